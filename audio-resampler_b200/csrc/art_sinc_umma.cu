/*
 * art_sinc_umma.cu -- the rational-ratio windowed-sinc kernel on the 5th-generation tensor cores
 * (tcgen05.mma, accumulators in tensor memory; sm_100a).
 *
 * Same reference functions as art_sinc_periodic.cu (resampler.c:523-526, :640-643, :1135-1157,
 * :1033-1044).  For ratio = L/M output n = L*q + j reads
 *
 *      y[q, j] = sum_k x[s_j + M*q + k] * h_j[k]
 *
 * which is a dense (periods x taps) by (taps x phases) product once the taps are counted from a
 * common origin s_0 ("flat" tap index f = k + s_j - s_0, 0 <= f < M + T): D[q, j] = sum_f A[q, f] * B[j, f]
 * with A[q, f] = x[s_0 + M*q + f] and B[j, f] = h_j[f - (s_j - s_0)] (zero outside the band).
 *
 *   operand A (signal)   The flat index is split as f = M*a + b: A[q, M*a + b] = x[s_0 + M*(q + a) + b] is row
 *                        q + a of the matrix X[r, b] = x[s_0 + M*r + b].  X is held ONCE in shared memory
 *                        (K-major, no swizzle, stored as 8-tap planes with rows at a 16-byte pitch) and the
 *                        row shift a is nothing but +16*a bytes on the descriptor's start address -- no
 *                        im2col copy of the overlapping windows is ever made.
 *   operand B (filters)  built once per launch by the prep kernel in exactly the shared-memory image of a
 *                        k-step (16 taps x Npad phases), streamed by TMA bulk copies through an 8-slot ring, two k-steps
 *                        per copy.
 *   exact accumulation   Tensor-memory accumulation truncates (measured on B200: -0.15 ulp per MMA, see
 *                        profiles/microbench/umma_probe.cu), which a chain of ~200 MMAs cannot afford at a
 *                        1e-6 bar.  Integer-valued fp16 operands whose sums stay below 2^24 accumulate
 *                        EXACTLY, so both operands are split in fixed point per tile:
 *                            x = qx * (X1 + 2^-11 x2),   h = qh * (H1 + 2^-11 h2 + 2^-22 h3)
 *                        X1, H1 integers of magnitude <= 2^11 (exact in fp16), x2/h2/h3 fp16 residuals.  Five
 *                        MMAs per k-step feed three accumulators: [X1*H1] exact; [X1*h2 + x2*H1] and
 *                        [X1*h3 + x2*h2] carry 2^-11 and 2^-22 of the weight, where truncation is harmless.
 *                        The epilogue adds them in fp32 and applies the power-of-two scale.
 *
 *   the prep kernel      (same launch sequence) writes the filter operand, the per-job origins, the new history, and the
 *                        block maximum of every tile's samples (-> q_x).
 *   generality           long periods: operand A becomes a ring of plane-pair slots; more than 160 phases: groups of
 *                        phases with their own tables; many interleaved channels: planar scratch (art_device.cu).
 *
 * Warp roles (one CTA of 640 threads per SM, persistent over tiles of 128 periods x 1 channel x 1 phase group):
 *   warp 0      TMA producer: filter k-steps -> stage ring (cp.async.bulk + mbarrier)
 *   warps 1-2   MMA issuers, alternating ring slots: 5 x tcgen05.mma (M=128, N=Npad, K=16) per k-step, tcgen05.commit
 *   warps 4-11  converters: global -> fixed-point split -> operand A, a pair of 8-tap planes at a time
 *   warps 12-19 epilogue: tcgen05.ld the three accumulators into registers (setmaxnreg gives them 120), release tensor
 *               memory, then transpose through shared memory and store
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include "art_kernels.cuh"
#include "art_device.h"

#define ART_U_THREADS 640           /* warpgroup 0: producer + 2 MMA warps (+1 idle); 1-2: converters; 3-4: epilogue */
#define ART_U_EPI     256           /* epilogue threads */
#define ART_U_CONV    256           /* converter threads */
#define ART_U_STAGES  8           /* filter ring depth (fewer when shared memory is short) */
#define ART_U_MAXKI   12            /* plane-pair slots (barriers) */
#define ART_U_MAXPAIRS 32           /* plane pairs per row: M <= 512 */
#define ART_U_GROUP   2             /* k-steps per ring slot: one wait / commit per slot */
#define ART_U_ROWS    128           /* periods per tile = M of the MMA */
#define ART_U_DX      11            /* signal digit: |X1| <= 2^11 */

__device__ __forceinline__ unsigned int u_smem (const void *p) { return (unsigned int) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void u_mbar_init (unsigned int bar, unsigned int count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void u_mbar_expect_tx (unsigned int bar, unsigned int bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void u_mbar_arrive (unsigned int bar)
{
    asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void u_mbar_wait (unsigned int bar, unsigned int parity)
{
    // the spin limit only turns a lost arrival (a bug) into a trap instead of a hung GPU
    for (unsigned int spins = 0; spins < (1u << 27); ++spins) {
        unsigned int done;
        asm volatile (
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    printf ("libresampler_b200: tensor-core pipeline stalled (block %d, thread %d)\n", blockIdx.x, threadIdx.x);
    __trap ();
}
/* for the roles that wait long (producer, converters, epilogue): sleep between polls, so that their spinning does not
 * take issue slots and shared-memory atomic bandwidth from the warps that are working */
__device__ __forceinline__ void u_mbar_wait_relaxed (unsigned int bar, unsigned int parity)
{
    for (unsigned int spins = 0; spins < (1u << 24); ++spins) {
        unsigned int done;
        asm volatile (
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        __nanosleep (200);
    }
    printf ("libresampler_b200: tensor-core pipeline stalled (block %d, thread %d)\n", blockIdx.x, threadIdx.x);
    __trap ();
}
__device__ __forceinline__ void u_bulk_g2s (unsigned int dstSmem, const void *srcGlobal, unsigned int bytes, unsigned int bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(dstSmem), "l"(srcGlobal), "r"(bytes), "r"(bar) : "memory");
}

/* shared-memory matrix descriptor: K-major, no swizzle; lbo = bytes between the two 8-tap planes of a
 * k-step, sbo = bytes between groups of 8 rows (128: rows sit at a uniform 16-byte pitch) */
__device__ __forceinline__ unsigned long long u_desc (unsigned int addr, unsigned int lbo, unsigned int sbo)
{
    return (unsigned long long) ((addr >> 4) & 0x3fff) | ((unsigned long long) ((lbo >> 4) & 0x3fff) << 16) |
           ((unsigned long long) ((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
/* instruction descriptor: fp32 accumulator, fp16 x fp16, both K-major, M x N */
__device__ __forceinline__ unsigned int u_idesc (int M, int N)
{
    return (1u << 4) | ((unsigned int) (N >> 3) << 17) | ((unsigned int) (M >> 4) << 24);
}
__device__ __forceinline__ void u_mma (unsigned int tmem, unsigned long long da, unsigned long long db, unsigned int idesc, unsigned int acc)
{
    asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                  :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
/* Shared memory is addressed through 32-bit shared-window addresses derived once from the block's base: a generic
 * pointer costs a read of the cluster CTA id (S2UR SR_CgaCtaId) at every use, which showed up in every store of the converters */
__device__ __forceinline__ void u_sts16 (unsigned int addr, unsigned short v) { asm volatile ("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ void u_sts32 (unsigned int addr, unsigned int v) { asm volatile ("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void u_stsf (unsigned int addr, float v) { asm volatile ("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float u_ldsf (unsigned int addr) { float v; asm volatile ("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ unsigned int u_lds32 (unsigned int addr) { unsigned int v; asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint2 u_lds64 (unsigned int addr) { uint2 v; asm volatile ("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void u_sts64 (unsigned int addr, uint2 v) { asm volatile ("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(v.x), "r"(v.y) : "memory"); }

__device__ __forceinline__ bool u_elect ()
{
    unsigned int pred;
    asm volatile ("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void u_commit (unsigned int bar)
{
    asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

/* optional role timing (ART_B200_UPROF=1): cycles spent waiting / working per role, summed over CTAs */
__device__ unsigned long long g_uprof[16];
#define UCLK() (prof ? clock64 () : 0ll)
#define UPROF_ADD(slot, cyc) do { if (prof) pacc[slot] += (unsigned int) (cyc); } while (0)      /* role-local, flushed once per thread */
#define UPROF_DECL()  unsigned int pacc[16] = { 0 }
#define UPROF_FLUSH() do { if (prof) { _Pragma ("unroll") for (int i_ = 0; i_ < 16; ++i_) if (pacc[i_]) atomicAdd (&g_uprof[i_], (unsigned long long) pacc[i_]); } } while (0)

__device__ __forceinline__ void u_tmem_ld8 (unsigned int addr, unsigned int (&r)[8])
{
    asm volatile ("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                  : "r"(addr));
}

__device__ __forceinline__ int u_find_job (const ArtJob *jobs, int numJobs, int tile)
{
    int lo = 0, hi = numJobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

/* ---- 1. prep: filter operand, per-job origins, history ------------------------------------------ */
/* One block per (table, phase).  Phase j's interpolated filter h_j (the same float values the FFMA form
 * uses: (float) (a + f (b - a)) in double, resampler.c:1155-1156 with the lerp folded into the taps) is
 * cut into fixed-point digits and written in the shared-memory image of every k-step. */
__global__ void __launch_bounds__ (128)
art_umma_prep_kernel (const ArtClass k, const ArtUmma u, const __grid_constant__ ArtJob single,
                      const ArtJob *__restrict__ jobs, int numJobs, int numTables, int histBlocksPerJob, int totalTiles, int dbg)
{
    const int tableBlocks = numTables * u.G * u.Npad;
    const int originBlocks = (numJobs * u.G + 127) / 128;
    int b = blockIdx.x;
    const int T = k.T, half = T / 2, F = k.F;
    if (b < tableBlocks) {
#ifdef ART_B200_ABLATE
        if (dbg & 16) return;
#endif
        const int tg = b / u.Npad, j = b - tg * u.Npad;                 // (table, phase group), phase inside the group
        const int tbl = tg / u.G, grp = tg - tbl * u.G;
        const int ph0 = grp * u.Lg, phj = ph0 + j;                      // the group's first phase, this phase
        const bool live = j < u.Lg && phj < u.L;
        const ArtJob &job = jobs ? jobs[jobs[tbl].repJob] : single;
        __shared__ int sh_row, sh_pass, sh_shift;
        __shared__ double sh_f;
        if (threadIdx.x == 0) {
            ArtLoopState st;
            st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = T;
            long long sj = 0, sb = 0;
            int row = 0, pass = -1;
            double f = 0.0;
            for (int which = 0; which < 2; ++which) {               // 0: the group's first phase (the origin), 1: this phase
                const int ph = which ? phj : ph0;
                if (which && !live) break;
                int w;
                const double pos = art_output_pos (&st, job.nStart + ph, &w);
                const double whole = floor (pos), fr = pos - whole;
                const long long s = (long long) whole - half + 1 + (long long) w * 15LL * T - job.origin;
                if (!which) { sb = s; continue; }
                sj = s;
                if (k.mode & ART_MODE_INTERP) {
                    double phs = fr * F;                             // resampler.c:1149-1152
                    row = (int) floor (phs);
                    f = phs - row;
                    if (row >= F) { row = F - 1; f = 1.0; }
                }
                else {
                    row = (int) floor (fr * F + 0.5);                // resampler.c:1137
                    if (!(k.mode & ART_MODE_LOWPASS) && row % F == 0)    // resampler.c:1141-1142
                        pass = half - 1 + (row ? 1 : 0);
                }
            }
            sh_row = row; sh_f = f; sh_pass = pass; sh_shift = (int) (sj - sb);
        }
        __syncthreads ();
        const int row = sh_row, pass = sh_pass, shift = sh_shift;
        const double f = sh_f;
        const float *ra = k.bank + (size_t) row * k.Tp, *rb = ra + k.Tp;
        unsigned short *tab = u.H + (size_t) tg * u.tableHalfs;
        const double qinv = (double) (1 << u.DH);
        for (int kk = threadIdx.x; kk < u.numK * 16; kk += 128) {
            const int ks = kk >> 4, e16 = kk & 15;
            const int bb = 16 * u.ki[ks] + e16;
            const int t = u.ka[ks] * u.M + bb - shift;
            float h = 0.0f;
            if (live && bb < u.M && t >= 0 && t < T) {
                if (pass >= 0)
                    h = t == pass ? 1.0f : 0.0f;
                else if (k.mode & ART_MODE_INTERP) {
                    const double a = ra[t], c = rb[t];
                    h = (float) (a + f * (c - a));
                }
                else
                    h = ra[t];
            }
            const double v = (double) h * qinv;
            const double H1 = rint (v);
            const double r1 = (v - H1) * 2048.0;
            const __half h2 = __double2half (r1);
            const double r2 = (r1 - (double) __half2float (h2)) * 2048.0;
            const __half h3 = __double2half (r2);
            const __half h1 = __double2half (H1);
            // [ks][split][plane][Npad][8]
            const size_t at = (((size_t) ks * 3 * 2 + (e16 >> 3)) * u.Npad + j) * 8 + (e16 & 7);
            const size_t splitStride = (size_t) 2 * u.Npad * 8;
            tab[at] = __half_as_ushort (h1);
            tab[at + splitStride] = __half_as_ushort (h2);
            tab[at + 2 * splitStride] = __half_as_ushort (h3);
        }
        return;
    }
    b -= tableBlocks;
    if (b < originBlocks) {
        const int e = b * 128 + threadIdx.x;
        if (e >= numJobs * u.G) return;
        const int seg = e / u.G, grp = e - seg * u.G;
        const ArtJob &job = jobs ? jobs[seg] : single;
        ArtLoopState st;
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = T;
        int w;
        const double pos = art_output_pos (&st, job.nStart + grp * u.Lg, &w);
        u.S0[e] = (int) ((long long) floor (pos) - half + 1 + (long long) w * 15LL * T - job.origin);
        return;
    }
    b -= originBlocks;
    if (b < totalTiles) {
        /* Block maximum of the samples a tile reads (-> its power-of-two quantum, |x| / 2^e <= 2^11, taken by the converters).
         * tileMax[] holds float bit patterns, zeroed before the launch, raised with atomicMax.  The C tiles of a period block
         * read the same frames: for interleaved input their C blocks each scan 1/C of the frames for ALL channels (every
         * sample read once, coalesced); otherwise each block scans its own channel.  (The origin is recomputed here: the
         * origin block may run later.) */
        const int tile = b;
        const int seg = jobs ? (numJobs > 1 ? u_find_job (jobs, numJobs, tile) : 0) : 0;
        if (threadIdx.x == 0) u.tileJob[tile] = seg;
        const ArtJob &job = jobs ? jobs[seg] : single;
        const int local = tile - job.tile0;
        const int C = k.C;
        const int qg = local / C, j = local - qg * C;                    // (period block, phase group), channel
        const int qb = qg / u.G, grp = qg - qb * u.G;
        ArtLoopState st;
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = T;
        int w;
        const double pos = art_output_pos (&st, job.nStart + grp * u.Lg, &w);
        const long long S0 = (long long) floor (pos) - half + 1 + (long long) w * 15LL * T - job.origin;
        const long long R0 = S0 + (long long) u.M * qb * 128;
        const int span = u.M * (u.rows - 1) + 16 * u.KI;
        // the part inside the caller's block, and the part that comes from the history; the rest is silence
        const long long lo = R0 > -job.prevAvail ? R0 : -job.prevAvail;
        const long long hi = R0 + span < (long long) job.inValid ? R0 + span : (long long) job.inValid;
        const long long hlo = R0 > -job.prevAvail - T ? R0 : -job.prevAvail - T;
        const long long hhi = R0 + span < -job.prevAvail ? R0 + span : -job.prevAvail;
        unsigned int *tileMax = reinterpret_cast<unsigned int *> (u.tileExp) + (tile - j);       // entries of this period block
        __shared__ unsigned int smaxC[128];

        // channel j's stretch of history (planar [C][T])
        float mh = 0.0f;
        {
            const float *hist = job.hist + (long long) j * T + T + job.prevAvail;
            for (long long i = hlo + threadIdx.x; i < hhi; i += 128) mh = fmaxf (mh, fabsf (hist[i]));
        }
        const bool interleaved = job.inPlanes == nullptr && job.inCS == 1 && job.inFS == C && C <= 128;
        if (interleaved) {
            const long long len = hi > lo ? hi - lo : 0;
            const long long s0 = lo + len * j / C, s1 = lo + len * (j + 1) / C;        // this block's frames
            smaxC[threadIdx.x] = 0u;
            __syncthreads ();
            if (mh > 0.0f) atomicMax (&smaxC[j], __float_as_uint (mh));
            const float *p0 = job.in + s0 * C, *p1 = job.in + s1 * C;
            if (C == 1 || C == 2 || C == 4) {
                // 16-byte loads over the aligned interior, eight in flight per thread; lane k of a vector is channel (off + k) % C
                float mk[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                const float *a0 = reinterpret_cast<const float *> ((reinterpret_cast<unsigned long long> (p0) + 15) & ~15ull);
                const float *a1 = reinterpret_cast<const float *> (reinterpret_cast<unsigned long long> (p1) & ~15ull);
                if (a1 < a0) { a0 = p1; a1 = p1; }
                for (const float *p = p0 + threadIdx.x; p < a0 && p < p1; p += 128) atomicMax (&smaxC[(int) ((p - p0) % C)], __float_as_uint (fabsf (*p)));
                for (const float *p = a1 + threadIdx.x; p < p1; p += 128) atomicMax (&smaxC[(int) ((p - p0) % C)], __float_as_uint (fabsf (*p)));
                const float4 *q = reinterpret_cast<const float4 *> (a0);
                const int n4 = (int) ((a1 - a0) >> 2);
                for (int i0 = threadIdx.x; i0 < n4; i0 += 8 * 128) {
                    float4 v[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int i = i0 + r * 128;
                        v[r] = i < n4 ? __ldg (q + i) : make_float4 (0.0f, 0.0f, 0.0f, 0.0f);
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        mk[0] = fmaxf (mk[0], fabsf (v[r].x)); mk[1] = fmaxf (mk[1], fabsf (v[r].y));
                        mk[2] = fmaxf (mk[2], fabsf (v[r].z)); mk[3] = fmaxf (mk[3], fabsf (v[r].w));
                    }
                }
                const int off = (int) ((a0 - p0) % C);
#pragma unroll
                for (int kq = 0; kq < 4; ++kq) {
                    float m = mk[kq];
#pragma unroll
                    for (int sh = 16; sh >= 1; sh >>= 1) m = fmaxf (m, __shfl_xor_sync (0xffffffffu, m, sh));
                    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax (&smaxC[(off + kq) % C], __float_as_uint (m));
                }
            }
            else if ((128 % C) == 0) {
                // thread t owns channel t % C and every (128 / C)-th frame: a warp reads consecutive floats; eight loads in flight
                const int c = threadIdx.x % C, fstep = 128 / C;
                float m = 0.0f;
                const float *base = job.in + c;
                for (long long i0 = s0 + threadIdx.x / C; i0 < s1; i0 += 8 * fstep) {
                    float v[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const long long i = i0 + (long long) r * fstep;
                        v[r] = i < s1 ? __ldg (base + i * C) : 0.0f;
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) m = fmaxf (m, fabsf (v[r]));
                }
                atomicMax (&smaxC[c], __float_as_uint (m));
            }
            else {
                const long long n = (s1 - s0) * C;
                for (long long e = threadIdx.x; e < n; e += 128) atomicMax (&smaxC[(int) (e % C)], __float_as_uint (fabsf (__ldg (p0 + e))));
            }
            __syncthreads ();
            if (threadIdx.x < C && smaxC[threadIdx.x]) atomicMax (&tileMax[threadIdx.x], smaxC[threadIdx.x]);
            return;
        }
        {
            float m = mh;
            const float *base = job.inPlanes ? job.inPlanes[j] : job.in + (long long) j * job.inCS;
            const long long fs = job.inFS;
            for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += 8 * 128) {
                float v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const long long i = i0 + r * 128;
                    v[r] = i < hi ? __ldg (base + i * fs) : 0.0f;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) m = fmaxf (m, fabsf (v[r]));
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) m = fmaxf (m, __shfl_xor_sync (0xffffffffu, m, off));
            if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax (&tileMax[j], __float_as_uint (m));
        }
        return;
    }
    b -= totalTiles;
    {
        const int seg = b / histBlocksPerJob, hb = b - seg * histBlocksPerJob;
        const ArtJob &job = jobs ? jobs[seg] : single;
        if (!job.histOut) return;
        const int total = k.C * T;
        for (int e = hb * 128 + threadIdx.x; e < total; e += histBlocksPerJob * 128) {
            const int c = e / T, i = e - c * T;
            job.histOut[e] = art_fetch (job, T, c, job.consumed - T + i);
        }
    }
}

/* ---- 2. the product ------------------------------------------------------------------------------- */
__global__ void __launch_bounds__ (ART_U_THREADS, 1)
art_sinc_umma_kernel (const ArtClass k, const __grid_constant__ ArtUmma u, const __grid_constant__ ArtJob single,
                      const ArtJob *__restrict__ jobs, int totalTiles, int profArg)
{
    const int prof = profArg & 1;
#ifdef ART_B200_ABLATE              /* measurement builds only: bits that SKIP work (wrong results by construction) */
    const int dbg = profArg >> 4;
#else
    constexpr int dbg = 0;
#endif
    extern __shared__ __align__ (1024) unsigned char smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = u.L, M = u.M, Npad = u.Npad, KI = u.KI, NS = u.NS, numK = u.numK, C = k.C, T = k.T;
    const unsigned int rep = (unsigned int) (KI / NS);    // uses of a plane-pair slot per tile
    const unsigned int planeBytes = (unsigned int) u.rows * 16u;
    const unsigned int splitBytes = planeBytes * 2u * (unsigned int) NS;
    const unsigned int stageBytes = 3u * 2u * (unsigned int) Npad * 16u;
    // [operand A: 2 splits][filter stage ring][epilogue transposition scratch][barriers and small tables]
    unsigned int xcBase;
    asm volatile ("mov.u32 %0, %1;" : "=r"(xcBase) : "r"(u_smem (smem)));       // opaque: never rematerialised from the generic pointer
    const unsigned int stBase = xcBase + 2u * splitBytes;
    const unsigned int scratchBase = stBase + (unsigned int) u.stages * stageBytes;
    const unsigned int ctl = scratchBase + 8u * 32u * 17u * 4u;
    const int units = u.stages / ART_U_GROUP;             // ring slots of ART_U_GROUP k-steps
#define hFullA(s)   (ctl + 8u * (unsigned int) (s))
#define hEmptyA(s)  (ctl + 64u + 8u * (unsigned int) (s))
#define pFullA(i)   (ctl + 128u + 8u * (unsigned int) (i))
#define pEmptyA(i)  (ctl + 224u + 8u * (unsigned int) (i))
#define accFullA    (ctl + 320u)
#define accEmptyA   (ctl + 328u)
#define tmemSlotA   (ctl + 336u)
#define goA(w)      (ctl + ((w) ? 368u : 344u))                         /* issue token of MMA issuer w */
#define sScaleA(j)  (ctl + 352u + 4u * (unsigned int) (j))
#define kTabA(ks)   (ctl + 384u + 8u * (unsigned int) (ks))          /* per k-step: low descriptor word of operand A, flags | plane pair << 8 */

    const int unitsPerTile = (numK + ART_U_GROUP - 1) / ART_U_GROUP;
    if (tid == 0) {
        for (int s = 0; s < units; ++s) { u_mbar_init (hFullA (s), 1); u_mbar_init (hEmptyA (s), 1); }
        for (int sl = 0; sl < NS; ++sl) {
            // a pair of planes is released by every issuer that used it (the planner guarantees that the pairs
            // sharing a slot agree on that)
            unsigned int who = 0;
            for (int ks = 0; ks < numK; ++ks)
                if (u.ki[ks] % NS == sl) who |= 1u << ((ks / ART_U_GROUP) & 1);
            u_mbar_init (pFullA (sl), ART_U_CONV / 32);
            u_mbar_init (pEmptyA (sl), who == 3u ? 2 : 1);
        }
        u_mbar_init (accFullA, unitsPerTile > 1 ? 2 : 1);
        u_mbar_init (accEmptyA, ART_U_EPI / 32);
        u_mbar_init (goA (0), 1);
        u_mbar_init (goA (1), 1);
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int ks = tid; ks < numK; ks += ART_U_THREADS) {
        const unsigned int i = u.ki[ks], a = u.ka[ks];
        const unsigned int aAddr = xcBase + 2u * (i % (unsigned int) NS) * planeBytes + 16u * a;      // row shift a = +16 bytes
        // bit 1: the last k-step of its pair of planes that this k-step's issuer handles
        const int mine = (ks / ART_U_GROUP) & 1;
        bool last = true;
        for (int k2 = ks + 1; k2 < numK && u.ki[k2] == i; ++k2)
            if (((k2 / ART_U_GROUP) & 1) == mine) last = false;
        u_sts64 (kTabA (ks), make_uint2 (((aAddr >> 4) & 0x3fffu) | (((planeBytes >> 4) & 0x3fffu) << 16),
                                         (a == 0 ? 1u : 0u) | (last ? 2u : 0u) | ((i % (unsigned int) NS) << 8) | ((i / (unsigned int) NS) << 16)));
    }
    if (warp == 0) {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmemSlotA), "r"(512));
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tm = u_lds32 (tmemSlotA);
    /* Registers follow the work (setmaxnreg, per warpgroup): the epilogue keeps a whole accumulator row block in registers so
     * that tensor memory is handed back to the MMAs before the slow part (transpose + stores) starts. 128*56 + 256*88 + 256*120 <= 640 * 96 */

    const ArtJob *const singlePtr = &single;
    auto jobOf = [=] (int tile) -> const ArtJob & {
        return jobs ? jobs[u.tileJob[tile]] : *singlePtr;
    };

    if (warp == 0) {
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
        /* ===== TMA producer: the tile's filter table, two k-steps per copy ===== */
        if (lane == 0) {
            UPROF_DECL ();
            unsigned int us = 0, ph = 0;
            for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x) {
                const ArtJob &job = jobOf (tile);
                const int grp = ((tile - job.tile0) / C) % u.G;
                const unsigned short *tab = u.H + ((size_t) job.table * u.G + grp) * u.tableHalfs;
                for (int ks = 0; ks < numK; ks += ART_U_GROUP) {
                    const unsigned int bytes = (unsigned int) (numK - ks < ART_U_GROUP ? numK - ks : ART_U_GROUP) * stageBytes;
                    long long t0 = UCLK ();
                    u_mbar_wait_relaxed (hEmptyA (us), ph ^ 1);
                    UPROF_ADD (0, UCLK () - t0);
                    if (dbg & 2) { u_mbar_arrive (hFullA (us)); }
                    else {
                        u_mbar_expect_tx (hFullA (us), bytes);
                        u_bulk_g2s (stBase + us * (ART_U_GROUP * stageBytes), tab + (size_t) ks * (stageBytes / 2), bytes, hFullA (us));
                    }
                    if (++us == (unsigned int) units) { us = 0; ph ^= 1; }
                }
            }
            UPROF_FLUSH ();
        }
    }
    else if (warp < 3) {
        /* ===== MMA issuers.  A single thread needs ~460 cycles to issue a k-step whose five MMAs run for 400, so two warps
         * share the work: ring slots (ART_U_GROUP k-steps) alternate between them.  Each whole warp walks the loop (uniform
         * control flow) and one elected lane issues.  Every tcgen05.commit costs the tensor pipe ~85 cycles (measured,
         * profiles/microbench/umma_probe.cu): one per slot, one per issuer and pair of planes, one per issuer and tile ===== */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
        const unsigned int me = (unsigned int) (warp - 1);
        UPROF_DECL ();
        const unsigned int idesc = u_idesc (ART_U_ROWS, Npad);
        const unsigned int bSplitU = (2u * (unsigned int) Npad * 16u) >> 4, stageU = stageBytes >> 4;
        const unsigned long long descHi = ((unsigned long long) (128u >> 4) << 32) | (1ull << 46);      // SBO = 128, version 1
        const unsigned int bLo0 = ((stBase >> 4) & 0x3fffu) | ((((unsigned int) Npad * 16u >> 4) & 0x3fffu) << 16);
        const unsigned int aSplitU = splitBytes >> 4;
        unsigned int us = 0, ph = 0, lt = 0, gw = 0;
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x, ++lt) {
            long long t1 = UCLK ();
            // the accumulators are free once the epilogue has drained them (the second issuer starts on the first one's token)
            if (me == 0) u_mbar_wait (accEmptyA, (lt & 1) ^ 1);
            long long t2 = UCLK ();
            if (lane == 0 && me == 0) UPROF_ADD (2, t2 - t1);
            asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned int unit = 0;
            for (int ks0 = 0; ks0 < numK; ks0 += ART_U_GROUP, ++unit) {
                if ((unit & 1u) == me) {
                    const int cnt = numK - ks0 < ART_U_GROUP ? numK - ks0 : ART_U_GROUP;
                    long long t0 = UCLK ();
                    for (int g = 0; g < cnt; ++g) {
                        const unsigned int f = u_lds64 (kTabA (ks0 + g)).y;
                        u_mbar_wait (pFullA ((f >> 8) & 0xffu), (lt * rep + (f >> 16)) & 1);   // the converters filled this pair of planes
                    }
                    long long t3 = UCLK ();
                    u_mbar_wait (hFullA (us), ph);
                    if (lane == 0 && me == 0) { UPROF_ADD (1, t3 - t0); UPROF_ADD (3, UCLK () - t3); }
                    // slots are issued strictly in order -- the two issuers pass a token -- so that every accumulator sees its
                    // MMAs in one fixed order: the truncating accumulation (classes 2 and 3) stays bit-reproducible
                    if (unit > 0) { u_mbar_wait (goA (me), gw & 1); ++gw; }
                    asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
                    long long t4 = UCLK ();
                    if (u_elect ()) {
                        for (int g = 0; g < cnt; ++g) {
                            const int ks = ks0 + g;
                            const uint2 kt = u_lds64 (kTabA (ks));
                            const unsigned int aLo = kt.x;
                            const unsigned int bLo = bLo0 + (us * ART_U_GROUP + (unsigned int) g) * stageU;
                            const unsigned long long dA1 = descHi | aLo, dA2 = descHi | (aLo + aSplitU);
                            const unsigned long long dB1 = descHi | bLo, dB2 = descHi | (bLo + bSplitU), dB3 = descHi | (bLo + 2u * bSplitU);
                            const unsigned int acc = ks > 0;
                            if (!(dbg & 8)) {
                            u_mma (tm, dA1, dB1, idesc, acc);                       // X1 * H1   (exact)
                            u_mma (tm + Npad, dA1, dB2, idesc, acc);                // X1 * h2
                            u_mma (tm + Npad, dA2, dB1, idesc, 1);                  // x2 * H1
                            u_mma (tm + 2 * Npad, dA1, dB3, idesc, acc);            // X1 * h3
                            u_mma (tm + 2 * Npad, dA2, dB2, idesc, 1);              // x2 * h2
                            }
                            if (kt.y & 2)
                                u_commit (pEmptyA ((kt.y >> 8) & 0xffu));            // this issuer is done with the pair of planes
                        }
                        long long t5 = UCLK ();
                        u_commit (hEmptyA (us));
                        if ((int) unit + 1 < unitsPerTile) u_mbar_arrive (goA (me ^ 1u));     // the next slot is the other issuer's
                        if (me == 0) { UPROF_ADD (11, t5 - t4); UPROF_ADD (12, UCLK () - t5); }
                    }
                    __syncwarp ();
                    if (lane == 0 && me == 0) UPROF_ADD (13, UCLK () - t4);
                }
                if (++us == (unsigned int) units) { us = 0; ph ^= 1; }
            }
            if (me < unit) {                                                     // this issuer had work in the tile
                if (u_elect ())
                    u_commit (accFullA);
                __syncwarp ();
            }
            if (lane == 0 && me == 0) UPROF_ADD (4, UCLK () - t2);
        }
        UPROF_FLUSH ();
    }
    else if (warp < 4) {
        /* idle: fills warpgroup 0 (register re-allocation is per warpgroup) */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
    }
    else if (warp < 12) {
        /* ===== converters: signal -> fixed-point operand A, one pair of 8-tap planes at a time ===== */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 88;");
        const int ctid = tid - 128, cw = ctid >> 5;
        UPROF_DECL ();
        const int span = M * (u.rows - 1) + 16 * KI;                       // samples the tile touches per channel
        /* a lane converts TWO neighbouring taps of a row per step (one conversion and one 32-bit store per digit for the two);
         * a warp covers 4 rows x 16 taps, the 8 warps take every 8th group of 4 rows */
        constexpr int UN = 5;                                               // row groups per warp: rows <= 160
        const int r0 = 4 * cw + (lane >> 3);                                // this lane's row in the warp's first group
        const int off0 = M * r0 + 2 * (lane & 7);                           // its first sample's offset inside the tile, plane pair 0
        unsigned int lt = 0;
        float v[2 * UN], vn[2 * UN];
        const int nU = ((u.rows + 3) / 4 - cw + 7) / 8;                     // row groups this warp owns

        /* what a lane needs to fetch its samples of one tile.  Inside the caller's block a sample is base[idx * fs];
         * tiles that touch the history or run past the end of the input take the same loads from a clamped address
         * and zero what must read as silence, so that all loads of a plane pair are still issued back to back */
        struct Src { const float *p, *p0, *h0, *dummy; long long fs; int loRel, hiRel, e; bool fast; };
        auto source = [&] (int tile) -> Src {
            Src sc;
            const ArtJob &job = jobOf (tile);
            const int seg = (int) (&job - (jobs ? jobs : singlePtr));
            const int local = tile - job.tile0;
            const int qg = local / C, c = local - qg * C;
            const int qb = qg / u.G, grp = qg - qb * u.G;
            const long long R0 = (long long) u.S0[seg * u.G + grp] + (long long) M * qb * ART_U_ROWS;
            const long long lo = -job.prevAvail, hi = job.inValid;
            sc.fast = R0 >= lo && R0 + span <= hi;
            sc.fs = job.inFS;
            const float *base = job.inPlanes ? job.inPlanes[c] : job.in + (long long) c * job.inCS;
            const float *hist = job.hist + (long long) c * T + T + job.prevAvail;     // hist[idx]: -T - prevAvail <= idx < -prevAvail
            // tile-relative coordinates (rel = idx - R0, an int): the boundary path classifies every sample it fetches
            const long long loR = lo - R0, hiR = hi - R0;
            sc.loRel = loR < -(1 << 30) ? -(1 << 30) : (loR > (1 << 30) ? (1 << 30) : (int) loR);
            sc.hiRel = hiR < -(1 << 30) ? -(1 << 30) : (hiR > (1 << 30) ? (1 << 30) : (int) hiR);
            sc.p0 = base + R0 * sc.fs;
            sc.h0 = hist + R0;
            sc.dummy = hist + lo - 1;                                         // any readable address: the newest history sample
            sc.p = sc.p0 + (long long) off0 * sc.fs;
            {
                const float m = __int_as_float (u.tileExp[tile]);             // block maximum (prep kernel)
                int e = 0;
                if (m > 0.0f) { (void) frexpf (m, &e); e -= ART_U_DX; e = e < -114 ? -114 : (e > 100 ? 100 : e); }      // m < 2^(e + 11)
                sc.e = e;
            }
            return sc;
        };
        auto rowOk = [&] (int uu) -> bool { return uu < nU && r0 + 32 * uu < u.rows; };
        auto fetch = [&] (const Src &sc, int i, float (&dst)[2 * UN]) {
            if (sc.fast) {
                const float *p = sc.p + (long long) (16 * i) * sc.fs;
                const long long rowStep = (long long) (32 * M) * sc.fs;
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    const bool ok = rowOk (uu) && !(dbg & 1);
                    dst[2 * uu] = ok ? __ldg (p) : 0.0f;
                    dst[2 * uu + 1] = ok ? __ldg (p + sc.fs) : 0.0f;
                    p += rowStep;
                }
            }
            else {
                const float *ptr[2 * UN];
                bool ok[2 * UN];
#pragma unroll
                for (int e = 0; e < 2 * UN; ++e) {
                    const int rel = off0 + 16 * i + 32 * M * (e >> 1) + (e & 1);
                    const bool inBlock = rel >= sc.loRel && rel < sc.hiRel, inHist = rel < sc.loRel && rel >= sc.loRel - T;
                    ok[e] = (inBlock || inHist) && rowOk (e >> 1);
                    ptr[e] = inBlock ? sc.p0 + (long long) rel * sc.fs : (inHist ? sc.h0 + rel : sc.dummy);
                }
#pragma unroll
                for (int e = 0; e < 2 * UN; ++e) dst[e] = __ldg (ptr[e]);
#pragma unroll
                for (int e = 0; e < 2 * UN; ++e) dst[e] = ok[e] ? dst[e] : 0.0f;
            }
        };

        Src cur = source (blockIdx.x < totalTiles ? blockIdx.x : 0);
        if ((int) blockIdx.x < totalTiles) fetch (cur, 0, vn);
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x, ++lt) {
            const int e = cur.e;
            const float invq = __int_as_float ((127 - e) << 23);
            if (ctid == 0)
                u_stsf (sScaleA (lt & 3), __int_as_float ((127 + e) << 23) * __int_as_float ((127 - u.DH) << 23));
            const bool more = tile + (int) gridDim.x < totalTiles;
            Src nxt = cur;
            for (int i = 0; i < KI; ++i) {
#pragma unroll
                for (int uu = 0; uu < 2 * UN; ++uu) v[uu] = vn[uu];
                // the next pair's loads (of the next tile after the last pair) fly while this pair is converted
                // (the next tile's job lookup is a chain of dependent global loads: done early, at the first pair, where
                //  the converters have a whole tile of slack, not in front of the last pair the MMAs are waiting for)
                if (i == 0 && more) nxt = source (tile + gridDim.x);
                if (i + 1 < KI) fetch (cur, i + 1, vn);
                else if (more) fetch (nxt, 0, vn);
                long long c0t = UCLK ();
                const unsigned int sl = (unsigned int) i % (unsigned int) NS, use = lt * rep + (unsigned int) i / (unsigned int) NS;
                u_mbar_wait_relaxed (pEmptyA (sl), (use & 1) ^ 1);
                long long cb = UCLK ();
                if (ctid == 0) UPROF_ADD (5, cb - c0t);
                // taps 2c, 2c+1 of the pair: plane 2i + (c >> 2), 4 bytes at (c & 3) * 4 of the row's 16-byte slot
                const unsigned int dst = xcBase + (2u * sl + ((lane >> 2) & 1)) * planeBytes + (unsigned int) (lane & 3) * 4u +
                                         (unsigned int) r0 * 16u;
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    if (rowOk (uu)) {
                        const float ua = v[2 * uu] * invq, ub = v[2 * uu + 1] * invq;
                        const float Xa = (ua + 12582912.0f) - 12582912.0f;          // round to nearest integer (|u| <= 2^11)
                        const float Xb = (ub + 12582912.0f) - 12582912.0f;
                        const float ra = (ua - Xa) * 2048.0f, rb = (ub - Xb) * 2048.0f;
                        const __half2 p1 = __floats2half2_rn (Xa, Xb), p2 = __floats2half2_rn (ra, rb);      // low half = first tap
                        u_sts32 (dst + uu * 512, *reinterpret_cast<const unsigned int *> (&p1));
                        u_sts32 (dst + splitBytes + uu * 512, *reinterpret_cast<const unsigned int *> (&p2));
                    }
                }
                long long cc = UCLK ();
                asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core reads
                long long cd = UCLK ();
                __syncwarp ();
                if (lane == 0) u_mbar_arrive (pFullA (sl));                          // one arrival per warp: 256 serialised arrivals cost ~500 cycles
                if (ctid == 0) { UPROF_ADD (7, cc - cb); UPROF_ADD (14, cd - cc); UPROF_ADD (15, UCLK () - cd); }
            }
            cur = nxt;
        }
        UPROF_FLUSH ();
    }
    else {
        /* ===== epilogue: accumulators -> registers (tensor memory released) -> transposed through shared memory -> output ===== */
        asm volatile ("setmaxnreg.inc.sync.aligned.u32 120;");
        const int quad = warp & 3;                                          // TMEM lanes 32*quad .. 32*quad+31 (a warp may only read quarter warp%4)
        UPROF_DECL ();
        const int slot = (warp - 12) >> 2;                                  // every quarter is served by two warps: even / odd 16-column chunks
        const unsigned int tmRow = tm + ((unsigned int) (quad * 32) << 16);
        const unsigned int scratch = scratchBase + (unsigned int) (warp - 12) * (32u * 17u * 4u);
        const int half16 = lane >> 4, col = lane & 15;
        constexpr int NCH = 5;                                              // chunks per warp: Npad <= 160
        unsigned int lt = 0;
        for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x, ++lt) {
            const ArtJob &job = jobOf (tile);
            const int local = tile - job.tile0;
            const int qg = local / C, c = local - qg * C;
            const int qb = qg / u.G, grp = qg - qb * u.G;
            const int phases = min (u.Lg, L - grp * u.Lg);                    // phases of this group
            long long e0 = UCLK ();
            u_mbar_wait_relaxed (accFullA, lt & 1);
            long long e1 = UCLK ();
            asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float scale = u_ldsf (sScaleA (lt & 3));
            float y[NCH][16];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c0 = 16 * slot + 32 * ch;
                if (c0 < Npad && !(dbg & 4)) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {                        // 8 columns at a time: 24 transient registers
                        unsigned int a1[8], a2[8], a3[8];
                        u_tmem_ld8 (tmRow + (unsigned int) (c0 + 8 * hh), a1);
                        u_tmem_ld8 (tmRow + (unsigned int) (Npad + c0 + 8 * hh), a2);
                        u_tmem_ld8 (tmRow + (unsigned int) (2 * Npad + c0 + 8 * hh), a3);
                        asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            y[ch][8 * hh + j] = ((__uint_as_float (a3[j]) * (1.0f / 4194304.0f) + __uint_as_float (a2[j]) * (1.0f / 2048.0f)) +
                                                 __uint_as_float (a1[j])) * scale;
                    }
                }
            }
            // the accumulators may be overwritten by the next tile from here on
            asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp ();
            if (lane == 0) u_mbar_arrive (accEmptyA);
            long long e2 = UCLK ();

            // job fields into registers: the output stores below may alias anything as far as the compiler knows
            const long long outputs = job.outputs, outFS = job.outFS;
            float *const obase = (job.outPlanes ? job.outPlanes[c] : job.out + (long long) c * job.outCS) + (long long) job.nStart * outFS;
            const long long q0 = (long long) qb * ART_U_ROWS + quad * 32;    // period of this warp's row 0
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c0 = 16 * slot + 32 * ch;
                if (c0 < Npad && !(dbg & 4)) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        u_stsf (scratch + (unsigned int) (lane * 17 + j) * 4u, y[ch][j]);
                    __syncwarp ();
                    const int ph = c0 + col;                                 // this lane's phase
                    // half-warps take rows rr and rr + 16: with the row pitch of 17 words their 16 columns fall on disjoint banks
                    long long nl = (q0 + 16 * half16) * L + grp * u.Lg + ph;  // output index inside the job, rows advance by 1
                    float *op = obase + nl * outFS;
                    const unsigned int sp = scratch + (unsigned int) (16 * half16 * 17 + col) * 4u;
                    const long long step = L, ostep = step * outFS;
                    if (ph < phases) {
#pragma unroll 8
                        for (int rr = 0; rr < 16; ++rr) {
                            const float v = u_ldsf (sp + (unsigned int) rr * (17u * 4u));
                            if (nl < outputs) *op = v;
                            nl += step; op += ostep;
                        }
                    }
                    __syncwarp ();
                }
            }
            if (tid == 12 * 32) { UPROF_ADD (8, e1 - e0); UPROF_ADD (9, e2 - e1); UPROF_ADD (6, UCLK () - e2); UPROF_ADD (10, 1); }
        }
        UPROF_FLUSH ();
    }

    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (warp == 0) {
        __syncwarp ();
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

/* ---- host side -------------------------------------------------------------------------------------- */

static size_t umma_smem (const ArtUmma &u)
{
    const size_t xc = (size_t) 2 * (2 * u.NS) * u.rows * 16;                 // two splits
    return xc + (size_t) u.stages * 3 * 2 * u.Npad * 16 + 8 * 32 * 17 * sizeof (float) + 384 + ART_U_MAXK * 8;
}

bool artPlanUmma (const ArtClass &k, double ratio, unsigned int maxOutputs, unsigned long long totalOutputs,
                  int smCount, ArtUmma &u)
{
    if (g_artTensorMode < 0) {
        const char *e = getenv ("ART_B200_UMMA");
        g_artTensorMode = e ? atoi (e) : 1;
    }
    const int enabled = g_artTensorMode;
    if (!enabled) return false;
    if (k.mode & ART_MODE_PRECISE) return false;                  // double accumulation: generic kernel
    // the block-scaled fixed point makes the last bit depend on where a tile starts: contexts that promise
    // chunking-invariant output (no interpolation, resampler.c:1135-1145) keep the FFMA form unless the caller
    // trades that promise for speed (mode 3)
    if (!(k.mode & ART_MODE_INTERP) && enabled < 3) return false;
    int L, M;
    if (!artRational (ratio, 1024, &L, &M)) return false;
    // short periods are grouped: g periods of L outputs form one row of g*L phases
    int g = L <= 160 ? 160 / L : 1;
    while (g > 1 && (long long) M * g > 176) --g;
    L *= g; M *= g;
    if (L < 48 || M > 16 * ART_U_MAXPAIRS) return false;
    if (maxOutputs < (unsigned) (16 * L)) return false;           // rows of a tile would be mostly idle
    // a launch costs this kernel ~25 us whatever its size (filter table + one tile per SM); the FFMA form runs at
    // ~11 Gsamples/s on a single stream, so it wins below ~0.3 Msamples
    (void) smCount;
    if (enabled != 2 && totalOutputs * (unsigned long long) k.C < 300000ull) return false;

    memset (&u, 0, sizeof u);
    u.L = L; u.M = M;
    // tensor memory holds 3 accumulators of at most 160 columns: more phases are handled in G groups, each with its own table
    u.G = (L + 159) / 160;
    u.Lg = (L + u.G - 1) / u.G;
    u.Npad = (u.Lg + 15) & ~15;
    u.KI = (M + 15) / 16;
    if (u.KI > ART_U_MAXPAIRS) return false;
    const int flatEnd = (int) (((long long) u.Lg * M + L - 1) / L) + 2 + k.T;    // taps are counted from the first tap of a group's first phase
    int n = 0, aMax = 0;
    for (int i = 0; i < u.KI; ++i)                                // plane pair outermost: its planes are handed
        for (int a = 0; a * M + 16 * i < flatEnd; ++a) {          // back to the converters after the last shift
            if (n >= ART_U_MAXK) return false;
            u.ka[n] = (unsigned char) a; u.ki[n] = (unsigned char) i; ++n;
            if (a > aMax) aMax = a;
        }
    u.numK = n;
    for (int i = 0; i < u.KI; ++i) {
        int c = 0;
        for (int a = 0; a * M + 16 * i < flatEnd; ++a) ++c;
        u.nA[i] = (unsigned char) c;
    }
    // rows: 128 + aMax, even, and 2..6 (mod 8) so that the 16-byte row slots of neighbouring planes fall on different banks
    int rows = ART_U_ROWS + aMax;
    while ((rows & 1) || (rows & 7) < 2 || (rows & 7) > 6) ++rows;
    if (rows > 144) return false;
    u.rows = rows;
    // filter quantum: the exact accumulator holds sum X1*H1 with |X1| <= 2^11 and sum |H1| <= absSum * 2^DH + T/2
    u.DH = 0;
    for (int dh = 11; dh >= 6; --dh)
        if ((double) k.absSum * (double) (1 << dh) + 0.5 * k.T < 8191.0) { u.DH = dh; break; }
    if (!u.DH) return false;
    u.tableHalfs = u.numK * 3 * 2 * u.Npad * 8;
    u.stages = ART_U_STAGES;
    // operand A: all KI plane pairs of a row if they fit, else a ring of NS | KI slots -- pairs are used in order, each
    // for all of its row shifts in a row, so a slot can take pair i + NS as soon as the MMAs of pair i are done.  Slots
    // shared by several pairs need those pairs to be released by the same issuers (both, i.e. >= 3 k-steps per pair).
    u.NS = 0;
    for (int ns = u.KI < ART_U_MAXKI ? u.KI : ART_U_MAXKI; ns >= 1; --ns) {
        if (u.KI % ns) continue;
        u.NS = ns;
        if (umma_smem (u) <= 224 * 1024) break;
        u.NS = 0;
    }
    if (!u.NS) return false;
    if (u.NS < u.KI)
        for (int i = 0; i < u.KI; ++i)
            if (u.nA[i] < 3) return false;
    while (u.stages > 2 * ART_U_GROUP && umma_smem (u) > 224 * 1024) u.stages -= ART_U_GROUP;
    if (umma_smem (u) > 224 * 1024) return false;
    if (getenv ("ART_B200_TRACE"))
        fprintf (stderr, "[art] umma L=%d M=%d G=%d Npad=%d KI=%d NS=%d numK=%d rows=%d DH=%d stages=%d smem=%zu\n",
                 u.L, u.M, u.G, u.Npad, u.KI, u.NS, u.numK, u.rows, u.DH, u.stages, umma_smem (u));
    return true;
}

int artUmmaTiles (const ArtUmma &u, int channels, unsigned int outputs)
{
    const long long Q = ((long long) outputs + u.L - 1) / u.L;
    return (int) (((Q + ART_U_ROWS - 1) / ART_U_ROWS) * channels * u.G);
}

static size_t umma_align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

size_t artUmmaTableBytes (const ArtUmma &u, int numTables, int numJobs, int totalTiles)
{
    return umma_align16 ((size_t) numTables * u.G * u.tableHalfs * sizeof (unsigned short)) + umma_align16 ((size_t) numJobs * u.G * sizeof (int)) +
           2 * (size_t) totalTiles * sizeof (int);
}

void artUmmaCarve (ArtUmma &u, void *tables, int numTables, int numJobs)
{
    unsigned char *p = reinterpret_cast<unsigned char *> (tables);
    u.H = reinterpret_cast<unsigned short *> (p);
    p += umma_align16 ((size_t) numTables * u.G * u.tableHalfs * sizeof (unsigned short));
    u.S0 = reinterpret_cast<int *> (p);
    p += umma_align16 ((size_t) numJobs * u.G * sizeof (int));
    u.tileExp = reinterpret_cast<int *> (p);
    u.tileJob = nullptr;            // follows tileExp: set by the launcher, which knows the tile count
}

void artLaunchUmma (const ArtClass &k, const ArtUmma &u, int totalTiles, int numJobs, int numTables, int smCount,
                    const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    if (totalTiles <= 0) return;
    static bool configured[16] = { false };
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    if (!configured[device & 15]) {
        // the kernel moves registers between its warpgroups (setmaxnreg: 128 threads down to 56, 256 down to 88, 256 up to 120);
        // the pool that comes from is threads x registers-per-thread as compiled, so check that it suffices -- a warpgroup
        // asking for registers that never become free would spin forever
        cudaFuncAttributes fa;
        ART_CUDA_CHECK (cudaFuncGetAttributes (&fa, art_sinc_umma_kernel));
        if (128 * (fa.numRegs - 56) + 256 * (fa.numRegs - 88) < 256 * (120 - fa.numRegs) || fa.numRegs < 88 || fa.numRegs > 120) {
            artRaise ("art_sinc_umma_kernel was compiled with %d registers per thread: its register re-allocation plan does not hold", fa.numRegs);
        }
        ART_CUDA_CHECK (cudaFuncSetAttribute (art_sinc_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        configured[device & 15] = true;
    }
    int histBlocks = (k.C * k.T + 127) / 128;
    if (histBlocks > 32) histBlocks = 32;
    const int prepBlocks = numTables * u.G * u.Npad + (numJobs * u.G + 127) / 128 + totalTiles + numJobs * histBlocks;
    ArtUmma uu = u;
    uu.tileJob = u.tileExp + totalTiles;
    ART_CUDA_CHECK (cudaMemsetAsync (u.tileExp, 0, (size_t) totalTiles * sizeof (int), stream));
    static int prepDbg = -1;
    if (prepDbg < 0) {
        prepDbg = 0;
#ifdef ART_B200_ABLATE
        if (const char *d = getenv ("ART_B200_UDBG")) prepDbg = atoi (d);
#endif
    }
    art_umma_prep_kernel<<<prepBlocks, 128, 0, stream>>> (k, uu, single, d_jobs, numJobs, numTables, histBlocks, totalTiles, prepDbg);
    ART_CUDA_CHECK (cudaGetLastError ());
    const int grid = totalTiles < smCount ? totalTiles : smCount;
    static int roleProf = -1;
    if (roleProf < 0) {
        roleProf = getenv ("ART_B200_UPROF") ? 1 : 0;
#ifdef ART_B200_ABLATE
        if (const char *d = getenv ("ART_B200_UDBG")) roleProf |= atoi (d) << 4;
#endif
        if (roleProf & 1) atexit ([] () {
            unsigned long long h[16];
            cudaDeviceSynchronize ();
            if (cudaMemcpyFromSymbol (h, g_uprof, sizeof h) != cudaSuccess) return;
            const double n = h[10] ? (double) h[10] : 1.0;
            fprintf (stderr, "[art] umma cycles per tile: producer wait-empty %.0f | mma wait-planes %.0f wait-acc %.0f wait-h %.0f tile %.0f | "
                     "convert wait-planes %.0f | epilogue wait %.0f drain %.0f | tiles %.0f | mma issue %.0f commit %.0f region %.0f | epilogue store %.0f | convert split %.0f fence %.0f arrive %.0f\n",
                     h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[8] / n, h[9] / n, n, h[11] / n, h[12] / n, h[13] / n,
                     h[6] / n, h[7] / n, h[14] / n, h[15] / n);
        });
    }
    void *prof;
    artProfileBegin (stream, &prof);
    art_sinc_umma_kernel<<<grid, ART_U_THREADS, umma_smem (u), stream>>> (k, uu, single, d_jobs, totalTiles, roleProf);
    artProfileEnd (stream, prof);
    ART_CUDA_CHECK (cudaGetLastError ());
    g_artLaunches += 2;
}
