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

#if !ART_WIDE       /* a float-path optimisation: the wide (PATH_WIDTH=64) build keeps to the any-ratio kernel */

#define ART_U_THREADS 640           /* warpgroup 0: producer + 2 MMA warps (+1 idle); 1-2: converters; 3-4: epilogue */
#define ART_U_EPI     256           /* epilogue threads */
#define ART_U_CONV    256           /* converter threads */
#define ART_U_STAGES  8           /* filter ring depth (fewer when shared memory is short) */
#define ART_U_MAXKI   12            /* plane-pair slots (barriers) */
#define ART_U_MAXPAIRS 32           /* plane pairs per row: M <= 512 */
#define ART_U_GROUP   2             /* k-steps per ring slot: one wait / commit per slot */
#define ART_U_ROWS    128           /* periods per tile = M of the MMA */
#define ART_U_DX      11            /* signal digit: |X1| <= 2^11 */
#define ART_U_SCRATCH (32u * 9u * 4u)  /* epilogue transposition scratch per warp: 32 rows x 8 columns, pitch 9 words */

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

/* raw-sample staging of the converters: per ring slot (= one pair of planes) [period group][tap][converter thread] x CGT floats */
__host__ __device__ constexpr unsigned int u_staging_groups (int cgt) { return cgt == 1 ? 5u : (cgt == 2 ? 3u : 2u); }
__host__ __device__ constexpr unsigned int u_staging_slot (int cgt) { return u_staging_groups (cgt) * 2u * ART_U_CONV * 4u * (unsigned int) cgt; }

__device__ __forceinline__ void u_cp_async (unsigned int dst, const void *src, int bytes, bool live)
{
    // src-size 0 writes zeros: samples that must read as silence cost the same instruction as real ones
    const unsigned int n = live ? (unsigned int) bytes : 0u;
    if (bytes == 16)     asm volatile ("cp.async.ca.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
    else if (bytes == 8) asm volatile ("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
    else                 asm volatile ("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
}

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
__device__ unsigned long long g_uprof[24];
__device__ unsigned long long g_utime[160][4];         // per CTA (last launch): globaltimer at entry, first MMA issue, last accFull commit, exit
__device__ __forceinline__ unsigned long long u_gtime () { unsigned long long t; asm volatile ("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define UCLK() (prof ? clock64 () : 0ll)
#define UPROF_ADD(slot, cyc) do { if (prof) pacc[slot] += (unsigned int) (cyc); } while (0)      /* role-local, flushed once per thread */
#define UPROF_DECL()  unsigned int pacc[24] = { 0 }
#define UPROF_FLUSH() do { if (prof) { _Pragma ("unroll") for (int i_ = 0; i_ < 24; ++i_) if (pacc[i_]) atomicAdd (&g_uprof[i_], (unsigned long long) pacc[i_]); } } while (0)

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

/* Block maximum of the samples one tile reads (-> its quantum), this warp's share: warp `wi` of `nw` takes every nw-th
 * 512-byte run.  Non-finite samples are left out: they then convert to Inf / NaN digits and poison exactly the outputs whose
 * windows hold them, as they do in the reference's float arithmetic.  BATCH 16-byte loads are in flight per lane. */
template <int BATCH>
__device__ __forceinline__ float u_scan_tile (const ArtJob &job, int CGT, int c0, long long R0, long long R1, int T, int wi, int nw, int lane)
{
    const float inf = __int_as_float (0x7f800000);
    float m = 0.0f;
#define U_TAKE(x) do { const float v_ = fabsf (x); m = fmaxf (m, v_ < inf ? v_ : 0.0f); } while (0)
    const long long lo = -job.prevAvail, hi = job.inValid;
    const long long a = R0 > lo ? R0 : lo, b = R1 < hi ? R1 : hi;
    const int stride = nw * 32;
    auto span = [&] (const float *base, long long n) {                     // n contiguous floats, all of them the tile's
        long long head = (long long) (((16u - (unsigned int) (reinterpret_cast<unsigned long long> (base) & 15u)) & 15u) >> 2);
        if (head > n) head = n;
        const long long n4 = (n - head) >> 2;
        const float4 *q = reinterpret_cast<const float4 *> (base + head);
        for (long long i0 = wi * 32 + lane; i0 < n4; i0 += (long long) BATCH * stride) {
            float4 v[BATCH];
#pragma unroll
            for (int r = 0; r < BATCH; ++r) {
                const long long i = i0 + (long long) r * stride;
                v[r] = i < n4 ? __ldg (q + i) : make_float4 (0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int r = 0; r < BATCH; ++r) { U_TAKE (v[r].x); U_TAKE (v[r].y); U_TAKE (v[r].z); U_TAKE (v[r].w); }
        }
        if (wi == 0) {                                                      // the unaligned ends
            for (long long e = lane; e < head; e += 32) U_TAKE (__ldg (base + e));
            for (long long e = head + 4 * n4 + lane; e < n; e += 32) U_TAKE (__ldg (base + e));
        }
    };
    if (b > a) {
        const long long fs = job.inFS;
        if (!job.inPlanes && job.inCS == 1 && fs == CGT)                    // the frames hold exactly the tile's channels
            span (job.in + c0 + a * fs, (b - a) * fs);
        else if (!job.inPlanes && job.inCS == 1 && CGT > 1 && (fs % CGT) == 0 &&
                 (reinterpret_cast<unsigned long long> (job.in + c0) & (4u * CGT - 1u)) == 0) {
            // an interleaved block of more channels than the tile takes: one vector per frame, eight frames in flight per lane
            const float *base = job.in + c0 + a * fs;
            const long long n = b - a;
            for (long long i0 = wi * 32 + lane; i0 < n; i0 += 8LL * stride) {
                float4 v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const long long i = i0 + (long long) r * stride;
                    if (i >= n) v[r] = make_float4 (0.0f, 0.0f, 0.0f, 0.0f);
                    else if (CGT == 4) v[r] = __ldg (reinterpret_cast<const float4 *> (base + i * fs));
                    else { const float2 t = __ldg (reinterpret_cast<const float2 *> (base + i * fs)); v[r] = make_float4 (t.x, t.y, 0.0f, 0.0f); }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) { U_TAKE (v[r].x); U_TAKE (v[r].y); U_TAKE (v[r].z); U_TAKE (v[r].w); }
            }
        }
        else
            for (int cc = 0; cc < CGT; ++cc) {
                const float *base = (job.inPlanes ? job.inPlanes[c0 + cc] : job.in + (long long) (c0 + cc) * job.inCS) + a * fs;
                if (fs == 1) span (base, b - a);
                else
                    for (long long i = wi * 32 + lane; i < b - a; i += stride) U_TAKE (__ldg (base + i * fs));
            }
    }
    {   // the stretch that comes from the history (tiles at the start of a call)
        const long long hlo = R0 > lo - T ? R0 : lo - T, hhi = R1 < lo ? R1 : lo;
        for (int cc = 0; cc < CGT; ++cc) {
            const float *hist = job.hist + (long long) (c0 + cc) * T + T + job.prevAvail;
            for (long long i = hlo + wi * 32 + lane; i < hhi; i += stride) U_TAKE (hist[i]);
        }
    }
#undef U_TAKE
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) m = fmaxf (m, __shfl_xor_sync (0xffffffffu, m, sh));
    return m;
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
    const int T = k.T, Tref = k.Tref, half = Tref / 2 + k.lead, F = k.F;          // T taps per row; positions run on Tref
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
            st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = Tref;
            long long sj = 0, sb = 0;
            int row = 0, pass = -1;
            double f = 0.0;
            for (int which = 0; which < 2; ++which) {               // 0: the group's first phase (the origin), 1: this phase
                const int ph = which ? phj : ph0;
                if (which && !live) break;
                int w;
                const double pos = art_output_pos (&st, job.nStart + ph, &w);
                const double whole = floor (pos), fr = pos - whole;
                const long long s = (long long) whole - half + 1 + (long long) w * 15LL * Tref - job.origin;
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
                        pass = half - 1 + (row ? 1 : 0);                 // (never with a folded-in pre-filter: that sets ART_MODE_LOWPASS)
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
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = Tref;
        int w;
        const double pos = art_output_pos (&st, job.nStart + grp * u.Lg, &w);
        u.S0[e] = (int) ((long long) floor (pos) - half + 1 + (long long) w * 15LL * Tref - job.origin);
        return;
    }
    b -= originBlocks;
    const int tileJobBlocks = (totalTiles + 127) / 128;
    if (b < tileJobBlocks) {
        // tile -> job: the product kernel's roles then find their job with one load instead of a binary search
        const int tile = b * 128 + threadIdx.x;
        if (tile < totalTiles)
            u.tileJob[tile] = jobs ? (numJobs > 1 ? u_find_job (jobs, numJobs, tile) : 0) : 0;
        return;
    }
    b -= tileJobBlocks;
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
/* CGT = channels per tile (1, 2 or 4).  A tile's 128 MMA rows are ART_U_ROWS / CGT periods x CGT channels with the channel
 * fastest: physical row CGT * r + cc of the signal operand is period r of channel c0 + cc, and the row shift a of a k-step is
 * +16 * CGT * a bytes on the descriptor.  The converters then read all CGT channels of a frame with ONE vector load (an
 * interleaved stereo block wastes no half sectors), and the epilogue's stores cover whole frames. */
template <int CGT, bool VEC>
__global__ void __launch_bounds__ (ART_U_THREADS, 1)
art_sinc_umma_kernel (const ArtClass k, const __grid_constant__ ArtUmma u, const __grid_constant__ ArtJob single,
                      const ArtJob *__restrict__ jobs, int totalTiles, int profArg)
{
#ifdef ART_B200_ABLATE              /* measurement builds only: role counters (ART_B200_UPROF) and bits that SKIP work (wrong results by construction) */
    const int prof = profArg & 1;
    const int dbg = (profArg >> 4) & 255;
#else
    constexpr int prof = 0, dbg = 0;
#endif
    extern __shared__ __align__ (1024) unsigned char smem[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (prof && tid == 0 && blockIdx.x < 160) g_utime[blockIdx.x][0] = u_gtime ();
    const int L = u.L, M = u.M, Npad = u.Npad, KI = u.KI, NS = u.NS, numK = u.numK, C = k.C, T = k.T;
    /* This CTA's i-th tile (-1: no more).  Channel groups of one frame range are neighbours in the launch and want to run at the
     * same time on different SMs (they share sectors and DRAM pages: BASELINE config 3 loses a third otherwise), which is what
     * round-robin (blockIdx + i * grid) gives.  When a tile holds all channels there is nothing to share, and round-robin has a
     * trap: it hands a CTA the same KIND of tile every time whenever the tiles per job divide the grid -- with BASELINE config 4's
     * four tiles per 2^15-frame block on 148 = 4 x 37 CTAs a quarter of the SMs got nothing but the tiles that start at the
     * history and ran 50 % longer than the rest.  Those launches give every CTA a contiguous range of tiles instead. */
    // tile i of this CTA = tStart + i * tStep, i < tCount
    const bool contiguous = (profArg & 4) != 0 && C == CGT;          // the host's choice (short jobs), see artLaunchUmma
    const int tileLo = (int) (((long long) totalTiles * blockIdx.x) / gridDim.x), tileHi = (int) (((long long) totalTiles * (blockIdx.x + 1)) / gridDim.x);
    const int tStart = contiguous ? tileLo : (int) blockIdx.x;
    const int tStep = contiguous ? 1 : (int) gridDim.x;
    const int tCount = contiguous ? tileHi - tileLo : (totalTiles - (int) blockIdx.x + (int) gridDim.x - 1) / (int) gridDim.x;
    const int CG = C / CGT;                                 // channel groups: tiles per (period block, phase group)
    constexpr int PR = ART_U_ROWS / CGT;                    // periods per tile
    const int digits = u.digits;                            // signal digits: 2, or 3 (one more MMA per k-step)
    const unsigned int rep = (unsigned int) (KI / NS);    // uses of a plane-pair slot per tile
    const unsigned int planeBytes = (unsigned int) u.rows * 16u;
    const unsigned int splitBytes = planeBytes * 2u * (unsigned int) NS;
    const unsigned int stageBytes = 3u * 2u * (unsigned int) Npad * 16u;
    // [operand A: `digits` splits][filter stage ring][epilogue transposition scratch][barriers and small tables]
    unsigned int xcBase;
    asm volatile ("mov.u32 %0, %1;" : "=r"(xcBase) : "r"(u_smem (smem)));       // opaque: never rematerialised from the generic pointer
    const unsigned int stBase = xcBase + (unsigned int) digits * splitBytes;
    const unsigned int scratchBase = stBase + (unsigned int) u.stages * stageBytes;
    const unsigned int stagingBase = scratchBase + 8u * ART_U_SCRATCH;          // converters' ring of raw samples (cp.async), see below
    const unsigned int ctl = stagingBase + (unsigned int) u.depth * u_staging_slot (CGT);
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
#define scanFullA(s)  (ctl + 384u + 8u * ART_U_MAXK + 8u * (unsigned int) (s))          /* tile maxima, ring of 4: written by the epilogue warps ... */
#define scanEmptyA(s) (ctl + 384u + 8u * ART_U_MAXK + 32u + 8u * (unsigned int) (s))    /* ... three tiles ahead, read by the converters */
#define sPartA(s, w)  (ctl + 384u + 8u * ART_U_MAXK + 64u + 32u * (unsigned int) (s) + 4u * (unsigned int) (w))

    const int unitsPerTile = (numK + ART_U_GROUP - 1) / ART_U_GROUP;
    if (tid == 0) {
        // a ring slot is free when BOTH issuers' MMAs that read it have completed (each commits to it)
        for (int s = 0; s < units; ++s) { u_mbar_init (hFullA (s), 1); u_mbar_init (hEmptyA (s), 2); }
        for (int sl = 0; sl < NS; ++sl) {
            u_mbar_init (pFullA (sl), ART_U_CONV / 32);
            u_mbar_init (pEmptyA (sl), 1);                 // released by the relay warp once the slot's last k-step has completed
        }
        u_mbar_init (accFullA, 2);
        u_mbar_init (accEmptyA, ART_U_EPI / 32);
        for (int sl = 0; sl < 4; ++sl) { u_mbar_init (scanFullA (sl), ART_U_EPI / 32); u_mbar_init (scanEmptyA (sl), ART_U_CONV / 32); }
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int ks = tid; ks < numK; ks += ART_U_THREADS) {
        const unsigned int i = u.ki[ks], a = u.ka[ks];
        const unsigned int aAddr = xcBase + 2u * (i % (unsigned int) NS) * planeBytes + 16u * (unsigned int) CGT * a;      // row shift a = +16 * CGT bytes
        // bit 1: the last k-step of its pair of planes (the k-steps of a pair are consecutive)
        const bool last = ks + 1 >= numK || u.ki[ks + 1] != i;
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
    /* tile -> (period block, phase group, first channel): tiles of a job run channel group fastest, then phase group */
    struct Where { int seg, qb, grp, c0; };
    auto whereIs = [=] (const ArtJob &job, int tile) -> Where {
        Where w;
        w.seg = (int) (&job - (jobs ? jobs : singlePtr));
        const int local = tile - job.tile0;
        const int qg = local / CG;
        w.c0 = (local - qg * CG) * CGT;
        w.qb = qg / u.G;
        w.grp = qg - w.qb * u.G;
        return w;
    };
    const int periods = u.periods;                                          // periods of a tile incl. the row shifts: PR + aMax
    const int span = M * (periods - 1) + 16 * KI;                           // samples the tile touches per channel

    if (warp == 0) {
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
        /* ===== TMA producer: the tile's filter table, two k-steps per copy ===== */
        if (lane == 0) {
            UPROF_DECL ();
            unsigned int us = 0, ph = 0;
            for (int it = 0, tile = tStart; it < tCount; ++it, tile += tStep) {
                const ArtJob &job = jobOf (tile);
                const int grp = ((tile - job.tile0) / CG) % u.G;
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
        /* ===== MMA issuers.  One thread needs 70-90 cycles to issue a tcgen05.mma that executes in 80, so a single issuer
         * leaves the tensor pipe waiting.  Two warps issue CONCURRENTLY, split by accumulator: issuer 0 feeds [X1*H1] and
         * [X1*h2 + x2*H1], issuer 1 feeds [X1*h3 + x2*h2 (+ x3*H1)].  Every accumulator sees its MMAs from one thread in program
         * order, so the truncating accumulation of the lower classes stays bit-reproducible without any hand-off between
         * the two.  Each whole warp walks the loop (uniform control flow) and one elected lane issues.  tcgen05.commit is not
         * free (~85 cycles of the pipe, profiles/microbench/umma_probe.cu): one per issuer and ring slot, one per issuer and
         * tile; plane pairs are released by the relay warp below instead ===== */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
        const unsigned int me = (unsigned int) (warp - 1);
        UPROF_DECL ();
        const unsigned int idesc = u_idesc (ART_U_ROWS, Npad);
        const unsigned int bSplitU = (2u * (unsigned int) Npad * 16u) >> 4, stageU = stageBytes >> 4;
        const unsigned long long descHi = ((unsigned long long) (128u >> 4) << 32) | (1ull << 46);      // SBO = 128, version 1
        const unsigned int bLo0 = ((stBase >> 4) & 0x3fffu) | ((((unsigned int) Npad * 16u >> 4) & 0x3fffu) << 16);
        const unsigned int aSplitU = splitBytes >> 4;
        unsigned int us = 0, ph = 0, lt = 0;
        for (int tile = tStart; (int) lt < tCount; ++lt, tile += tStep) {
            long long t1 = UCLK ();
            u_mbar_wait (accEmptyA, (lt & 1) ^ 1);                            // the epilogue has drained the accumulators
            long long t2 = UCLK ();
            if (lane == 0 && me == 0) UPROF_ADD (2, t2 - t1);
            asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ks0 = 0; ks0 < numK; ks0 += ART_U_GROUP) {
                const int cnt = numK - ks0 < ART_U_GROUP ? numK - ks0 : ART_U_GROUP;
                long long t0 = UCLK ();
                for (int g = 0; g < cnt; ++g) {
                    const unsigned int f = u_lds64 (kTabA (ks0 + g)).y;
                    u_mbar_wait (pFullA ((f >> 8) & 0xffu), (lt * rep + (f >> 16)) & 1);   // the converters filled this pair of planes
                }
                long long t3 = UCLK ();
                u_mbar_wait (hFullA (us), ph);
                if (lane == 0 && me == 0) { UPROF_ADD (1, t3 - t0); UPROF_ADD (3, UCLK () - t3); }
                asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
                long long t4 = UCLK ();
                if (prof && me == 0 && lane == 0 && lt == 0 && ks0 == 0 && blockIdx.x < 160) g_utime[blockIdx.x][1] = u_gtime ();
                if (u_elect ()) {
                    for (int g = 0; g < cnt; ++g) {
                        const int ks = ks0 + g;
                        const unsigned int aLo = u_lds32 (kTabA (ks));
                        const unsigned int bLo = bLo0 + (us * ART_U_GROUP + (unsigned int) g) * stageU;
                        const unsigned long long dA1 = descHi | aLo, dA2 = descHi | (aLo + aSplitU);
                        const unsigned long long dB1 = descHi | bLo, dB2 = descHi | (bLo + bSplitU);
                        const unsigned int acc = ks > 0;
                        if (!(dbg & 8)) {
                            if (me == 0) {
                                u_mma (tm, dA1, dB1, idesc, acc);                       // X1 * H1   (exact)
                                u_mma (tm + Npad, dA1, dB2, idesc, acc);                // X1 * h2
                                u_mma (tm + Npad, dA2, dB1, idesc, 1);                  // x2 * H1
                            }
                            else {
                                u_mma (tm + 2 * Npad, dA1, descHi | (bLo + 2u * bSplitU), idesc, acc);      // X1 * h3
                                u_mma (tm + 2 * Npad, dA2, dB2, idesc, 1);              // x2 * h2
                                if (digits == 3)
                                    u_mma (tm + 2 * Npad, descHi | (aLo + 2u * aSplitU), dB1, idesc, 1);    // x3 * H1
                            }
                        }
                    }
                    long long t5 = UCLK ();
                    u_commit (hEmptyA (us));
                    if (me == 0) { UPROF_ADD (11, t5 - t4); UPROF_ADD (12, UCLK () - t5); }
                }
                __syncwarp ();
                if (lane == 0 && me == 0) UPROF_ADD (13, UCLK () - t4);
                if (++us == (unsigned int) units) { us = 0; ph ^= 1; }
            }
            if (u_elect ())
                u_commit (accFullA);
            __syncwarp ();
            if (prof && me == 0 && lane == 0 && blockIdx.x < 160) g_utime[blockIdx.x][2] = u_gtime ();
            if (lane == 0 && me == 0) UPROF_ADD (4, UCLK () - t2);
        }
        UPROF_FLUSH ();
    }
    else if (warp < 4) {
        /* ===== relay: a pair of planes may be overwritten once the MMAs of its last k-step have completed, which the ring
         * slot's barrier already reports (both issuers commit to it).  This warp follows those completions in order and
         * passes them on to the converters -- a plain mbarrier arrive instead of ~20 more tcgen05.commit per tile ===== */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0) {
            unsigned int us = 0, ph = 0;
            for (int it = 0, tile = tStart; it < tCount; ++it, tile += tStep) {
                for (int ks0 = 0; ks0 < numK; ks0 += ART_U_GROUP) {
                    const int cnt = numK - ks0 < ART_U_GROUP ? numK - ks0 : ART_U_GROUP;
                    u_mbar_wait (hEmptyA (us), ph);
                    for (int g = 0; g < cnt; ++g) {
                        const unsigned int f = u_lds64 (kTabA (ks0 + g)).y;
                        if (f & 2u) u_mbar_arrive (pEmptyA ((f >> 8) & 0xffu));
                    }
                    if (++us == (unsigned int) units) { us = 0; ph ^= 1; }
                }
            }
        }
    }
    else if (warp < 12) {
        /* ===== converters: signal -> fixed-point operand A, one pair of 8-tap planes at a time ===== */
        asm volatile ("setmaxnreg.dec.sync.aligned.u32 88;");
        const int ctid = tid - 128, cw = ctid >> 5;
        UPROF_DECL ();
        /* a lane converts TWO neighbouring taps of a period, for all CGT channels of the tile (one conversion and one 32-bit
         * store per digit and channel for the two); a warp covers 4 periods x 16 taps, the 8 warps take every 8th group of 4 */
        constexpr int UN = (int) u_staging_groups (CGT);                    // period groups per warp: periods <= 160 / 96 / 64
        constexpr int NV = 2 * UN * CGT;
        const int r0 = 4 * cw + (lane >> 3);                                // this lane's period in the warp's first group
        const int off0 = M * r0 + 2 * (lane & 7);                           // its first sample's offset inside the tile, plane pair 0
        unsigned int lt = 0;

        /* The samples reach the converters through a ring of `depth` raw pairs in shared memory filled by cp.async: a lane
         * requests ITS samples of the pair `depth - 1` iterations ahead (no registers are tied up while they fly -- with loads
         * into registers only one pair could be in flight, and the ~1400 cycles a pair's loads take were exposed ten times per
         * tile) and reads them back itself, so the ring needs no synchronisation beyond cp.async.wait_group. */
        const int depth = u.depth;
        constexpr unsigned int planeStride = ART_U_CONV * 4u * CGT, slotBytes = u_staging_slot (CGT);
        const unsigned int stg = stagingBase + (unsigned int) ctid * (4u * CGT);

        /* what a lane needs to request its samples of one tile.  Inside the caller's block a sample is base[idx * fs];
         * tiles that touch the history or run past the end of the input classify every sample (history / silence) */
        struct Src { const float *pc[CGT]; const float *h0, *dummy; long long fs; int loRel, hiRel; bool fast, vec; };
        auto source = [&] (int tile) -> Src {
            Src sc;
            const ArtJob &job = jobOf (tile);
            const Where w = whereIs (job, tile);
            const long long R0 = (long long) u.S0[w.seg * u.G + w.grp] + (long long) M * w.qb * PR;
            const long long lo = -job.prevAvail, hi = job.inValid;
            sc.fast = R0 >= lo && R0 + span <= hi;
            sc.fs = job.inFS;
#pragma unroll
            for (int cc = 0; cc < CGT; ++cc)
                sc.pc[cc] = (job.inPlanes ? job.inPlanes[w.c0 + cc] : job.in + (long long) (w.c0 + cc) * job.inCS) + R0 * sc.fs;
            // all channels of a frame from one copy: interleaved with the tile's channels adjacent and the vector aligned
            sc.vec = CGT > 1 && job.inPlanes == nullptr && job.inCS == 1 && (sc.fs % CGT) == 0 &&
                     (reinterpret_cast<unsigned long long> (job.in + w.c0) & (4u * CGT - 1u)) == 0;
            const float *hist = job.hist + (long long) w.c0 * T + T + job.prevAvail;     // hist[idx]: -T - prevAvail <= idx < -prevAvail (+ cc * T)
            // tile-relative coordinates (rel = idx - R0, an int): the boundary path classifies every sample it fetches
            const long long loR = lo - R0, hiR = hi - R0;
            sc.loRel = loR < -(1 << 30) ? -(1 << 30) : (loR > (1 << 30) ? (1 << 30) : (int) loR);
            sc.hiRel = hiR < -(1 << 30) ? -(1 << 30) : (hiR > (1 << 30) ? (1 << 30) : (int) hiR);
            sc.h0 = hist + R0;
            sc.dummy = hist + lo - 1;                                         // any readable address: the newest history sample
            if (!sc.fast) {
                // a tile that touches the history or the end of the input (every tile of a short block): the classification is
                // only needed by the warps whose rows actually reach outside the block -- decided once per tile and warp
                bool plain = true;
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    const int rel = off0 + 32 * M * uu;                       // this lane's first sample of pair 0; its last: + 16 * KI - 1
                    plain = plain && (r0 + 32 * uu >= periods || (rel >= sc.loRel && rel + 16 * KI <= sc.hiRel));
                }
                sc.fast = __all_sync (0xffffffffu, plain);
            }
            return sc;
        };
        /* the tile's block maximum -> its power-of-two quantum.  The epilogue warps scan every tile three tiles ahead of its
         * conversion (each leaves its share of the maximum in the ring slot) */
        auto quantum = [&] (unsigned int kk) -> int {
            const unsigned int slot = kk & 3u;
            u_mbar_wait_relaxed (scanFullA (slot), (kk >> 2) & 1u);
            unsigned int bits = lane < ART_U_EPI / 32 ? u_lds32 (sPartA (slot, lane)) : 0u;
#pragma unroll
            for (int sh = 4; sh >= 1; sh >>= 1) bits = max (bits, __shfl_xor_sync (0xffffffffu, bits, sh));
            bits = __shfl_sync (0xffffffffu, bits, 0);
            __syncwarp ();
            if (lane == 0) u_mbar_arrive (scanEmptyA (slot));
            const float m = __uint_as_float (bits);
            int e = 0;
            if (m > 0.0f) { (void) frexpf (m, &e); e -= ART_U_DX; e = e < -114 ? -114 : (e > 116 ? 116 : e); }      // m < 2^(e + 11): every finite float fits
            return e;
        };
        unsigned int okMask = 0;
#pragma unroll
        for (int uu = 0; uu < UN; ++uu) okMask |= (r0 + 32 * uu < periods ? 1u : 0u) << uu;
        auto rowOk = [&] (int uu) -> bool { return (okMask >> uu) & 1u; };
        // request this lane's samples of plane pair i of the tile `sc` describes into ring slot `slot`
        auto request = [&] (const Src &sc, int i, unsigned int slot) {
            const unsigned int sb = stg + slot * slotBytes;
            if (dbg & 1) return;
            if (sc.fast) {
                // tile-relative element offsets fit 32 bits (a tile spans ~10^4 frames)
                const int fs = (int) sc.fs, at = (off0 + 16 * i) * fs, rowStep = 32 * M * fs;
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    if (!rowOk (uu)) continue;
                    const int o = at + uu * rowStep;
                    if (VEC) {                                              // (host-checked for every job of the launch)
                        u_cp_async (sb + (unsigned int) (2 * uu) * planeStride, sc.pc[0] + o, 4 * CGT, true);
                        u_cp_async (sb + (unsigned int) (2 * uu + 1) * planeStride, sc.pc[0] + o + sc.fs, 4 * CGT, true);
                    }
                    else {
#pragma unroll
                        for (int cc = 0; cc < CGT; ++cc) {
                            u_cp_async (sb + (unsigned int) (2 * uu) * planeStride + 4u * cc, sc.pc[cc] + o, 4, true);
                            u_cp_async (sb + (unsigned int) (2 * uu + 1) * planeStride + 4u * cc, sc.pc[cc] + o + sc.fs, 4, true);
                        }
                    }
                }
            }
            else {
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    if (!rowOk (uu)) continue;
#pragma unroll
                    for (int e = 0; e < 2 * CGT; ++e) {
                        const int tap = e / CGT, cc = e - tap * CGT;
                        const int rel = off0 + 16 * i + 32 * M * uu + tap;
                        const bool inBlock = rel >= sc.loRel && rel < sc.hiRel, inHist = rel < sc.loRel && rel >= sc.loRel - T;
                        const float *ptr = inBlock ? sc.pc[cc] + (long long) rel * sc.fs : (inHist ? sc.h0 + (long long) cc * T + rel : sc.dummy);
                        u_cp_async (sb + (unsigned int) (2 * uu + tap) * planeStride + 4u * cc, ptr, 4, inBlock || inHist);
                    }
                }
            }
        };
        // the request side runs `depth - 1` pairs ahead of the conversion, across tile boundaries
        int rit = 0, rtile = tStart, ri = 0;                                 // rit < tCount: a tile is left to request
        unsigned int rslot = 0;
        Src rsrc = source (tCount > 0 ? rtile : 0);
        auto requestNext = [&] () {
            if (rit < tCount) {
                request (rsrc, ri, rslot);
                if (++ri == KI) {
                    ri = 0;
                    ++rit; rtile += tStep;
                    if (rit < tCount) rsrc = source (rtile);
                }
            }
            asm volatile ("cp.async.commit_group;" ::: "memory");           // one group per pair, empty or not: the wait below counts groups
            if (++rslot == (unsigned int) depth) rslot = 0;
        };
        for (int d = 0; d + 1 < depth; ++d) requestNext ();
        unsigned int cslot = 0;
        unsigned int useBase = 0;                                           // uses of operand slot 0 before this tile

        for (int tile = tStart; (int) lt < tCount; ++lt, tile += tStep) {
            long long q0t = UCLK ();
            const int e = quantum (lt);
            if (ctid == 0) UPROF_ADD (17, UCLK () - q0t);
            const float invq = __int_as_float ((127 - e) << 23);
            const unsigned long long invq2 = art_pack2 (invq, invq), magic2 = art_pack2 (12582912.0f, 12582912.0f),
                                     neg2 = art_pack2 (-1.0f, -1.0f), k2048 = art_pack2 (2048.0f, 2048.0f);
            if (ctid == 0)
                u_stsf (sScaleA (lt & 3), __int_as_float ((127 + e) << 23) * __int_as_float ((127 - u.DH) << 23));
            unsigned int sl = 0, use = useBase;                           // operand slot of pair i = i % NS, and how often it was used before
            for (int i = 0; i < KI; ++i) {
                long long h0t = UCLK ();
                requestNext ();
                // all but the newest depth - 1 groups have landed: this pair's samples are in its slot
                if (depth == 4)      asm volatile ("cp.async.wait_group 3;" ::: "memory");
                else if (depth == 3) asm volatile ("cp.async.wait_group 2;" ::: "memory");
                else                 asm volatile ("cp.async.wait_group 1;" ::: "memory");
                const unsigned int sb = stg + cslot * slotBytes;
                if (++cslot == (unsigned int) depth) cslot = 0;
                float v[NV];                                                 // [(group * 2 + tap) * CGT + channel]
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
#pragma unroll
                    for (int tap = 0; tap < 2; ++tap) {
                        const unsigned int at = sb + (unsigned int) (2 * uu + tap) * planeStride;
                        if (CGT == 1) v[2 * uu + tap] = rowOk (uu) ? u_ldsf (at) : 0.0f;
                        else if (CGT == 2) {
                            const uint2 w2 = rowOk (uu) ? u_lds64 (at) : make_uint2 (0u, 0u);
                            v[(2 * uu + tap) * CGT] = __uint_as_float (w2.x); v[(2 * uu + tap) * CGT + (CGT > 1)] = __uint_as_float (w2.y);
                        }
                        else {
                            const uint2 w2 = rowOk (uu) ? u_lds64 (at) : make_uint2 (0u, 0u), w3 = rowOk (uu) ? u_lds64 (at + 8u) : make_uint2 (0u, 0u);
                            v[(2 * uu + tap) * CGT] = __uint_as_float (w2.x); v[(2 * uu + tap) * CGT + (CGT > 1)] = __uint_as_float (w2.y);
                            v[(2 * uu + tap) * CGT + 2 * (CGT > 2)] = __uint_as_float (w3.x); v[(2 * uu + tap) * CGT + 3 * (CGT > 2)] = __uint_as_float (w3.y);
                        }
                    }
                }
                long long c0t = UCLK ();
                if (ctid == 0) UPROF_ADD (18, c0t - h0t);
                u_mbar_wait_relaxed (pEmptyA (sl), (use & 1) ^ 1);
                long long cb = UCLK ();
                if (ctid == 0) UPROF_ADD (5, cb - c0t);
                // taps 2c, 2c+1 of the pair: plane 2i + (c >> 2), 4 bytes at (c & 3) * 4 of the row's 16-byte slot; row = CGT * period + channel
                const unsigned int dst = xcBase + (2u * sl + ((lane >> 2) & 1)) * planeBytes + (unsigned int) (lane & 3) * 4u +
                                         (unsigned int) (CGT * r0) * 16u;
#pragma unroll
                for (int uu = 0; uu < UN; ++uu) {
                    if (rowOk (uu) && !(dbg & 128)) {
#pragma unroll
                        for (int cc = 0; cc < CGT; ++cc) {
                            // both taps at once in packed fp32 (fma.rn.f32x2): u = x / q;  X = rint (u) by the 1.5 * 2^23 trick (|u| <= 2^11);
                            // r = (u - X) * 2^11.  Every step is exact or rounds exactly as the scalar form would (x / q is a power-of-two scaling)
                            const unsigned long long x2 = art_pack2 (v[(2 * uu) * CGT + cc], v[(2 * uu + 1) * CGT + cc]);
                            unsigned long long t2 = magic2, nX = magic2, d2 = 0ull, r2 = 0ull;
                            art_ffma2 (t2, x2, invq2);                                   // u + magic
                            art_ffma2 (nX, t2, neg2);                                    // magic - (u + magic) = -X
                            d2 = nX; art_ffma2 (d2, x2, invq2);                          // u - X
                            art_ffma2 (r2, d2, k2048);                                   // (u - X) * 2^11
                            float nXa, nXb, ra, rb;
                            art_unpack2 (nX, nXa, nXb);
                            art_unpack2 (r2, ra, rb);
                            const __half2 n1 = __floats2half2_rn (nXa, nXb), p2 = __floats2half2_rn (ra, rb);     // low half = first tap
                            const unsigned int at = dst + (unsigned int) (uu * 512 * CGT + cc * 16);
                            u_sts32 (at, *reinterpret_cast<const unsigned int *> (&n1) ^ 0x80008000u);            // X = -(-X)
                            u_sts32 (at + splitBytes, *reinterpret_cast<const unsigned int *> (&p2));
                            if (digits == 3) {
                                // third digit: what the fp16 rounding of the second one dropped (11 more bits): every sample keeps >= 22
                                // significant bits whatever the tile's maximum is
                                const float2 back = __half22float2 (p2);
                                unsigned long long e2 = r2, r3 = 0ull;
                                art_ffma2 (e2, art_pack2 (back.x, back.y), neg2);          // r - fp16 (r)
                                art_ffma2 (r3, e2, k2048);
                                float ea, eb;
                                art_unpack2 (r3, ea, eb);
                                const __half2 p3 = __floats2half2_rn (ea, eb);
                                u_sts32 (at + 2u * splitBytes, *reinterpret_cast<const unsigned int *> (&p3));
                            }
                        }
                    }
                }
                long long cc_ = UCLK ();
                if (!(dbg & 64))
                    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
                long long cd = UCLK ();
                __syncwarp ();
                if (lane == 0) u_mbar_arrive (pFullA (sl));                          // one arrival per warp: 256 serialised arrivals cost ~500 cycles
                if (ctid == 0) { UPROF_ADD (7, cc_ - cb); UPROF_ADD (14, cd - cc_); UPROF_ADD (15, UCLK () - cd); }
                if (++sl == (unsigned int) NS) { sl = 0; ++use; }
            }
            useBase += rep;
        }
        asm volatile ("cp.async.wait_all;" ::: "memory");
        UPROF_FLUSH ();
    }
    else {
        /* ===== epilogue: accumulators -> registers (tensor memory released) -> transposed through shared memory -> output ===== */
        asm volatile ("setmaxnreg.inc.sync.aligned.u32 120;");
        const int quad = warp & 3;                                          // TMEM lanes 32*quad .. 32*quad+31 (a warp may only read quarter warp%4)
        UPROF_DECL ();
        const int slot = (warp - 12) >> 2;                                  // every quarter is served by two warps: even / odd 16-column chunks
        const unsigned int tmRow = tm + ((unsigned int) (quad * 32) << 16);
        const unsigned int scratch = scratchBase + (unsigned int) (warp - 12) * ART_U_SCRATCH;
        constexpr int NCH = 5;                                              // chunks per warp: Npad <= 160
        /* ---- the tile scanner: block maximum of the samples tile number kk of this CTA will read.  A separate pass over the input
         * in front of the kernel costs more than these eight warps lose to it (both measured, round 2); they look three tiles
         * ahead, which also pulls that tile's samples into L2 before the converters ask for them ---- */
        const int ew = warp - 12;
        auto scanTile = [&] (int tile, unsigned int kk) {
            const unsigned int sslot = kk & 3u;
            u_mbar_wait_relaxed (scanEmptyA (sslot), ((kk >> 2) & 1u) ^ 1u);
            const ArtJob &job = jobOf (tile);
            const Where w = whereIs (job, tile);
            const long long R0 = (long long) u.S0[w.seg * u.G + w.grp] + (long long) M * w.qb * PR;
            const float m = u_scan_tile<20> (job, CGT, w.c0, R0, R0 + span, T, ew, ART_U_EPI / 32, lane);
            if (lane == 0) {
                u_sts32 (sPartA (sslot, ew), __float_as_uint (m));
                u_mbar_arrive (scanFullA (sslot));
            }
            __syncwarp ();
        };
        for (int ahead = 0; ahead < 3; ++ahead)
            if (ahead < tCount)
                scanTile (tStart + ahead * tStep, (unsigned int) ahead);
        // store mapping after the transposition: a lane owns channel lane % CGT and phase (lane / CGT) % 8 of period lane / (8 * CGT):
        // one store instruction covers 4 / CGT periods x 8 phases x CGT channels, i.e. whole frames, 32 bytes or more per period
        const int scc = lane % CGT, sjj = (lane / CGT) & 7, spp = lane / (8 * CGT);
        constexpr int PPI = 4 / CGT;
        unsigned int lt = 0;
        for (int tile = tStart; (int) lt < tCount; ++lt, tile += tStep) {
            const ArtJob &job = jobOf (tile);
            const Where w = whereIs (job, tile);
            const int grp = w.grp;
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

            // Output addressing in 32-bit arithmetic relative to one 64-bit pointer per tile: element (period p, phase ph) of this
            // lane's channel sits at tb[(p * L + ph) * outFS]; `room` is how many of the job's outputs lie at or after the tile's first
            const long long first = ((long long) w.qb * PR + (quad * 32) / CGT) * L + (long long) grp * u.Lg;      // job-relative index of (warp's period 0, phase 0)
            const long long left = (long long) job.outputs - first;
            const int room = left < 0 ? 0 : (left > 0x3fffffff ? 0x3fffffff : (int) left);
            const int ofs = (int) job.outFS;
            float *const tb = (job.outPlanes ? job.outPlanes[w.c0 + scc] : job.out + (long long) (w.c0 + scc) * job.outCS) +
                              ((long long) job.nStart + first) * job.outFS;
            const int stepN = PPI * L;                                        // outputs between the periods of successive store instructions
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int c0 = 16 * slot + 32 * ch;
                if (c0 < Npad && !(dbg & 4)) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            u_stsf (scratch + (unsigned int) (lane * 9 + j) * 4u, y[ch][8 * hh + j]);
                        __syncwarp ();
                        const int ph = c0 + 8 * hh + sjj;                    // this lane's phase
                        const unsigned int sp = scratch + (unsigned int) ((CGT * spp + scc) * 9 + sjj) * 4u;
                        float vv[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it)                       // 32 / CGT periods per warp, PPI per instruction
                            vv[it] = u_ldsf (sp + (unsigned int) it * (unsigned int) (CGT * PPI * 9 * 4));
                        int nl = spp * L + ph;                                // output index relative to `first`
                        if (ph < phases && !(dbg & 32)) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                if (nl < room) tb[nl * ofs] = vv[it];
                                nl += stepN;
                            }
                        }
                        __syncwarp ();
                    }
                }
            }
            long long e3 = UCLK ();
            if ((int) lt + 3 < tCount)
                scanTile (tile + 3 * tStep, lt + 3u);
            if (tid == 12 * 32) { UPROF_ADD (8, e1 - e0); UPROF_ADD (9, e2 - e1); UPROF_ADD (6, e3 - e2); UPROF_ADD (10, 1); UPROF_ADD (16, UCLK () - e3); }
        }
        UPROF_FLUSH ();
    }

    asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads ();
    if (prof && tid == 0 && blockIdx.x < 160) g_utime[blockIdx.x][3] = u_gtime ();
    if (warp == 0) {
        __syncwarp ();
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
    }
}

/* ---- host side -------------------------------------------------------------------------------------- */

int g_artTensorDigits = -1;        // signal digits of the tensor-core form: 3 (default) or 2; ART_B200_DIGITS sets the initial value

static size_t umma_smem (const ArtUmma &u)
{
    const size_t xc = (size_t) u.digits * (2 * u.NS) * u.rows * 16;          // one split per signal digit
    return xc + (size_t) u.stages * 3 * 2 * u.Npad * 16 + 8 * ART_U_SCRATCH + (size_t) u.depth * u_staging_slot (u.cg) +
           384 + ART_U_MAXK * 8 + 64 + 4 * 32;
}

#define ART_U_SMEM_MAX (227 * 1024)

bool artPlanUmma (const ArtClass &k, double ratio, unsigned int maxOutputs, unsigned long long totalOutputs,
                  int smCount, ArtUmma &u)
{
    if (g_artTensorMode < 0) {
        const char *e = getenv ("ART_B200_UMMA");
        g_artTensorMode = e ? atoi (e) : 1;
    }
    if (g_artTensorDigits < 0) {
        const char *e = getenv ("ART_B200_DIGITS");
        g_artTensorDigits = e && atoi (e) == 2 ? 2 : 3;
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
    u.digits = g_artTensorDigits;
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
    // filter quantum: the exact accumulator holds sum X1*H1 with |X1| <= 2^11 and sum |H1| <= absSum * 2^DH + T/2
    u.DH = 0;
    for (int dh = 11; dh >= 6; --dh)
        if ((double) k.absSum * (double) (1 << dh) + 0.5 * k.T < 8191.0) { u.DH = dh; break; }
    if (!u.DH) return false;
    u.tableHalfs = u.numK * 3 * 2 * u.Npad * 8;

    // channels per tile: the 128 MMA rows are 128 / cg periods x cg channels (channel fastest), so that the converters read
    // whole frames (or 16-byte slices of the frames of a many-channel block) and the epilogue stores them; channel counts
    // that are not a multiple of 4 go through planar scratch from 8 channels on (art_device.cu).  Fewer channels per tile when
    // the row shifts no longer fit the operand.
    for (int cg = (k.C % 4 == 0 ? 4 : (k.C % 2 == 0 ? 2 : 1)); cg >= 1; cg >>= 1) {
        u.cg = cg;
        u.periods = ART_U_ROWS / cg + aMax;
        // rows: cg * periods, padded so that the 16-byte row slots a warp of converters writes into the two planes of a pair
        // fall on different banks (plane pitch = rows * 16 bytes)
        int rows = cg * u.periods;
        if (cg == 1) while ((rows & 1) || (rows & 7) < 2 || (rows & 7) > 6) ++rows;
        else if (cg == 2) while (!(rows & 1)) ++rows;
        else while ((rows & 7) == 0 || (rows & 7) == 4) ++rows;
        if (rows > 152) continue;
        u.rows = rows;
        // Shared memory is shared out between operand A (a ring of NS | KI plane-pair slots -- pairs are used in order, each
        // for all of its row shifts in a row, so a slot can take pair i + NS as soon as the MMAs of pair i are done), the
        // filter stage ring and the converters' raw-sample ring.  What the converters need most is loads in flight
        // (min (depth - 1, NS) pairs), then a filter ring of 6 k-steps, then operand slots.
        u.NS = 0;
        {
            long best = -1;
            int bNS = 0, bDepth = 0, bStages = 0;
            for (int depth = 4; depth >= 2; --depth)
                for (int ns = u.KI < ART_U_MAXKI ? u.KI : ART_U_MAXKI; ns >= 1; --ns) {
                    if (u.KI % ns) continue;
                    if (ns < 2 && u.KI > 1) continue;              // a ring slot of two k-steps may straddle two pairs: both must be resident
                    u.NS = ns; u.depth = depth;
                    for (u.stages = ART_U_STAGES; u.stages > 2 * ART_U_GROUP && umma_smem (u) > ART_U_SMEM_MAX; u.stages -= ART_U_GROUP) { }
                    if (umma_smem (u) > ART_U_SMEM_MAX) continue;
                    const int flying = depth - 1 < ns ? depth - 1 : ns;
                    const long score = 1000L * flying + 100L * (u.stages < 6 ? u.stages : 6) + ns;
                    if (score > best) { best = score; bNS = ns; bDepth = depth; bStages = u.stages; }
                }
            u.NS = bNS; u.depth = bDepth; u.stages = bStages;
        }
        if (!u.NS) continue;
        if (getenv ("ART_B200_TRACE"))
            fprintf (stderr, "[art] umma L=%d M=%d G=%d Npad=%d KI=%d NS=%d numK=%d cg=%d rows=%d DH=%d digits=%d stages=%d depth=%d smem=%zu\n",
                     u.L, u.M, u.G, u.Npad, u.KI, u.NS, u.numK, u.cg, u.rows, u.DH, u.digits, u.stages, u.depth, umma_smem (u));
        return true;
    }
    return false;
}

int artUmmaTiles (const ArtUmma &u, int channels, unsigned int outputs)
{
    const long long Q = ((long long) outputs + u.L - 1) / u.L;
    const int pr = ART_U_ROWS / u.cg;
    return (int) (((Q + pr - 1) / pr) * (channels / u.cg) * u.G);
}

static size_t umma_align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

size_t artUmmaTableBytes (const ArtUmma &u, int numTables, int numJobs, int totalTiles)
{
    return umma_align16 ((size_t) numTables * u.G * u.tableHalfs * sizeof (unsigned short)) + umma_align16 ((size_t) numJobs * u.G * sizeof (int)) +
           (size_t) totalTiles * sizeof (int);
}

void artUmmaCarve (ArtUmma &u, void *tables, int numTables, int numJobs)
{
    unsigned char *p = reinterpret_cast<unsigned char *> (tables);
    u.H = reinterpret_cast<unsigned short *> (p);
    p += umma_align16 ((size_t) numTables * u.G * u.tableHalfs * sizeof (unsigned short));
    u.S0 = reinterpret_cast<int *> (p);
    p += umma_align16 ((size_t) numJobs * u.G * sizeof (int));
    u.tileJob = reinterpret_cast<int *> (p);
}

template <int CGT, bool VEC>
static void umma_launch_one (const ArtClass &k, const ArtUmma &u, int totalTiles, int grid, const ArtJob &single, const ArtJob *d_jobs,
                             int roleProf, cudaStream_t stream)
{
    static bool configured[16] = { false };                 // per instantiation and device
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    if (!configured[device & 15]) {
        // the kernel moves registers between its warpgroups (setmaxnreg: 128 threads down to 56, 256 down to 88, 256 up to 120);
        // the pool that comes from is threads x registers-per-thread as compiled, so check that it suffices -- a warpgroup
        // asking for registers that never become free would spin forever
        cudaFuncAttributes fa;
        ART_CUDA_CHECK (cudaFuncGetAttributes (&fa, art_sinc_umma_kernel<CGT, VEC>));
        if (128 * (fa.numRegs - 56) + 256 * (fa.numRegs - 88) < 256 * (120 - fa.numRegs) || fa.numRegs < 88 || fa.numRegs > 120)
            artRaise ("art_sinc_umma_kernel was compiled with %d registers per thread: its register re-allocation plan does not hold", fa.numRegs);
        ART_CUDA_CHECK (cudaFuncSetAttribute (art_sinc_umma_kernel<CGT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ART_U_SMEM_MAX));
        configured[device & 15] = true;
    }
    void *prof;
    artProfileBegin (stream, &prof);
    art_sinc_umma_kernel<CGT, VEC><<<grid, ART_U_THREADS, umma_smem (u), stream>>> (k, u, single, d_jobs, totalTiles, roleProf);
    artProfileEnd (stream, prof);
    ART_CUDA_CHECK (cudaGetLastError ());
}

void artLaunchUmma (const ArtClass &k, const ArtUmma &u, int totalTiles, int numJobs, int numTables, int smCount,
                    const ArtJob &single, const ArtJob *d_jobs, bool vecIn, cudaStream_t stream)
{
    if (totalTiles <= 0) return;
    int histBlocks = (k.C * k.T + 127) / 128;
    if (histBlocks > 32) histBlocks = 32;
    const int prepBlocks = numTables * u.G * u.Npad + (numJobs * u.G + 127) / 128 + (totalTiles + 127) / 128 + numJobs * histBlocks;
    static int prepDbg = -1;
    if (prepDbg < 0) {
        prepDbg = 0;
#ifdef ART_B200_ABLATE
        if (const char *d = getenv ("ART_B200_UDBG")) prepDbg = atoi (d);
#endif
    }
    const ArtUmma &uu = u;
    art_umma_prep_kernel<<<prepBlocks, 128, 0, stream>>> (k, uu, single, d_jobs, numJobs, numTables, histBlocks, totalTiles, prepDbg);
    ART_CUDA_CHECK (cudaGetLastError ());
    const int grid = totalTiles < smCount ? totalTiles : smCount;
    static int roleProf = -1;
    if (roleProf < 0) {
        roleProf = 0;
#ifdef ART_B200_ABLATE
        roleProf = getenv ("ART_B200_UPROF") ? 1 : 0;
        if (const char *d = getenv ("ART_B200_UDBG")) roleProf |= atoi (d) << 4;
#endif
        if (roleProf & 1) atexit ([] () {
            unsigned long long h[24];
            cudaDeviceSynchronize ();
            if (cudaMemcpyFromSymbol (h, g_uprof, sizeof h) != cudaSuccess) return;
            const double n = h[10] ? (double) h[10] : 1.0;
            unsigned long long tl[160][4];
            if (cudaMemcpyFromSymbol (tl, g_utime, sizeof tl) == cudaSuccess) {
                unsigned long long t0 = ~0ull, t3 = 0;
                for (int i = 0; i < 148; ++i) { if (tl[i][0] && tl[i][0] < t0) t0 = tl[i][0]; if (tl[i][3] > t3) t3 = tl[i][3]; }
                double a1 = 0, a2 = 0, a3 = 0, m1 = 0, m2 = 0, m3 = 0, e0 = 0; int nn = 0;
                double lo3 = 1e30;
                for (int i = 0; i < 148; ++i) if (tl[i][0]) {
                    const double s = (double) (tl[i][0] - t0), f = (double) (tl[i][1] - t0), l = (double) (tl[i][2] - t0), x = (double) (tl[i][3] - t0);
                    e0 += s; a1 += f; a2 += l; a3 += x; if (f > m1) m1 = f; if (l > m2) m2 = l; if (x > m3) m3 = x; if (x < lo3) lo3 = x; ++nn;
                }
                if (nn) fprintf (stderr, "[art] umma timeline of the last launch (us from first CTA entry; avg / max over %d CTAs): entry %.1f | first MMA %.1f / %.1f | last MMA commit %.1f / %.1f | exit %.1f / %.1f (min %.1f)\n",
                                 nn, e0 / nn / 1e3, a1 / nn / 1e3, m1 / 1e3, a2 / nn / 1e3, m2 / 1e3, a3 / nn / 1e3, m3 / 1e3, lo3 / 1e3);
            }
            fprintf (stderr, "[art] umma cycles per tile: producer wait-empty %.0f | mma wait-planes %.0f wait-acc %.0f wait-h %.0f tile %.0f | "
                     "convert wait-planes %.0f | epilogue wait %.0f drain %.0f | tiles %.0f | mma issue %.0f commit %.0f region %.0f | epilogue store %.0f scan %.0f | convert split %.0f fence %.0f arrive %.0f quantum-wait %.0f loads+head %.0f\n",
                     h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[8] / n, h[9] / n, n, h[11] / n, h[12] / n, h[13] / n,
                     h[6] / n, h[16] / n, h[7] / n, h[14] / n, h[15] / n, h[17] / n, h[18] / n);
        });
    }
    // vecIn: every job reads an interleaved block whose frames hold the tile's channels adjacent and aligned, so the converters
    // copy whole frames (one 8- or 16-byte cp.async per tap); a compile-time property of the kernel, not a per-lane branch
    // tile order (see the kernel): contiguous ranges per CTA when the jobs are short -- a good part of their tiles then starts at a
    // history or ends at the end of the input, and round-robin can hand all of those to the same CTAs; round-robin otherwise (it
    // measured ~1.5 % faster on long jobs)
    const int orderBit = totalTiles / (numJobs > 0 ? numJobs : 1) < 32 ? 4 : 0;
    const int launchArg = roleProf | orderBit;
    if (u.cg == 4)      { if (vecIn) umma_launch_one<4, true> (k, uu, totalTiles, grid, single, d_jobs, launchArg, stream);
                          else       umma_launch_one<4, false> (k, uu, totalTiles, grid, single, d_jobs, launchArg, stream); }
    else if (u.cg == 2) { if (vecIn) umma_launch_one<2, true> (k, uu, totalTiles, grid, single, d_jobs, launchArg, stream);
                          else       umma_launch_one<2, false> (k, uu, totalTiles, grid, single, d_jobs, launchArg, stream); }
    else                umma_launch_one<1, false> (k, uu, totalTiles, grid, single, d_jobs, launchArg, stream);
    g_artLaunches += 2;
}

#else   /* ART_WIDE */

int g_artTensorDigits = -1;
bool artPlanUmma (const ArtClass &, double, unsigned int, unsigned long long, int, ArtUmma &) { return false; }
int artUmmaTiles (const ArtUmma &, int, unsigned int) { return 0; }
size_t artUmmaTableBytes (const ArtUmma &, int, int, int) { return 0; }
void artUmmaCarve (ArtUmma &, void *, int, int) { }
void artLaunchUmma (const ArtClass &, const ArtUmma &, int, int, int, int, const ArtJob &, const ArtJob *, bool, cudaStream_t) { }

#endif
