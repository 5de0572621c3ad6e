/*
 * art_sinc_periodic.cu -- the rational-ratio windowed-sinc kernel (sm_100a).
 *
 * Same reference functions as art_sinc_generic.cu (resampler.c:523-526, :640-643, :1135-1157,
 * :1033-1044), restructured for ratios L/M with a small numerator (44.1k->48k = 160/147,
 * 48k->44.1k = 147/160, 96k->44.1k = 147/320, ... -- every BASELINE config but the ASRC one).
 *
 * With ratio = L/M the read position advances by exactly M input samples every L outputs, so
 * output n = L*q + j ("phase j of period q") applies the SAME two filter rows with the SAME
 * interpolation weight for every q:
 *
 *      y[j, q] = sum_k h_j[k] * x[s_j + M*q + k],   h_j = (1 - f_j) * row(fi_j) + f_j * row(fi_j + 1)
 *
 * i.e. a (phases x taps) by (taps x periods) product whose left operand is a banded matrix of L
 * pre-interpolated filters and whose right operand is the input read at stride M.  Both operands
 * are now reused (each h_j by every period and channel, each input sample by every phase), which
 * the per-output dot product of the generic kernel cannot offer, and the interpolation costs T
 * multiply-adds per phase per call instead of T per output sample: T FMAs per sample, not 2T.
 *
 *   1. art_phase_table_kernel: for the L phases of a segment, evaluate the reference's position
 *      arithmetic exactly (art_plan.h) at the segment's first period, then h_j in double -> float.
 *      Later periods reuse (fi_j, f_j, s_j + M*q): the exact positions drift from that by
 *      < 1e-9 sample over a segment (the host caps segment length accordingly), far below the
 *      1e-6 * peak tolerance; the planner still owns input_used/output_generated/position exactly.
 *   2. art_sinc_periodic_kernel: a CTA owns up to 32 phases (4 rows of 8) and a run of periods.
 *      The phases' filters sit in shared memory, shifted to a common origin ("union window" of
 *      Kp taps); the input is staged chunk by chunk.  A warp computes an 8-phase x 8-column tile
 *      (columns = periods x channels): lanes split the taps, 64 accumulators per lane, 16 shared
 *      loads per 64 FMAs, then two transposing shuffle reductions and coalesced-by-phase stores.
 *
 * Non-interpolated contexts (resampleFixedRatioInit's reduced bank, resampler.c:323-335) use the
 * same kernel: h_j is the bank row itself, or a unit impulse where the reference returns the
 * stored sample verbatim (resampler.c:1141-1142) -- 1.0 * x + 0 * rest is exact.
 */
#include <cstdio>
#include <cstdlib>
#include "art_kernels.cuh"
#include "art_device.h"

#if !ART_WIDE       /* a float-path optimisation: the wide (PATH_WIDTH=64) build keeps to the any-ratio kernel */

#define ART_P_THREADS 256
#define ART_P_WARPS   (ART_P_THREADS / 32)
#define ART_P_ROWS_MAX 4

template <int CV> struct ArtPVec;
template <> struct ArtPVec<1> { typedef float  type; __device__ static float get (const float  &x, int)   { return x; } };
template <> struct ArtPVec<2> { typedef float2 type; __device__ static float get (const float2 &x, int v) { return v ? x.y : x.x; } };
template <> struct ArtPVec<4> { typedef float4 type; __device__ static float get (const float4 &x, int v) { return v == 0 ? x.x : v == 1 ? x.y : v == 2 ? x.z : x.w; } };

/* ---- TMA bulk copy + mbarrier (sm_90+/sm_100: cp.async.bulk -> SASS UBLKCP) ---------------------- */
__device__ __forceinline__ unsigned int art_smem_u32 (const void *p)
{
    return (unsigned int) __cvta_generic_to_shared (p);
}
__device__ __forceinline__ void art_mbar_init (unsigned long long *bar, unsigned int count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(art_smem_u32 (bar)), "r"(count));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void art_mbar_expect_tx (unsigned long long *bar, unsigned int bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(art_smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void art_mbar_wait (unsigned long long *bar, unsigned int parity)
{
    // try_wait suspends the thread for a bounded time itself; the spin limit only turns a lost copy
    // (a bug) into a trap instead of a hung GPU
    for (unsigned int spins = 0; spins < (1u << 26); ++spins) {
        unsigned int done;
        asm volatile (
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(art_smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    printf ("libresampler_b200: bulk copy never completed (block %d)\n", blockIdx.x);
    __trap ();
}
/* bytes must be a multiple of 16, both addresses 16-byte aligned */
__device__ __forceinline__ void art_bulk_g2s (void *dstSmem, const void *srcGlobal, unsigned int bytes, unsigned long long *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(art_smem_u32 (dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(art_smem_u32 (bar)) : "memory");
}

/* ---- 1. the phase table ----------------------------------------------------------------------- */
/* One block per (padded) phase.  Writes the phase's interpolated filter straight into the layout the
 * product kernel keeps in shared memory: [row][step][half-row][lane][4 phases], shifted so that tap 0 of
 * the phase block's first phase sits at m = 0. */
__device__ __forceinline__ void art_phase_table_block (const ArtClass &k, const ArtPeriodic &p, const ArtJob &single,
                                                        const ArtJob *__restrict__ jobs, int j, int tbl)
{
    const ArtJob &job = jobs ? jobs[jobs[tbl].repJob] : single;
    const int T = k.T, Tref = k.Tref, half = Tref / 2 + k.lead, F = k.F;          // T taps per row; positions run on Tref
    const int perBlock = p.rowsPerCta * 8;
    const int pb = j / perBlock, jj = j - pb * perBlock, jb = pb * perBlock;
    __shared__ int sh_row, sh_pass, sh_shift;
    __shared__ double sh_f;

    if (threadIdx.x == 0) {
        ArtLoopState st;
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = Tref;
        long long sj = 0, sb = 0;
        int row = 0, pass = -1;
        double f = 0.0;
        for (int which = 0; which < 2; ++which) {               // 0: the block's first phase, 1: this phase
            const int ph = which ? j : jb;
            if (ph >= p.L) break;
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
                    pass = half - 1 + (row ? 1 : 0);
            }
        }
        sh_row = row; sh_f = f; sh_pass = pass; sh_shift = (int) (sj - sb);
    }
    __syncthreads ();
    const int row = sh_row, pass = sh_pass, shift = sh_shift;
    const double f = sh_f;
    const int NIg = p.Kp >> 5;
    float *dst = p.Hblk + ((size_t) tbl * p.PB + pb) * perBlock * p.Kp;
    const float *ra = k.bank + (size_t) row * k.Tp, *rb = ra + k.Tp;
    for (int m = threadIdx.x; m < p.Kp; m += 128) {
        const int t = m - shift;
        float h = 0.0f;
        if (j < p.L && t >= 0 && t < T) {
            if (pass >= 0)
                h = t == pass ? 1.0f : 0.0f;
            else if (k.mode & ART_MODE_INTERP) {
                const double a = ra[t], b = rb[t];
                h = (float) (a + f * (b - a));
            }
            else
                h = ra[t];
        }
        // [row][step][phases 0-3 | 4-7][lane][4]: a lane fetches the 8 phases of its tap with two LDS.128
        dst[((((jj >> 3) * NIg + (m >> 5)) * 2 + ((jj >> 2) & 1)) * 32 + (m & 31)) * 4 + (jj & 3)] = h;
    }
}

/* Everything the product kernel needs before it can start, in ONE launch (a call's GPU time is
 * dominated by launch latency for small blocks): phase tables, per-job block origins, and -- because it
 * only reads what the product kernel also only reads -- the history update of resampler.c's ring. */
__global__ void __launch_bounds__ (128)
art_periodic_prep_kernel (const ArtClass k, const ArtPeriodic p, const __grid_constant__ ArtJob single,
                          const ArtJob *__restrict__ jobs, int numJobs, int numTables, int histBlocksPerJob)
{
    const int padded = p.PB * p.rowsPerCta * 8;
    const int tableBlocks = numTables * padded;
    const int originBlocks = (numJobs * p.PB + 127) / 128;
    int b = blockIdx.x;
    if (b < tableBlocks) {
        art_phase_table_block (k, p, single, jobs, b % padded, b / padded);
        return;
    }
    b -= tableBlocks;
    if (b < originBlocks) {
        const int e = b * 128 + threadIdx.x;
        if (e >= numJobs * p.PB) return;
        const int seg = e / p.PB, pb = e - seg * p.PB;
        const ArtJob &job = jobs ? jobs[seg] : single;
        ArtLoopState st;
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = k.Tref;
        int w;
        const double pos = art_output_pos (&st, job.nStart + pb * p.rowsPerCta * 8, &w);
        p.S0[e] = (int) ((long long) floor (pos) - (k.Tref / 2 + k.lead) + 1 + (long long) w * 15LL * k.Tref - job.origin);
        return;
    }
    b -= originBlocks;
    {
        const int seg = b / histBlocksPerJob, hb = b - seg * histBlocksPerJob;
        const ArtJob &job = jobs ? jobs[seg] : single;
        if (!job.histOut) return;
        const int total = k.C * k.T;
        for (int e = hb * 128 + threadIdx.x; e < total; e += histBlocksPerJob * 128) {
            const int c = e / k.T, i = e - c * k.T;
            job.histOut[e] = art_fetch (job, k.T, c, job.consumed - k.T + i);
        }
    }
}

/* ---- 2. the banded product --------------------------------------------------------------------- */
/* THREADS = 256 with two CTAs per SM (shared memory <= 110 KB), or 512 with one large CTA per SM */
template <int CV, int THREADS>
__global__ void __launch_bounds__ (THREADS, THREADS == 256 ? 2 : 1)
art_sinc_periodic_kernel (const ArtClass k, const ArtPeriodic p, const __grid_constant__ ArtJob single,
                          const ArtJob *__restrict__ jobs)
{
    typedef typename ArtPVec<CV>::type VecT;
    constexpr int QT = 8 / CV;                                  // periods per warp tile

    extern __shared__ __align__ (128) unsigned char smem_raw[];
    float *Hs = reinterpret_cast<float *> (smem_raw);            // [rows][NIg][2][32][4]
    float *xsBuf = Hs + (size_t) p.rowsPerCta * 8 * p.Kp;        // [Wc + 4][CV] (+4: alignment slack of the bulk copy)
    __shared__ __align__ (8) unsigned long long bars[2];         // [0] filters, [1] input chunk

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // channel group varies fastest across the grid: the CTAs that read the same frames of an interleaved
    // block run together, so each 128-byte line comes from DRAM once and from L2 for the other groups
    const int groups = (k.C + CV - 1) / CV;
    const int cta = blockIdx.x / groups, cgroup = blockIdx.x - cta * groups;
    const int seg = jobs ? (k.numJobs > 1 ? art_find_job_warp (jobs, k.numJobs, cta) : 0) : 0;
    const ArtJob &job = jobs ? jobs[seg] : single;
    const int L = p.L, M = p.M, T = k.T;
    const int Q = (int) ((job.outputs + L - 1) / L);             // periods in this segment
    const int R = (L + 7) >> 3;                                   // phase rows
    const int local = cta - job.tile0;
    const int pb = local % p.PB, qb = local / p.PB;               // phase block fastest: neighbours share input in L2
    const int row0 = pb * p.rowsPerCta;
    const int nrows = min (p.rowsPerCta, R - row0);
    const int j0 = row0 * 8;
    const int qStart = qb * p.Qblk, qEnd = min (Q, qStart + p.Qblk);
    if (qStart >= qEnd)
        return;
    const int c0 = cgroup * CV;
    const int NIg = p.Kp >> 5;
    const long long S0 = p.S0[(size_t) seg * p.PB + pb];

    if (tid == 0) {
        art_mbar_init (&bars[0], 1);
        art_mbar_init (&bars[1], 1);
    }
    __syncthreads ();
    /* the CTA's filters: one bulk copy of the block the table kernel laid out */
    if (tid == 0) {
        const unsigned int bytes = (unsigned int) (p.rowsPerCta * 8 * p.Kp * sizeof (float));
        art_mbar_expect_tx (&bars[0], bytes);
        art_bulk_g2s (Hs, p.Hblk + ((size_t) job.table * p.PB + pb) * p.rowsPerCta * 8 * p.Kp, bytes, &bars[0]);
    }

    // the bulk path needs the chunk to be one contiguous, fully valid span of the caller's interleaved block
    const bool bulkOk = (job.inPlanes == nullptr) && job.inCS == 1 && job.inFS == CV && k.C == CV;
    unsigned int xPhase = 0;
    bool filtersReady = false;

    /* Stage chunk `q0` into buffer `buf`: asynchronously by TMA when it is a plain span of the caller's block
     * (returns the float offset of sample 0 inside the buffer, >= 0), otherwise cooperatively by all threads
     * (returns -1 - offset; the caller must __syncthreads before reading). */
    auto stage = [&] (int q0) -> int {
        const int nq = min (p.Qc, qEnd - q0);
        const long long a = S0 + (long long) M * q0;             // region index of the chunk's first sample
        const int len = (nq - 1) * M + p.Kp;                      // samples per channel
        float *xsRaw = xsBuf;
        const float *g = job.in + a * CV;
        if (bulkOk && a >= -job.prevAvail && a + len + 4 <= (long long) job.inValid) {
            // align the source down to 16 bytes; the same slack appears in front of the data in shared memory
            const int xoff = (int) ((reinterpret_cast<unsigned long long> (g) >> 2) & 3);
            if (tid == 0) {
                const unsigned int bytes = (unsigned int) (((size_t) len * CV + xoff + 3) & ~(size_t) 3) * sizeof (float);
                art_mbar_expect_tx (&bars[1], bytes);
                art_bulk_g2s (xsRaw, g - xoff, bytes, &bars[1]);
            }
            return xoff;
        }
        if (CV > 1 && job.inPlanes == nullptr && job.inCS == 1 && c0 + CV <= k.C &&
            a >= -job.prevAvail && a + len <= (long long) job.inValid &&
            ((reinterpret_cast<unsigned long long> (job.in + a * job.inFS + c0) & (sizeof (VecT) - 1)) == 0) &&
            ((job.inFS * sizeof (float)) & (sizeof (VecT) - 1)) == 0) {
            // a CV-channel group of a wider interleaved block: one vector load per frame, four in flight
            const float *g0 = job.in + a * job.inFS + c0;
            VecT *dst = reinterpret_cast<VecT *> (xsRaw);
            int i = tid;
            for (; i + 3 * THREADS < len; i += 4 * THREADS) {
                VecT v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = __ldg (reinterpret_cast<const VecT *> (g0 + (long long) (i + u * THREADS) * job.inFS));
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dst[i + u * THREADS] = v[u];
            }
            for (; i < len; i += THREADS)
                dst[i] = __ldg (reinterpret_cast<const VecT *> (g0 + (long long) i * job.inFS));
            return -1;
        }
        const int total = len * CV;
        for (int e = tid; e < total; e += THREADS) {
            const int i = e / CV, v = e - i * CV;
            xsRaw[e] = (c0 + v < k.C) ? art_fetch (job, T, c0 + v, a + i) : 0.0f;
        }
        return -1;
    };

    /* (a second buffer with the next chunk's copy in flight was measured: no gain -- the second resident
     * CTA already covers the copy -- so the shared memory goes to longer chunks instead) */
    for (int q0 = qStart; q0 < qEnd; q0 += p.Qc) {
        const int nq = min (p.Qc, qEnd - q0);
        __syncthreads ();                                          // previous chunk fully consumed
        const int mine = stage (q0);
        if (!filtersReady) {
            art_mbar_wait (&bars[0], 0);
            filtersReady = true;
        }
        int xoff = 0;
        if (mine >= 0) {
            art_mbar_wait (&bars[1], xPhase);
            xPhase ^= 1;
            xoff = mine;
        }
        else
            __syncthreads ();                                      // cooperative stores of this chunk are complete
        const float *xs = xsBuf + xoff;

        const int qTiles = (nq + QT - 1) / QT;
        for (int tile = warp; tile < nrows * qTiles; tile += (THREADS / 32)) {
            const int row = tile % nrows, qloc = (tile / nrows) * QT;

            /* 64 accumulators as 32 packed pairs: acc2[pp][col] = phases (2pp, 2pp+1) of the row x column col
             * (col = period * CV + channel), in SLOT coordinates: a lane may hold its periods and its two
             * half-rows in a lane-dependent order, because that order is only a matter of which pointer it
             * loads from.  The lane bits chosen below make the three big stages of the transposing reduction
             * (56 of 62 shuffles) need no select at all: every lane sends its upper slots and keeps its lower
             * ones, and the XOR-permuted slot order guarantees partner lanes exchange matching values. */
            constexpr int QB = CV == 1 ? 3 : (CV == 2 ? 2 : 1);         // period bits of a tile
            // lane bits (from bit 4 down) are spent on: period bits, then the half-row bit, then the rest
            const int qmask = (lane >> (5 - QB)) & (QT - 1);            // XOR applied to the period slot index
            const int hsel = (lane >> (4 - QB)) & 1;                    // 1: this lane's slots 0,1 hold phases 4..7

            unsigned long long acc2[4][8];
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
#pragma unroll
                for (int b2 = 0; b2 < 8; ++b2) acc2[a2][b2] = 0ull;

            const float4 *hpA = reinterpret_cast<const float4 *> (Hs) + (size_t) row * NIg * 64 + lane + hsel * 32;
            const float4 *hpB = reinterpret_cast<const float4 *> (Hs) + (size_t) row * NIg * 64 + lane + (1 - hsel) * 32;
            const VecT *xp[QT];
#pragma unroll
            for (int qq = 0; qq < QT; ++qq)
                xp[qq] = reinterpret_cast<const VecT *> (xs) + (size_t) min (qloc + (qq ^ qmask), p.Qc - 1) * M + lane;

#pragma unroll 2
            for (int i = 0; i < NIg; ++i) {
                const float4 ha = hpA[i * 64], hb = hpB[i * 64];
                unsigned long long h2[4];
                h2[0] = art_pack2 (ha.x, ha.y); h2[1] = art_pack2 (ha.z, ha.w);
                h2[2] = art_pack2 (hb.x, hb.y); h2[3] = art_pack2 (hb.z, hb.w);
#pragma unroll
                for (int qq = 0; qq < QT; ++qq) {
                    const VecT xv = xp[qq][32 * i];
#pragma unroll
                    for (int v = 0; v < CV; ++v) {
                        const float x = ArtPVec<CV>::get (xv, v);
                        const unsigned long long x2 = art_pack2 (x, x);         // folded into a scalar operand
#pragma unroll
                        for (int pp = 0; pp < 4; ++pp)
                            art_ffma2 (acc2[pp][qq * CV + v], h2[pp], x2);
                    }
                }
            }

            /* flat value index: bit 0 = low/high phase of a pair, bits 1-3 = column slot, bits 4-5 = pair slot */
            float v[64];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    art_unpack2 (acc2[pp][c], v[(pp << 4) | (c << 1)], v[(pp << 4) | (c << 1) | 1]);

            // stage order: which flat bit each shuffle offset (16, 8, 4, 2, 1) folds away, and whether that
            // bit was XOR-permuted per lane (select-free) or not
            constexpr int chBits = CV == 1 ? 0 : (CV == 2 ? 1 : 2);
            int consumed = 0;
#pragma unroll
            for (int stage = 0; stage < 5; ++stage) {
                const int off = 16 >> stage;
                int bit;                                        // flat bit folded in this stage
                bool isFree;
                if (stage < QB)            { bit = 3 - stage;                 isFree = true; }     // period bits, high to low
                else if (stage == QB)      { bit = 5;                         isFree = true; }     // half-row (pair bit 1)
                else if (stage < QB + 1 + chBits) { bit = chBits - (stage - QB - 1);  isFree = false; }   // channel bits
                else                       { bit = 4;                         isFree = false; }    // pair bit 0
                const int bmask = 1 << bit;
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    if ((i & consumed) != 0 || (i & bmask) != 0) continue;
                    if (isFree)
                        v[i] += __shfl_xor_sync (0xffffffffu, v[i | bmask], off);
                    else {
                        const float send = upper ? v[i] : v[i | bmask];
                        const float keep = upper ? v[i | bmask] : v[i];
                        v[i] = keep + __shfl_xor_sync (0xffffffffu, send, off);
                    }
                }
                consumed |= bmask;
            }

            /* every lane now owns one phase pair of one column: v[0], v[1] */
            {
                int lb = 4;                                     // walk the lane bits in the order they were spent
                int qa = 0;
#pragma unroll
                for (int t = 0; t < QB; ++t) qa = (qa << 1) | ((lane >> lb--) & 1);
                const int pp1 = (lane >> lb--) & 1;
                int ch = 0;
#pragma unroll
                for (int t = 0; t < chBits; ++t) ch = (ch << 1) | ((lane >> lb--) & 1);
                const int pp0 = lane & 1;
                const int j = j0 + row * 8 + ((pp1 << 1) | pp0) * 2;
                const long long nl = (long long) (q0 + qloc + qa) * L + j;          // output index inside the segment
                if (qloc + qa < nq && c0 + ch < k.C) {
                    if (j < L && nl < (long long) job.outputs)
                        *art_out_ptr (job, c0 + ch, (long long) job.nStart + nl) = v[0];
                    if (j + 1 < L && nl + 1 < (long long) job.outputs)
                        *art_out_ptr (job, c0 + ch, (long long) job.nStart + nl + 1) = v[1];
                }
            }
        }
    }
}

/* ---- host side ----------------------------------------------------------------------------------- */

/* Smallest-denominator fraction L/M equal to `ratio` within double rounding, L <= maxL. */
bool artRational (double ratio, int maxL, int *Lout, int *Mout)
{
    if (!(ratio > 1e-6) || !(ratio < 1e6))
        return false;
    double x = ratio;
    long long p0 = 0, q0 = 1, p1 = 1, q1 = 0;                    // convergents p/q of ratio
    for (int it = 0; it < 64; ++it) {
        const double a = floor (x);
        if (a > 1e9) break;
        const long long p2 = (long long) a * p1 + p0, q2 = (long long) a * q1 + q0;
        if (p2 > maxL || q2 > (1LL << 30)) break;
        p0 = p1; q0 = q1; p1 = p2; q1 = q2;
        if (q1 > 0) {
            const double err = fabs ((double) p1 / (double) q1 - ratio);
            if (err <= 4.0e-16 * ratio) {
                // the period must close on the input axis too: L / ratio == M to rounding
                if (fabs ((double) p1 / ratio - (double) q1) <= 1e-12 * (double) q1) {
                    *Lout = (int) p1; *Mout = (int) q1;
                    return true;
                }
                return false;
            }
        }
        const double fr = x - a;
        if (fr < 1e-15) break;
        x = 1.0 / fr;
    }
    return false;
}

static size_t periodic_smem (const ArtPeriodic &p, int CV)
{
    return ((size_t) p.rowsPerCta * 8 * p.Kp + (size_t) (p.Wc + 4) * CV) * sizeof (float) + 128;
}

/* Tile geometry for one launch; returns false when the periodic form does not pay or fit. */
bool artPlanPeriodic (const ArtClass &k, double ratio, unsigned int maxOutputs, unsigned long long totalOutputs,
                      int smCount, ArtPeriodic &p, int &CV)
{
    int L, M;
    if (k.mode & ART_MODE_PRECISE) return false;                 // double accumulation: generic kernel
    if (!artRational (ratio, 1024, &L, &M)) return false;
    if (L < 5) return false;                                      // an 8-phase tile would idle: generic kernel
    if (maxOutputs < (unsigned) (4 * L)) return false;
    CV = k.C >= 4 ? 4 : (k.C >= 2 ? 2 : 1);
    const int QT = 8 / CV;

    p.L = L; p.M = M;
    const int R = (L + 7) / 8;
    int bestRows = 0, bestQc = 0;
    double bestScore = 0.0;
    // two shared-memory budgets: 110 KB keeps two CTAs (16 warps) per SM; long filters or large decimation
    // factors need more room per CTA to have at least a tile per warp, and then run one CTA per SM
    for (int pass = 0; pass < 2; ++pass) {
        const size_t budget = pass ? 200 * 1024 : 110 * 1024;
        for (int rows = ART_P_ROWS_MAX; rows >= 1; --rows) {
            // shifts inside a phase block: consecutive phases advance by M/L samples
            const int spread = (int) (((long long) (rows * 8 - 1) * M + L - 1) / L) + 2;
            const int Kp = (k.T + spread + 31) & ~31;
            for (int Qc = 64; Qc >= QT; Qc -= QT) {
                ArtPeriodic t = p;
                t.rowsPerCta = rows; t.Kp = Kp; t.Qc = Qc; t.Wc = (Qc - 1) * M + Kp;
                if (periodic_smem (t, CV) > budget) continue;
                // staged input is reused by rows*8 phases, the filters by Qc periods: weigh both, prefer
                // enough tiles per chunk to keep 8 warps busy
                const int tiles = rows * (Qc / QT);
                const int warps = pass ? 16 : 8;
                double score = (double) rows * 8 * Qc / (rows * 8 + Qc) * (tiles >= warps ? 1.0 : (double) tiles / warps);
                score *= (double) k.T / Kp;                       // union-window padding is wasted FMAs
                const int usedRows = (R + rows - 1) / rows * rows;
                score *= (double) R / usedRows;                   // idle rows in the last phase block
                if (pass) score *= 0.8;                           // one CTA per SM: no second CTA to hide the staging
                if (score > bestScore) { bestScore = score; bestRows = rows; bestQc = Qc; }
                break;                                            // largest Qc that fits for this row count
            }
        }
    }
    if (!bestRows) return false;
    p.rowsPerCta = bestRows;
    {
        const int spread = (int) (((long long) (bestRows * 8 - 1) * M + L - 1) / L) + 2;
        p.Kp = (k.T + spread + 31) & ~31;
    }
    p.Qc = bestQc;
    p.Wc = (p.Qc - 1) * M + p.Kp;
    // periods per CTA: several chunks (the filter block is fetched once per CTA), but keep the whole
    // launch a few waves deep
    p.PB = (R + p.rowsPerCta - 1) / p.rowsPerCta;
    const long long periods = (long long) ((totalOutputs + L - 1) / L);
    const int groups = (k.C + CV - 1) / CV;
    // candidates: 1..16 chunks per CTA; score = how full the last wave is (2 CTAs per SM) x how well
    // the once-per-CTA filter fetch is amortised
    const long long perJob = (maxOutputs + L - 1) / L;
    const long long jobsApprox = perJob ? (periods + perJob - 1) / perJob : 1;
    const long long slots = (long long) smCount * (periodic_smem (p, CV) > 110 * 1024 ? 1 : 2);
    int bestChunks = 1;
    double best = -1.0;
    for (int chunks = 1; chunks <= 16; ++chunks) {
        const long long perCta = (long long) p.Qc * chunks;
        const long long ctas = jobsApprox * p.PB * groups * ((perJob + perCta - 1) / perCta);
        const long long waves = (ctas + slots - 1) / slots;
        const double fill = (double) ctas / (double) (waves * slots);
        const double amort = (double) chunks / (chunks + 0.35);
        const double sc = fill * amort;
        if (sc > best) { best = sc; bestChunks = chunks; }
    }
    p.Qblk = p.Qc * bestChunks;

    // experiment knobs (measurement builds only): ART_P_ROWS / ART_P_QC / ART_P_CHUNKS override the choice
#ifdef ART_B200_ABLATE
    if (const char *e = getenv ("ART_P_ROWS")) {
        p.rowsPerCta = atoi (e);
        const int spread = (int) (((long long) (p.rowsPerCta * 8 - 1) * M + L - 1) / L) + 2;
        p.Kp = (k.T + spread + 31) & ~31;
        p.PB = (R + p.rowsPerCta - 1) / p.rowsPerCta;
    }
    if (const char *e = getenv ("ART_P_QC")) p.Qc = atoi (e);
#endif
    p.Wc = (p.Qc - 1) * M + p.Kp;
    p.Qblk = p.Qc * bestChunks;
#ifdef ART_B200_ABLATE
    if (const char *e = getenv ("ART_P_CHUNKS")) p.Qblk = p.Qc * atoi (e);
#endif
    if (periodic_smem (p, CV) > 200 * 1024) return false;
    if (getenv ("ART_B200_TRACE"))
        fprintf (stderr, "[art] periodic L=%d M=%d rows=%d Kp=%d Qc=%d Qblk=%d PB=%d CV=%d smem=%zu\n",
                 p.L, p.M, p.rowsPerCta, p.Kp, p.Qc, p.Qblk, p.PB, CV, periodic_smem (p, CV));
    return true;
}

/* longest segment whose drift from exact periodicity stays below ~1e-9 sample */
unsigned int artPeriodicSegmentOutputs (const ArtPeriodic &p, double ratio)
{
    const double perPeriod = fabs ((double) p.L / ratio - (double) p.M) + 4.0e-16 * p.M;
    double periods = 1.0e-9 / perPeriod;
    if (periods > 1.0e6) periods = 1.0e6;
    if (periods < 64.0) periods = 64.0;
    return (unsigned int) periods * (unsigned int) p.L;
}

int artPeriodicCtas (const ArtPeriodic &p, unsigned int outputs)
{
    const long long Q = ((long long) outputs + p.L - 1) / p.L;
    return (int) (p.PB * ((Q + p.Qblk - 1) / p.Qblk));
}

template <int CV, int THREADS>
static void launch_periodic (const ArtClass &k, const ArtPeriodic &p, int totalCtas, int numJobs, int numTables,
                             const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    auto kern = art_sinc_periodic_kernel<CV, THREADS>;
    static size_t configured[16] = { 0 };
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    const size_t smem = periodic_smem (p, CV);
    if (smem > configured[device & 15]) {
        const size_t want = smem > 112 * 1024 ? 200 * 1024 : 112 * 1024;
        ART_CUDA_CHECK (cudaFuncSetAttribute (kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) want));
        configured[device & 15] = want;
    }
    int histBlocks = (k.C * k.T + 127) / 128;
    if (histBlocks > 32) histBlocks = 32;
    const int prepBlocks = numTables * p.PB * p.rowsPerCta * 8 + (numJobs * p.PB + 127) / 128 + numJobs * histBlocks;
    art_periodic_prep_kernel<<<prepBlocks, 128, 0, stream>>> (k, p, single, d_jobs, numJobs, numTables, histBlocks);
    ART_CUDA_CHECK (cudaGetLastError ());
    const unsigned int grid = (unsigned int) totalCtas * (unsigned int) ((k.C + CV - 1) / CV);
    void *prof;
    artProfileBegin (stream, &prof);
    kern<<<grid, THREADS, smem, stream>>> (k, p, single, d_jobs);
    artProfileEnd (stream, prof);
    ART_CUDA_CHECK (cudaGetLastError ());
    g_artLaunches += 2;
}

void artLaunchPeriodic (const ArtClass &k, const ArtPeriodic &p, int CV, int totalCtas, int numJobs, int numTables,
                        const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    if (totalCtas <= 0) return;
    const bool big = periodic_smem (p, CV) > 110 * 1024;         // one CTA per SM: give it 16 warps
#define ART_LP(CVV)                                                                                    \
    do {                                                                                               \
        if (big) launch_periodic<CVV, 512> (k, p, totalCtas, numJobs, numTables, single, d_jobs, stream); \
        else     launch_periodic<CVV, 256> (k, p, totalCtas, numJobs, numTables, single, d_jobs, stream); \
    } while (0)
    if (CV == 4) ART_LP (4);
    else if (CV == 2) ART_LP (2);
    else ART_LP (1);
#undef ART_LP
}

#else   /* ART_WIDE */

bool artRational (double, int, int *, int *) { return false; }
bool artPlanPeriodic (const ArtClass &, double, unsigned int, unsigned long long, int, ArtPeriodic &, int &) { return false; }
unsigned int artPeriodicSegmentOutputs (const ArtPeriodic &, double) { return 0; }
int artPeriodicCtas (const ArtPeriodic &, unsigned int) { return 0; }
void artLaunchPeriodic (const ArtClass &, const ArtPeriodic &, int, int, int, int, const ArtJob &, const ArtJob *, cudaStream_t) { }

#endif
