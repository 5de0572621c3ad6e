/*
 * art_sinc_periodic2.cu -- rational-ratio sinc kernel, register-tiled GEMM form (sm_100a).
 *
 * Same mathematics as art_sinc_periodic.cu (see there for the derivation and the reference lines):
 *
 *      y[j, q, c] = sum_m Hfull[j][m] * x[(S0 + M*q + m) * C + c]
 *
 * with Hfull[j][m] = h_j[m - d_j] the pre-interpolated filter of phase j shifted by its window offset
 * d_j inside a phase tile.  What changes is who owns what.  The first kernel splits the TAPS across
 * the lanes of a warp: no value is shared between lanes, every shared-memory wavefront carries 32
 * distinct floats, and an 8x8 register tile then needs exactly one wavefront per four FMAs -- the
 * machine's own ratio -- plus a 62-shuffle reduction per 64 outputs.  Here a THREAD owns outputs
 * (10 phases x 8 columns, columns = periods x channels) and walks the taps itself:
 *
 *   - the 10 filter values of a tap are the same for the whole warp (one broadcast LDS.64 per pair),
 *   - the 8 input values are 32 consecutive vectors across the warp (conflict-free),
 *   - 13 wavefronts feed 40 FFMA2 (80 FMAs) per tap and warp, so the FMA pipe binds, not shared memory,
 *   - nobody reduces anything: each accumulator is a finished output.
 *
 * The price is that the input operand must sit in shared memory as X[m][column], i.e. the input read
 * at stride M ("im2col").  It is built on the fly per 32-tap chunk with cp.async (LDGSTS) straight
 * from the caller's interleaved block -- 32 consecutive frames per period, coalesced -- into a
 * double-buffered [32][256 + CV] tile; the row stride is congruent to CV mod 32 so both the vector
 * stores of the staging and the vector loads of the tiles are conflict-free.  The filter chunk
 * (32 taps x 80 phases, contiguous in the table the prep kernel lays out) arrives by TMA bulk copy.
 * Warps skip the chunks in which all of their phases are outside the filter band.
 *
 * CTA = 256 threads = 8 warps = 8 x 10 phases; 32 lanes x (8 / CV) periods x CV channels = 256 columns.
 */
#include <cstdio>
#include <cstdlib>
#include "art_kernels.cuh"
#include "art_device.h"

#define P2_THREADS 256
#define P2_TJ      10                   /* phases per thread                       */
#define P2_JT      80                   /* phases per CTA tile = 8 warps x 10      */
#define P2_KC      32                   /* taps per chunk                          */
#define P2_NT      256                  /* columns per CTA tile                    */

template <int CV> struct P2Vec;
template <> struct P2Vec<1> { typedef float  type; __device__ static float get (const float  &x, int)   { return x; } };
template <> struct P2Vec<2> { typedef float2 type; __device__ static float get (const float2 &x, int v) { return v ? x.y : x.x; } };
template <> struct P2Vec<4> { typedef float4 type; __device__ static float get (const float4 &x, int v) { return v == 0 ? x.x : v == 1 ? x.y : v == 2 ? x.z : x.w; } };

__device__ __forceinline__ unsigned int p2_smem_u32 (const void *p) { return (unsigned int) __cvta_generic_to_shared (p); }
__device__ __forceinline__ void p2_mbar_init (unsigned long long *bar, unsigned int count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(p2_smem_u32 (bar)), "r"(count));
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void p2_mbar_expect_tx (unsigned long long *bar, unsigned int bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(p2_smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p2_mbar_wait (unsigned long long *bar, unsigned int parity)
{
    for (unsigned int spins = 0; spins < (1u << 26); ++spins) {
        unsigned int done;
        asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                      : "=r"(done) : "r"(p2_smem_u32 (bar)), "r"(parity) : "memory");
        if (done) return;
    }
    printf ("libresampler_b200: filter chunk copy never completed (block %d)\n", blockIdx.x);
    __trap ();
}
__device__ __forceinline__ void p2_bulk_g2s (void *dst, const void *src, unsigned int bytes, unsigned long long *bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(p2_smem_u32 (dst)), "l"(src), "r"(bytes), "r"(p2_smem_u32 (bar)) : "memory");
}
template <int BYTES> __device__ __forceinline__ void p2_cp_async (void *dst, const void *src)
{
    if (BYTES == 16)
        asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(p2_smem_u32 (dst)), "l"(src) : "memory");
    else
        asm volatile ("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(p2_smem_u32 (dst)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void p2_cp_commit () { asm volatile ("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void p2_cp_wait_all () { asm volatile ("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int p2_find_job (const ArtJob *jobs, int numJobs, int cta)
{
    int lo = 0, hi = numJobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].tile0 <= cta) lo = mid; else hi = mid - 1;
    }
    return lo;
}

/* ---- prep: phase tables in [tile][tap][phase] layout, shifts, origins, history ---------------------- */
__global__ void __launch_bounds__ (128)
art_periodic2_prep_kernel (const ArtClass k, const ArtPeriodic2 p, const __grid_constant__ ArtJob single,
                           const ArtJob *__restrict__ jobs, int numJobs, int numTables, int histBlocksPerJob)
{
    const int padded = p.PB * P2_JT;
    const int tableBlocks = numTables * padded;
    const int originBlocks = (numJobs * p.PB + 127) / 128;
    const int T = k.T, half = T / 2, F = k.F;
    int b = blockIdx.x;

    if (b < tableBlocks) {
        const int tbl = b / padded, j = b - tbl * padded;
        const int pb = j / P2_JT, jj = j - pb * P2_JT, jb = pb * P2_JT;
        const ArtJob &job = jobs ? jobs[jobs[tbl].repJob] : single;
        __shared__ int sh_row, sh_pass, sh_shift;
        __shared__ double sh_f;
        if (threadIdx.x == 0) {
            ArtLoopState st;
            st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = T;
            long long sj = 0, sb = 0;
            int row = 0, pass = -1;
            double f = 0.0;
            for (int which = 0; which < 2; ++which) {           // 0: the tile's first phase, 1: this phase
                const int ph = which ? j : jb;
                if (ph >= p.L) break;
                int w;
                const double pos = art_output_pos (&st, job.nStart + ph, &w);
                const double whole = floor (pos), fr = pos - whole;
                const long long s = (long long) whole - half + 1 + (long long) w * 15LL * T - job.origin;
                if (!which) { sb = s; continue; }
                sj = s;
                if (k.mode & ART_MODE_INTERP) {
                    double phs = fr * F;                         // resampler.c:1149-1152
                    row = (int) floor (phs);
                    f = phs - row;
                    if (row >= F) { row = F - 1; f = 1.0; }
                }
                else {
                    row = (int) floor (fr * F + 0.5);            // resampler.c:1137
                    if (!(k.mode & ART_MODE_LOWPASS) && row % F == 0)    // resampler.c:1141-1142
                        pass = half - 1 + (row ? 1 : 0);
                }
            }
            sh_row = row; sh_f = f; sh_pass = pass; sh_shift = (int) (sj - sb);
            p.D[(size_t) tbl * padded + j] = j < p.L ? (int) (sj - sb) : -1;
        }
        __syncthreads ();
        const int row = sh_row, pass = sh_pass, shift = sh_shift;
        const double f = sh_f;
        float *dst = p.Hg + ((size_t) tbl * p.PB + pb) * p.Kt * P2_JT;
        const float *ra = k.bank + (size_t) row * k.Tp, *rb = ra + k.Tp;
        for (int m = threadIdx.x; m < p.Kt; m += 128) {
            const int t = m - shift;
            float h = 0.0f;
            if (j < p.L && t >= 0 && t < T) {
                if (pass >= 0)
                    h = t == pass ? 1.0f : 0.0f;
                else if (k.mode & ART_MODE_INTERP) {
                    const double a = ra[t], bb = rb[t];
                    h = (float) (a + f * (bb - a));
                }
                else
                    h = ra[t];
            }
            dst[(size_t) m * P2_JT + jj] = h;
        }
        return;
    }
    b -= tableBlocks;
    if (b < originBlocks) {
        const int e = b * 128 + threadIdx.x;
        if (e >= numJobs * p.PB) return;
        const int seg = e / p.PB, pb = e - seg * p.PB;
        const ArtJob &job = jobs ? jobs[seg] : single;
        ArtLoopState st;
        st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = T;
        int w;
        const double pos = art_output_pos (&st, job.nStart + pb * P2_JT, &w);
        p.S0[e] = (int) ((long long) floor (pos) - half + 1 + (long long) w * 15LL * T - job.origin);
        return;
    }
    b -= originBlocks;
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

/* ---- the product -------------------------------------------------------------------------------------- */
template <int CV>
__global__ void __launch_bounds__ (P2_THREADS, 2)
art_sinc_periodic2_kernel (const ArtClass k, const ArtPeriodic2 p, const __grid_constant__ ArtJob single,
                           const ArtJob *__restrict__ jobs)
{
    typedef typename P2Vec<CV>::type VecT;
    constexpr int NS = 8 / CV;                          // period slots per thread
    constexpr int PT = 32 * NS;                         // periods per CTA tile
    constexpr int XSTR = P2_NT + CV;                    // floats per staged row (== CV mod 32)

    extern __shared__ __align__ (128) unsigned char smem_raw[];
    float *Hs = reinterpret_cast<float *> (smem_raw);                   // [2][KC][JT]
    float *Xs = Hs + 2 * P2_KC * P2_JT;                                  // [2][KC][XSTR]
    __shared__ __align__ (8) unsigned long long bars[2];

    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int groups = (k.C + CV - 1) / CV;
    const int cta = blockIdx.x / groups, cgroup = blockIdx.x - cta * groups;
    const int seg = jobs ? (k.numJobs > 1 ? p2_find_job (jobs, k.numJobs, cta) : 0) : 0;
    const ArtJob &job = jobs ? jobs[seg] : single;
    const int L = p.L, M = p.M, T = k.T;
    const int Q = (int) ((job.outputs + L - 1) / L);
    const int local = cta - job.tile0;
    const int pb = local % p.PB, qt = local / p.PB;     // phase tile fastest: neighbours share input in L2
    const int q0 = qt * PT;
    if (q0 >= Q)
        return;
    const int c0 = cgroup * CV;
    const long long S0 = p.S0[(size_t) seg * p.PB + pb];
    const int NC = p.Kt / P2_KC;
    const float *Hg = p.Hg + ((size_t) job.table * p.PB + pb) * p.Kt * P2_JT;

    // chunks in which this warp's ten phases have any non-zero tap
    int cFirst = NC, cLast = -1;
    {
        const int *D = p.D + (size_t) job.table * p.PB * P2_JT + pb * P2_JT + ty * P2_TJ;
        int dlo = -1, dhi = -1;
        for (int jj = 0; jj < P2_TJ; ++jj) {
            const int d = D[jj];
            if (d >= 0) { if (dlo < 0) dlo = d; dhi = d; }
        }
        if (dlo >= 0) { cFirst = dlo / P2_KC; cLast = (dhi + T - 1) / P2_KC; }
    }

    if (tid == 0) {
        p2_mbar_init (&bars[0], 1);
        p2_mbar_init (&bars[1], 1);
    }
    __syncthreads ();

    // staging: is the whole tile a plain span of the caller's interleaved block?  (job fields are copied to
    // registers here: the asm statements below clobber memory, so anything left in *job would be reloaded
    // for every staged element)
    const int lastQ = min (PT, Q - q0) - 1;             // slow path: periods beyond the segment re-read the last valid one
    const float *const jobIn = job.in;
    const long long inFS = job.inFS;
    const bool vecOk = job.inPlanes == nullptr && job.inCS == 1 && c0 + CV <= k.C &&
                       ((inFS * sizeof (float)) % sizeof (VecT)) == 0 &&
                       ((reinterpret_cast<unsigned long long> (jobIn + c0) % sizeof (VecT)) == 0);
    const long long tileLo = S0 + (long long) M * q0, tileHi = S0 + (long long) M * (q0 + PT - 1) + p.Kt;
    const bool fastX = vecOk && lastQ == PT - 1 && tileLo >= -job.prevAvail && tileHi <= (long long) job.inValid;
    // fast path: thread owns tap m = tid & 31 of periods (tid >> 5) + 8 * it; source and destination advance by
    // constants from one iteration to the next
    const int mLane = tid & (P2_KC - 1), qBase = tid >> 5;
    const float *const srcLane = jobIn + c0 + (S0 + (long long) M * (q0 + qBase) + mLane) * inFS;
    const long long srcStep = 8LL * M * inFS;
    const unsigned int dstLane = (unsigned int) (mLane * XSTR + qBase * CV) * sizeof (float);
    const unsigned int xsBase = p2_smem_u32 (Xs);

    auto stage = [&] (int c, int buf) {
        float *hs = Hs + buf * P2_KC * P2_JT;
        float *xs = Xs + buf * P2_KC * XSTR;
        if (tid == 0) {
            const unsigned int bytes = P2_KC * P2_JT * sizeof (float);
            p2_mbar_expect_tx (&bars[buf], bytes);
            p2_bulk_g2s (hs, Hg + (size_t) c * P2_KC * P2_JT, bytes, &bars[buf]);
        }
        // X[m][q * CV + v] = x[(S0 + M * (q0 + q) + c * KC + m) * C + c0 + v]; lanes walk m (coalesced)
        if (fastX) {
            const float *src = srcLane + (long long) c * P2_KC * inFS;
            unsigned int dst = xsBase + (unsigned int) (buf * P2_KC * XSTR * sizeof (float)) + dstLane;
#pragma unroll
            for (int it = 0; it < PT / 8; ++it) {
                if (CV == 4)
                    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src));
                else
                    asm volatile ("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(dst), "l"(src), "n"(CV * 4));
                src += srcStep;
                dst += 8 * CV * sizeof (float);
            }
        }
        else {
            for (int e = tid; e < P2_KC * PT; e += P2_THREADS) {
                const int m = e & (P2_KC - 1), q = e >> 5;
                const long long a = S0 + (long long) M * (q0 + min (q, lastQ)) + c * P2_KC + m;
                float *dst = xs + m * XSTR + q * CV;
#pragma unroll
                for (int v = 0; v < CV; ++v)
                    dst[v] = (c0 + v < k.C) ? art_fetch (job, T, c0 + v, a) : 0.0f;
            }
        }
        p2_cp_commit ();
    };

    unsigned long long acc2[P2_TJ / 2][8];
#pragma unroll
    for (int a = 0; a < P2_TJ / 2; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) acc2[a][b2] = 0ull;

    /* Chunk order is outside-in (0, NC-1, 1, NC-2, ...), the centre of the window last: a float accumulator
     * loses half an ulp OF THE RUNNING SUM per addition, and the sum only becomes large once the main lobe
     * of the sinc is in.  Same reason the reference sums each dot product from both ends towards the middle
     * (resampler.c:1030-1043).  Measured: max|d|/peak 1.1e-6 in natural order, ~3e-7 this way. */
    auto chunkAt = [&] (int i) { return (i & 1) ? NC - 1 - (i >> 1) : (i >> 1); };
    unsigned int phase[2] = { 0, 0 };
    stage (chunkAt (0), 0);
    for (int i = 0, buf = 0; i < NC; ++i, buf ^= 1) {
        const int c = chunkAt (i);
        p2_cp_wait_all ();
        p2_mbar_wait (&bars[buf], phase[buf]);
        phase[buf] ^= 1;
        __syncthreads ();                               // chunk c visible to all; everyone is done with the previous one
        if (i + 1 < NC)
            stage (chunkAt (i + 1), buf ^ 1);
        if (c < cFirst || c > cLast)
            continue;

        const float *hs = Hs + buf * P2_KC * P2_JT + ty * P2_TJ;
        const VecT *xs = reinterpret_cast<const VecT *> (Xs + buf * P2_KC * XSTR) + tx;
#pragma unroll 4
        for (int m = 0; m < P2_KC; ++m) {
            unsigned long long h2[P2_TJ / 2];
#pragma unroll
            for (int pp = 0; pp < P2_TJ / 2; ++pp) {
                const float2 hv = *reinterpret_cast<const float2 *> (hs + m * P2_JT + 2 * pp);   // warp-uniform: broadcast
                h2[pp] = art_pack2 (hv.x, hv.y);
            }
#pragma unroll
            for (int r = 0; r < NS; ++r) {
                const VecT xv = *reinterpret_cast<const VecT *> (reinterpret_cast<const float *> (xs) + m * XSTR + r * 32 * CV);
#pragma unroll
                for (int v = 0; v < CV; ++v) {
                    const float x = P2Vec<CV>::get (xv, v);
                    const unsigned long long x2 = art_pack2 (x, x);
#pragma unroll
                    for (int pp = 0; pp < P2_TJ / 2; ++pp)
                        art_ffma2 (acc2[pp][r * CV + v], h2[pp], x2);
                }
            }
        }
    }

    /* every accumulator is a finished output: phases j0 .. j0+9 (consecutive frames) x NS periods x CV channels */
    const int j0 = pb * P2_JT + ty * P2_TJ;
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        const int q = q0 + tx + 32 * r;
        if (q >= Q) continue;
#pragma unroll
        for (int pp = 0; pp < P2_TJ / 2; ++pp) {
            float lo[CV], hi[CV];
#pragma unroll
            for (int v = 0; v < CV; ++v)
                art_unpack2 (acc2[pp][r * CV + v], lo[v], hi[v]);
#pragma unroll
            for (int hsel = 0; hsel < 2; ++hsel) {
                const int j = j0 + 2 * pp + hsel;
                const long long nl = (long long) q * L + j;
                if (j >= L || nl >= (long long) job.outputs) continue;
#pragma unroll
                for (int v = 0; v < CV; ++v)
                    if (c0 + v < k.C)
                        *art_out_ptr (job, c0 + v, (long long) job.nStart + nl) = hsel ? hi[v] : lo[v];
            }
        }
    }
}

/* ---- host side ------------------------------------------------------------------------------------------ */

static size_t periodic2_smem (int CV)
{
    return (size_t) 2 * (P2_KC * P2_JT + P2_KC * (P2_NT + CV)) * sizeof (float) + 128;
}

bool artPlanPeriodic2 (const ArtClass &k, double ratio, unsigned int maxOutputs, ArtPeriodic2 &p, int &CV)
{
    int L, M;
    if (k.mode & ART_MODE_PRECISE) return false;
    if (!artRational (ratio, 1024, &L, &M)) return false;
    if (L < 20) return false;                           // a tile is 80 phases wide; tiny L: first kernel / generic
    if (maxOutputs < (unsigned) (4 * L)) return false;
    CV = k.C >= 4 ? 4 : (k.C >= 2 ? 2 : 1);
    p.L = L; p.M = M;
    p.PB = (L + P2_JT - 1) / P2_JT;
    const int spread = (int) (((long long) (P2_JT - 1) * M + L - 1) / L) + 2;
    p.Kt = (k.T + spread + P2_KC - 1) / P2_KC * P2_KC;
    if (getenv ("ART_B200_TRACE"))
        fprintf (stderr, "[art] periodic2 L=%d M=%d PB=%d Kt=%d CV=%d smem=%zu\n", p.L, p.M, p.PB, p.Kt, CV, periodic2_smem (CV));
    return true;
}

int artPeriodic2Ctas (const ArtPeriodic2 &p, int CV, unsigned int outputs)
{
    const long long Q = ((long long) outputs + p.L - 1) / p.L;
    const int PT = 32 * (8 / CV);
    return (int) (p.PB * ((Q + PT - 1) / PT));
}

size_t artPeriodic2TableBytes (const ArtPeriodic2 &p, int numTables, int numJobs)
{
    return ((size_t) numTables * p.PB * p.Kt * P2_JT) * sizeof (float) + ((size_t) numTables * p.PB * P2_JT + (size_t) numJobs * p.PB) * sizeof (int);
}

void artPeriodic2Carve (ArtPeriodic2 &p, void *tables, int numTables, int numJobs)
{
    p.Hg = reinterpret_cast<float *> (tables);
    p.D = reinterpret_cast<int *> (p.Hg + (size_t) numTables * p.PB * p.Kt * P2_JT);
    p.S0 = p.D + (size_t) numTables * p.PB * P2_JT;
    (void) numJobs;
}

template <int CV>
static void launch_periodic2 (const ArtClass &k, const ArtPeriodic2 &p, int totalCtas, int numJobs, int numTables,
                              const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    auto kern = art_sinc_periodic2_kernel<CV>;
    static bool configured[16] = { false };
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    const size_t smem = periodic2_smem (CV);
    if (!configured[device & 15]) {
        ART_CUDA_CHECK (cudaFuncSetAttribute (kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        configured[device & 15] = true;
    }
    int histBlocks = (k.C * k.T + 127) / 128;
    if (histBlocks > 32) histBlocks = 32;
    const int prepBlocks = numTables * p.PB * P2_JT + (numJobs * p.PB + 127) / 128 + numJobs * histBlocks;
    art_periodic2_prep_kernel<<<prepBlocks, 128, 0, stream>>> (k, p, single, d_jobs, numJobs, numTables, histBlocks);
    ART_CUDA_CHECK (cudaGetLastError ());
    const unsigned int grid = (unsigned int) totalCtas * (unsigned int) ((k.C + CV - 1) / CV);
    void *prof;
    artProfileBegin (stream, &prof);
    kern<<<grid, P2_THREADS, smem, stream>>> (k, p, single, d_jobs);
    artProfileEnd (stream, prof);
    ART_CUDA_CHECK (cudaGetLastError ());
    g_artLaunches += 2;
}

void artLaunchPeriodic2 (const ArtClass &k, const ArtPeriodic2 &p, int CV, int totalCtas, int numJobs, int numTables,
                         const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    if (totalCtas <= 0) return;
    if (CV == 4) launch_periodic2<4> (k, p, totalCtas, numJobs, numTables, single, d_jobs, stream);
    else if (CV == 2) launch_periodic2<2> (k, p, totalCtas, numJobs, numTables, single, d_jobs, stream);
    else launch_periodic2<1> (k, p, totalCtas, numJobs, numTables, single, d_jobs, stream);
}
