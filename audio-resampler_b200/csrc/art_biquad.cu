/*
 * art_biquad.cu -- the reference's biquad recurrence (biquad.c:106-163) on the GPU.
 *
 * A direct-form-I section is a serial recurrence, and the reference has only `channels`
 * independent chains (64 in BASELINE config 3).  Time is therefore cut into chunks and the
 * cascade of S sections is treated as one linear system with state z = (delayed inputs and
 * outputs of every section):
 *
 *   A. zero-state pass   (thread per channel x chunk, double): run the chunk from z = 0
 *                         -> the forced response's end state  ZS[c][k]
 *   M. transition pass   (thread per channel x state component, double): run the chunk
 *                         length with zero input from each unit state -> matrix M[c]
 *   B. propagation       (one warp per channel): z_{k+1} = ZS_k + M z_k, storing the true
 *                         state at the START of every chunk
 *   C. output pass       (thread per channel x chunk, float): rerun the chunk from its true
 *                         start state with the reference's float arithmetic -- same
 *                         expression order, unfused multiply/add -- writing in place.
 *
 * Passes A/M/B run in double so the start states carry no more error than the reference's
 * own float recurrence does (float direct form is ~2e-7 from exact);
 * pass C is bit-faithful to the reference given its start state.  The buffer is read
 * twice and written once: 12 bytes per sample per cascade, independent of S.
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "art_kernels.cuh"
#include "art_device.h"

namespace {

struct BqGeom {
    long long frames;
    int stride, channels, chunk, numChunks, lastLen;
};

template <typename R, int S, int ORD, int HIST = ORD>
struct Cascade {
    R a[S][ORD + 1], b[S][ORD + 1];
    R xh[S][HIST], yh[S][HIST];        // [0] = newest; the arithmetic reads the first ORD, the rest
                                       // only mirrors the 4-deep rings of the reference's struct

    __device__ void loadCoeffs (const ArtBiquadStage *st, int channels, int c)
    {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const ArtBiquadStage &q = st[s * channels + c];
#pragma unroll
            for (int d = 0; d <= ORD; ++d) { a[s][d] = q.a[d]; b[s][d] = q.b[d]; }
        }
    }
    __device__ void zero ()
    {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < HIST; ++d) { xh[s][d] = 0; yh[s][d] = 0; }
    }
    template <typename Src> __device__ void loadState (const Src *z)
    {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < ORD; ++d) { xh[s][d] = (R) z[(s * 2) * ORD + d]; yh[s][d] = (R) z[(s * 2 + 1) * ORD + d]; }
    }
    template <typename Dst> __device__ void storeState (Dst *z) const
    {
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < ORD; ++d) { z[(s * 2) * ORD + d] = (Dst) xh[s][d]; z[(s * 2 + 1) * ORD + d] = (Dst) yh[s][d]; }
    }
    /* one sample through all sections; expression order of biquad.c:141-143 */
    __device__ __forceinline__ R step (R in);
};

template <typename R, int S, int ORD, int HIST>
__device__ __forceinline__ R Cascade<R, S, ORD, HIST>::step (R in)
{
#pragma unroll
    for (int s = 0; s < S; ++s) {
        R acc;
        if constexpr (sizeof (R) == 4) {
            acc = __fmul_rn (in, a[s][0]);
#pragma unroll
            for (int d = 1; d <= ORD; ++d)
                acc = __fsub_rn (__fadd_rn (acc, __fmul_rn (xh[s][d - 1], a[s][d])), __fmul_rn (b[s][d], yh[s][d - 1]));
        }
        else {
#if ART_WIDE            /* the wide build's output pass runs here: the reference's operation order, nothing contracted */
            acc = __dmul_rn (in, a[s][0]);
#pragma unroll
            for (int d = 1; d <= ORD; ++d)
                acc = __dsub_rn (__dadd_rn (acc, __dmul_rn (xh[s][d - 1], a[s][d])), __dmul_rn (b[s][d], yh[s][d - 1]));
#else
            acc = in * a[s][0];
#pragma unroll
            for (int d = 1; d <= ORD; ++d)
                acc = acc + xh[s][d - 1] * a[s][d] - b[s][d] * yh[s][d - 1];
#endif
        }
#pragma unroll
        for (int d = HIST - 1; d > 0; --d) { xh[s][d] = xh[s][d - 1]; yh[s][d] = yh[s][d - 1]; }
        xh[s][0] = in;
        yh[s][0] = acc;
        in = acc;
    }
    return in;
}

template <int S, int ORD>
__global__ void bq_zero_state_kernel (BqGeom g, const ArtBiquadStage *st, const artsample_t *buf, double *ZS)
{
    constexpr int NST = 2 * S * ORD;
    const long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x;     // channel fastest
    const int k = (int) (e / g.channels), c = (int) (e - (long long) k * g.channels);
    if (k >= g.numChunks) return;
    Cascade<double, S, ORD> f;
    f.loadCoeffs (st, g.channels, c);
    f.zero ();
    const long long f0 = (long long) k * g.chunk;
    const int len = k == g.numChunks - 1 ? g.lastLen : g.chunk;
    const artsample_t *p = buf + f0 * g.stride + c;
    // the recurrence is one dependent chain per thread: fetch 16 samples at a time so that a load's latency is paid once
    // per batch, not once per step (the un-batched loop ran at 0.7 TB/s)
    for (int i0 = 0; i0 < len; i0 += 16) {
        artsample_t v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = i0 + r < len ? __ldg (p + (long long) (i0 + r) * g.stride) : (artsample_t) 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) if (i0 + r < len) f.step ((double) v[r]);
    }
    f.storeState (ZS + ((long long) c * g.numChunks + k) * NST);
}

template <int S, int ORD>
__global__ void bq_transition_kernel (BqGeom g, const ArtBiquadStage *st, double *M)
{
    constexpr int NST = 2 * S * ORD;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = e / NST, j = e - c * NST;
    if (c >= g.channels) return;
    Cascade<double, S, ORD> f;
    f.loadCoeffs (st, g.channels, c);
    double z[NST];
    double *Mc = M + (long long) c * 2 * NST * NST;
#pragma unroll
    for (int which = 0; which < 2; ++which) {               // 0: full chunk, 1: the (shorter) last chunk
        const int len = which ? g.lastLen : g.chunk;
#pragma unroll
        for (int i = 0; i < NST; ++i) z[i] = i == j ? 1.0 : 0.0;
        f.loadState (z);
        for (int i = 0; i < len; ++i)
            f.step (0.0);
        f.storeState (z);
#pragma unroll
        for (int i = 0; i < NST; ++i)
            Mc[(which * NST + i) * NST + j] = z[i];
    }
}

/* One block per channel.  The start state of chunk k+1 is M z_k + zs_k (M: the chunk's transition, zs_k: its zero-state
 * response) -- a serial chain over the chunks, which a single warp walked in 0.35 us per chunk (the longest of the four
 * kernels).  Two levels instead: warp w takes a group of G consecutive chunks; (1) every warp runs its group from a zero start
 * state, (2) warp 0 chains the groups with M^G, (3) every warp re-runs its group from its true start and writes the chunk start
 * states.  Serial depth G + groups + G instead of numChunks.  Lane r owns component r of the state. */
template <int S, int ORD>
__global__ void __launch_bounds__ (512)
bq_propagate_kernel (BqGeom g, const ArtBiquadStage *st, const double *ZS, const double *M, double *Zstart)
{
    constexpr int NST = 2 * S * ORD;
    __shared__ double shM[NST * NST], shP[NST * NST], shY[16][NST], shZ[16][NST];
    const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int r = lane < NST ? lane : NST - 1;
    const int G = (g.numChunks + warps - 1) / warps;                        // chunks per group
    const int groups = (g.numChunks + G - 1) / G;
    const double *Mc = M + (long long) c * 2 * NST * NST;                   // [0]: a full chunk's transition
    for (int i = threadIdx.x; i < NST * NST; i += blockDim.x) shM[i] = Mc[i];
    double row[NST];
#pragma unroll
    for (int j = 0; j < NST; ++j) row[j] = Mc[r * NST + j];
    __syncthreads ();
    const int k0 = warp * G, k1 = min (k0 + G, g.numChunks);
    const double *zs = ZS + (long long) c * g.numChunks * NST;
    auto advance = [&] (double z, int k) -> double {                        // M z + zs_k, component r
        double acc = zs[(long long) k * NST + r];
#pragma unroll
        for (int j = 0; j < NST; ++j) acc += row[j] * __shfl_sync (0xffffffffu, z, j);
        return acc;
    };
    if (warp < groups) {
        double z = 0.0;
        for (int k = k0; k < k1; ++k) z = advance (z, k);
        if (lane < NST) shY[warp][lane] = z;
    }
    __syncthreads ();
    if (warp == 0) {
        // row r of M^G by G - 1 multiplications with M (G <= a few dozen, NST <= 32); the power lives in shared memory
        if (lane < NST)
#pragma unroll
            for (int j = 0; j < NST; ++j) shP[lane * NST + j] = row[j];
        __syncwarp ();
        for (int t = 1; t < G; ++t) {
            double nx[NST];
#pragma unroll
            for (int j = 0; j < NST; ++j) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < NST; ++i) a += shP[r * NST + i] * shM[i * NST + j];
                nx[j] = a;
            }
            __syncwarp ();
            if (lane < NST)
#pragma unroll
                for (int j = 0; j < NST; ++j) shP[lane * NST + j] = nx[j];
            __syncwarp ();
        }
        // start state of the whole call: the caller's Biquad structs
        double z;
        {
            const int s = r / (2 * ORD), rem = r - s * 2 * ORD;
            const ArtBiquadStage &q = st[s * g.channels + c];
            z = rem < ORD ? (double) q.x[rem] : (double) q.y[rem - ORD];
        }
        for (int w = 0; w < groups; ++w) {
            if (lane < NST) shZ[w][lane] = z;
            double acc = shY[w][r];
#pragma unroll
            for (int j = 0; j < NST; ++j) acc += shP[r * NST + j] * __shfl_sync (0xffffffffu, z, j);
            z = acc;
        }
    }
    __syncthreads ();
    if (warp < groups) {
        double z = shZ[warp][r];
        for (int k = k0; k < k1; ++k) {
            if (lane < NST) Zstart[((long long) c * g.numChunks + k) * NST + r] = z;
            z = advance (z, k);
        }
    }
}

template <int S, int ORD>
__global__ void bq_output_kernel (BqGeom g, const ArtBiquadStage *st, ArtBiquadStage *stOut, artsample_t *buf, const double *Zstart)
{
    constexpr int NST = 2 * S * ORD;
    const long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x;     // channel fastest
    const int k = (int) (e / g.channels), c = (int) (e - (long long) k * g.channels);
    if (k >= g.numChunks) return;
    Cascade<artsample_t, S, ORD, 4> f;
    f.loadCoeffs (st, g.channels, c);
    f.zero ();
    if (k == 0) {
        // the call's start state is the caller's struct itself, all four delays of it
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < 4; ++d) { f.xh[s][d] = st[s * g.channels + c].x[d]; f.yh[s][d] = st[s * g.channels + c].y[d]; }
    }
    else {
        const double *z = Zstart + ((long long) c * g.numChunks + k) * NST;
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < ORD; ++d) { f.xh[s][d] = (artsample_t) z[(s * 2) * ORD + d]; f.yh[s][d] = (artsample_t) z[(s * 2 + 1) * ORD + d]; }
    }
    const long long f0 = (long long) k * g.chunk;
    const int len = k == g.numChunks - 1 ? g.lastLen : g.chunk;
    artsample_t *p = buf + f0 * g.stride + c;
    for (int i0 = 0; i0 < len; i0 += 16) {
        artsample_t v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = i0 + r < len ? p[(long long) (i0 + r) * g.stride] : (artsample_t) 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) if (i0 + r < len) v[r] = f.step (v[r]);
#pragma unroll
        for (int r = 0; r < 16; ++r) if (i0 + r < len) p[(long long) (i0 + r) * g.stride] = v[r];
    }
    if (k == g.numChunks - 1) {
        // leave the state the reference would leave: the four newest inputs/outputs of every
        // section (the last chunk is >= 4 samples unless the whole call is shorter, in which
        // case k == 0 and the older entries are the caller's own, shifted)
#pragma unroll
        for (int s = 0; s < S; ++s) {
            ArtBiquadStage &q = stOut[s * g.channels + c];
#pragma unroll
            for (int d = 0; d < 4; ++d) { q.x[d] = f.xh[s][d]; q.y[d] = f.yh[s][d]; }
        }
    }
}

template <int S, int ORD>
void run_cascade (const BqGeom &g, const ArtBiquadStage *d_st, ArtBiquadStage *d_stOut, artsample_t *d_buf, cudaStream_t stream)
{
    constexpr int NST = 2 * S * ORD;
    static_assert (NST <= 32, "state must fit one warp");
    double *ZS = nullptr, *Zstart = nullptr, *M = nullptr;
    const size_t stateBytes = sizeof (double) * (size_t) g.channels * g.numChunks * NST;
    ART_CUDA_CHECK (cudaMallocAsync (&ZS, stateBytes, stream));
    ART_CUDA_CHECK (cudaMallocAsync (&Zstart, stateBytes, stream));
    ART_CUDA_CHECK (cudaMallocAsync (&M, sizeof (double) * (size_t) g.channels * 2 * NST * NST, stream));

    const long long work = (long long) g.channels * g.numChunks;
    const int tx = 64;
    const unsigned int gridCK = (unsigned int) ((work + tx - 1) / tx);
    bq_zero_state_kernel<S, ORD><<<gridCK, tx, 0, stream>>> (g, d_st, d_buf, ZS);
    bq_transition_kernel<S, ORD><<<(g.channels * NST + 63) / 64, 64, 0, stream>>> (g, d_st, M);
    bq_propagate_kernel<S, ORD><<<g.channels, g.numChunks >= 128 ? 512 : 128, 0, stream>>> (g, d_st, ZS, M, Zstart);
    bq_output_kernel<S, ORD><<<gridCK, tx, 0, stream>>> (g, d_st, d_stOut, d_buf, Zstart);
    ART_CUDA_CHECK (cudaGetLastError ());
    g_artLaunches += 4;

    ART_CUDA_CHECK (cudaFreeAsync (ZS, stream));
    ART_CUDA_CHECK (cudaFreeAsync (Zstart, stream));
    ART_CUDA_CHECK (cudaFreeAsync (M, stream));
}

template <int S>
void run_cascade_order (int order, const BqGeom &g, const ArtBiquadStage *d_st, ArtBiquadStage *d_stOut, artsample_t *d_buf, cudaStream_t stream)
{
    if (order <= 2) run_cascade<S, 2> (g, d_st, d_stOut, d_buf, stream);
    else            run_cascade<S, 4> (g, d_st, d_stOut, d_buf, stream);
}

}   // namespace

extern "C" int artBiquadRun (ArtBiquadStage *stages, int numStages, int channels, artsample_t *buffer,
                             long long frames, int stride, int onDevice, void *streamPtr)
{
    ART_GUARD_BEGIN
    if (frames <= 0 || numStages <= 0 || channels <= 0)
        return 0;
    int count = 0;
    if (cudaGetDeviceCount (&count) != cudaSuccess || count == 0) {
        artRaise ("biquad needs a CUDA device; this library has no CPU path");
    }
    cudaStream_t stream = (cudaStream_t) streamPtr;
    const size_t span = (size_t) (frames - 1) * stride + channels;

    artsample_t *d_buf = buffer;
    if (!onDevice) {
        ART_CUDA_CHECK (cudaMallocAsync (&d_buf, span * sizeof (artsample_t), stream));
        ART_CUDA_CHECK (cudaMemcpyAsync (d_buf, buffer, span * sizeof (artsample_t), cudaMemcpyHostToDevice, stream));
    }
    ArtBiquadStage *d_st = nullptr, *d_stOut = nullptr;
    const size_t stBytes = sizeof (ArtBiquadStage) * (size_t) numStages * channels;
    ART_CUDA_CHECK (cudaMallocAsync (&d_st, stBytes, stream));
    ART_CUDA_CHECK (cudaMallocAsync (&d_stOut, stBytes, stream));
    ART_CUDA_CHECK (cudaMemcpyAsync (d_st, stages, stBytes, cudaMemcpyHostToDevice, stream));

    BqGeom g;
    g.frames = frames; g.stride = stride; g.channels = channels;
    int chunk = 64;
    while (chunk < 2048 && frames / chunk > 1024) chunk <<= 1;
    g.chunk = chunk;
    g.numChunks = frames / chunk > 0 ? (int) (frames / chunk) : 1;          // the last chunk absorbs the remainder,
    g.lastLen = (int) (frames - (long long) (g.numChunks - 1) * chunk);     // so it is never shorter than `chunk`

    int order = 1;
    for (int i = 0; i < numStages * channels; ++i)
        if (stages[i].order > order) order = stages[i].order;

    // sections are applied four at a time (state of 4 order-4 sections = 32 = one warp)
    for (int s0 = 0; s0 < numStages; s0 += 4) {
        const int S = numStages - s0 < 4 ? numStages - s0 : 4;
        const ArtBiquadStage *base = d_st + (size_t) s0 * channels;
        ArtBiquadStage *baseOut = d_stOut + (size_t) s0 * channels;
        switch (S) {
            case 1: run_cascade_order<1> (order, g, base, baseOut, d_buf, stream); break;
            case 2: run_cascade_order<2> (order, g, base, baseOut, d_buf, stream); break;
            case 3: run_cascade_order<3> (order, g, base, baseOut, d_buf, stream); break;
            default: run_cascade_order<4> (order, g, base, baseOut, d_buf, stream); break;
        }
    }

    // the four newest inputs/outputs of every section come back into the caller's structs
    std::vector<ArtBiquadStage> back ((size_t) numStages * channels);
    ART_CUDA_CHECK (cudaMemcpyAsync (back.data (), d_stOut, stBytes, cudaMemcpyDeviceToHost, stream));
    if (!onDevice)
        ART_CUDA_CHECK (cudaMemcpyAsync (buffer, d_buf, span * sizeof (artsample_t), cudaMemcpyDeviceToHost, stream));
    ART_CUDA_CHECK (cudaStreamSynchronize (stream));
    for (size_t i = 0; i < back.size (); ++i)
        for (int d = 0; d < 4; ++d) { stages[i].x[d] = back[i].x[d]; stages[i].y[d] = back[i].y[d]; }
    ART_CUDA_CHECK (cudaFreeAsync (d_st, stream));
    ART_CUDA_CHECK (cudaFreeAsync (d_stOut, stream));
    if (!onDevice)
        ART_CUDA_CHECK (cudaFreeAsync (d_buf, stream));
    return 0;
    ART_GUARD_END (-1)
}
