/*
 * art_kernels.cuh -- structures shared by the CUDA translation units (sm_100a only).
 *
 * HBM layout
 *   bank      : (F+2) rows x Tp floats, Tp = T rounded up to 32, zero padded.  Row r, tap t is
 *               filters[r][t] of the reference (resampler.c:146-168); row F+1 is all zero so a
 *               clamped row index can never read out of bounds.  128-byte aligned rows.
 *   history   : C x T floats, planar -- the newest T samples each channel has consumed.  This
 *               replaces the reference's 16*T ring + memmove compaction (resampler.c:139,
 *               :497-503): an output window never reaches further back than T samples before
 *               the first sample of the current call (art_plan.h), so "history ++ input block"
 *               is all a call can touch.
 *   input     : caller's block, interleaved [frame][C] or planar (frame stride / channel stride,
 *               or a table of per-channel pointers).
 *   output    : likewise.
 *
 * Sample coordinates: "region index" i counts frames from the first frame of the call's input
 * region (i >= 0: this call's input, zero beyond inValid; i < 0: history, or -- in block mode --
 * the frames that precede this block in the same buffer).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <exception>
#include <atomic>
#include "art_plan.h"
#include "art_sample.h"

struct ArtJob {
    double        P, ratio;          // loop-entry outputOffset, effective ratio
    int           I;                 // loop-entry inputIndex
    int           origin;            // ring index that region index 0 maps to (I - pre)
    unsigned int  outputs;           // frames to produce
    int           inValid;           // frames of caller data in the input region
    long long     prevAvail;         // consumed frames preceding the region in the same buffer
    long long     consumed;          // region frames that enter the history (history kernel)
    int           tile0;             // first tile / CTA of this job inside the launch
    unsigned int  nStart;            // call-relative index of this job's first output (segments of one call)
    const artsample_t *hist;         // [C][T] history at call entry
    artsample_t  *histOut;           // [C][T] history after the call (may be null)
    const artsample_t *in;           // base of region frame 0, channel 0
    artsample_t  *out;
    long long     inFS, inCS, outFS, outCS;      // frame / channel strides in floats
    const artsample_t *const *inPlanes;    // optional per-channel pointer tables (device memory)
    artsample_t *const *outPlanes;
    int           table;             // periodic kernel: which phase table this job reads
    int           repJob;            // periodic kernel: entry t holds the job whose state defines table t
};

struct ArtClass {
    const artsample_t *bank;
    int T, Tp, F, C, mode;   // T = taps of a bank row = depth of the history (Tref + lead)
    int Tref;        // numTaps of the reference context: what the control loop's position arithmetic (art_plan.h) runs on
    int lead;        // taps in front of the reference's window: a folded-in pre-filter (art_context.c) extends every filter into the
                     //   past, so a window starts at floor (pos) - Tref / 2 + 1 - lead and has T taps
    int NB;          // output frames per tile
    int Cg;          // channels per CTA (smem planes; a multiple of the CV the kernel was built for)
    int Wp;          // floats per smem plane
    int numJobs;
    int sort;        // 1: group a tile's outputs by filter row before convolving
    int unity;       // 1: every ratio of the launch is so close to 1 that consecutive outputs share a filter-row pair for >= 8 outputs
                     //    on average (asynchronous sample-rate conversion): the any-ratio kernel then register-blocks consecutive outputs
    float absSum;    // max over bank rows of sum |tap|: bounds sum |h| of any interpolated filter
};

/* rational-ratio kernel: per-launch geometry and the phase tables it reads (art_sinc_periodic.cu) */
struct ArtPeriodic {
    int L, M;            // outputs / inputs per period (ratio = L / M in lowest terms)
    int rowsPerCta;      // rows of 8 phases per CTA
    int Kp;              // union-window taps per phase block, multiple of 32
    int Qc;              // periods per staged chunk
    int Qblk;            // periods per CTA
    int Wc;              // staged samples per chunk = (Qc - 1) * M + Kp
    int PB;              // phase blocks = ceil(ceil(L / 8) / rowsPerCta)
    float *Hblk;         // [tables][PB][rowsPerCta*8*Kp]  interpolated filters of a phase block, already
                         //   shifted to the block's origin and in the kernel's shared-memory layout
    int   *S0;           // [jobs][PB]  region index of the block's first tap, period 0
};

/* tensor-core form of the rational-ratio kernel (art_sinc_umma.cu): y[period, phase] as a product of the
 * input read at stride M (rows = periods) and the banded matrix of pre-interpolated filters, on tcgen05 */
#define ART_U_MAXK 96
struct ArtUmma {
    int L, M;            // outputs / inputs per period
    int G, Lg;           // the L phases are handled in G groups of Lg (tensor memory holds 160 accumulator columns x 3)
    int Npad;            // phases of a group rounded up to 16: the N of every MMA
    int KI;              // k-steps (16 taps) per input period: ceil(M / 16) = pairs of 8-tap planes per row
    int NS;              // plane-pair slots of operand A held in shared memory (a divisor of KI; pair i lives in slot i % NS)
    int numK;            // k-steps per tile
    int cg;              // channels per tile (1, 2, 4): MMA row = cg * period + channel
    int periods;         // periods per tile incl. the largest row shift: 128 / cg + aMax
    int digits;          // signal digits (2 or 3): splits of operand A, 5 or 6 MMAs per k-step
    int rows;            // 16-byte row slots of a plane of the signal operand: cg * periods, padded
    int DH;              // filter quantum is 2^-DH
    int stages;          // depth of the filter stage ring
    int depth;           // raw plane pairs the converters keep in flight (cp.async ring in shared memory): 2 .. 4
    int tableHalfs;      // fp16 elements per table: numK * 3 * 2 * Npad * 8
    unsigned short *H;   // [tables][G][numK][3 splits][2 k-planes][Npad][8]  fp16 bit patterns
    int *S0;             // [jobs][G]  region index of tap 0 of the group's first phase, period 0
    int *tileJob;        // [tiles] index of the job a tile belongs to (written by the prep kernel: the product kernel's roles
                         //   then find their job with one load instead of a binary search over the job list)
    unsigned char ka[ART_U_MAXK], ki[ART_U_MAXK];      // k-step -> (row shift a, 16-tap group i), i outermost
    unsigned char nA[32];                              // row shifts per 16-tap group
};

/* Sum NV register values per lane across the warp so that lane L ends up with the total of
 * value (L * NV / 32).  log2(NV) exchange stages halve the value count while consuming one
 * lane bit each; the remaining lane bits are folded with a plain butterfly. */
template <int NV, typename AccT>
__device__ __forceinline__ AccT art_transpose_reduce (AccT (&v)[NV], int lane)
{
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            AccT send = upper ? v[i] : v[i + n / 2];
            AccT keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync (0xffffffffu, send, off);
        }
    }
#pragma unroll
    for (; off >= 1; off >>= 1)
        v[0] += __shfl_xor_sync (0xffffffffu, v[0], off);
    return v[0];
}

/* packed FP32 pairs: fma.rn.f32x2 (SASS FFMA2) performs two IEEE fused multiply-adds per instruction;
 * ptxas turns a {x, x} operand into a scalar broadcast, so (pair) x (scalar) + (pair) costs one issue slot */
__device__ __forceinline__ unsigned long long art_pack2 (float lo, float hi)
{
    unsigned long long r;
    asm ("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void art_unpack2 (unsigned long long v, float &lo, float &hi)
{
    asm ("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void art_ffma2 (unsigned long long &acc, unsigned long long a, unsigned long long b)
{
    asm ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

/* sample arithmetic in the reference's operation order, never contracted into FMAs (the byte / state parity of the integer and
 * biquad stages depends on it), for either sample width */
__device__ __forceinline__ float  art_mul (float a, float b)   { return __fmul_rn (a, b); }
__device__ __forceinline__ float  art_add (float a, float b)   { return __fadd_rn (a, b); }
__device__ __forceinline__ float  art_sub (float a, float b)   { return __fsub_rn (a, b); }
__device__ __forceinline__ double art_mul (double a, double b) { return __dmul_rn (a, b); }
__device__ __forceinline__ double art_add (double a, double b) { return __dadd_rn (a, b); }
__device__ __forceinline__ double art_sub (double a, double b) { return __dsub_rn (a, b); }

__device__ __forceinline__ artsample_t art_fetch (const ArtJob &j, int T, int c, long long idx)
{
    if (idx >= j.inValid)
        return 0;
    if (idx >= -j.prevAvail) {
        const artsample_t *p = j.inPlanes ? j.inPlanes[c] + idx * j.inFS : j.in + idx * j.inFS + c * j.inCS;
        return __ldg (p);
    }
    const long long h = T + idx + j.prevAvail;
    return h >= 0 ? j.hist[(long long) c * T + h] : (artsample_t) 0;
}

__device__ __forceinline__ artsample_t *art_out_ptr (const ArtJob &j, int c, long long frame)
{
    return j.outPlanes ? j.outPlanes[c] + frame * j.outFS : j.out + frame * j.outFS + c * j.outCS;
}

/* The job a tile / CTA belongs to: the last job whose first tile is <= tile (jobs[0].tile0 == 0).  Called by whole warps with a
 * warp-uniform `tile`: every lane probes one candidate (a 33-ary search), so a launch of a thousand jobs -- an ASRC sequence of
 * 1024 blocks, BASELINE config 4's contexts -- costs two dependent loads per CTA instead of ten. */
__device__ __forceinline__ int art_find_job_warp (const ArtJob *jobs, int numJobs, int tile)
{
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = numJobs - 1;
    while (lo < hi) {
        const int step = (hi - lo + 31) / 32;
        int idx = lo + (lane + 1) * step;
        if (idx > hi) idx = hi;
        const bool ok = jobs[idx].tile0 <= tile;                 // monotone in the lane index
        const int c = __popc (__ballot_sync (0xffffffffu, ok));
        if (c == 32) { lo = hi; break; }
        const int firstFalse = min (lo + (c + 1) * step, hi);
        if (c) lo = min (lo + c * step, hi);
        hi = firstFalse - 1;
    }
    return lo;
}

/* launchers (host side, C++ linkage, defined next to their kernels) */
struct ArtLaunchGeom { int totalTiles; size_t smemBytes; int CV; };

void artPlanGenericGeometry (ArtClass &k, double minRatio, unsigned int maxOutputs, int smCount, ArtLaunchGeom &g);
/* `single` is used when d_jobs is null (one job, passed by value); otherwise d_jobs[numJobs] */
void artLaunchGeneric (const ArtClass &k, const ArtLaunchGeom &g, const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream);
void artLaunchHistory (const ArtClass &k, const ArtJob &single, const ArtJob *d_jobs, int numJobs, cudaStream_t stream);

bool artRational (double ratio, int maxL, int *L, int *M);
bool artPlanPeriodic (const ArtClass &k, double ratio, unsigned int maxOutputs, unsigned long long totalOutputs,
                      int smCount, ArtPeriodic &p, int &CV);
unsigned int artPeriodicSegmentOutputs (const ArtPeriodic &p, double ratio);
int  artPeriodicCtas (const ArtPeriodic &p, unsigned int outputs);
void artLaunchPeriodic (const ArtClass &k, const ArtPeriodic &p, int CV, int totalCtas, int numJobs, int numTables,
                        const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream);

bool artPlanUmma (const ArtClass &k, double ratio, unsigned int maxOutputs, unsigned long long totalOutputs,
                  int smCount, ArtUmma &u);
int  artUmmaTiles (const ArtUmma &u, int channels, unsigned int outputs);
size_t artUmmaTableBytes (const ArtUmma &u, int numTables, int numJobs, int totalTiles);
void artUmmaCarve (ArtUmma &u, void *tables, int numTables, int numJobs);
void artLaunchUmma (const ArtClass &k, const ArtUmma &u, int totalTiles, int numJobs, int numTables, int smCount,
                    const ArtJob &single, const ArtJob *d_jobs, bool vecIn, cudaStream_t stream);

extern std::atomic<unsigned long long> g_artLaunches;
extern int g_artTensorMode;
extern int g_artTensorDigits;

/* per-kernel event timing, active only after artDevProfileEnable(1) */
void artProfileBegin (cudaStream_t stream, void **token);
void artProfileEnd (cudaStream_t stream, void *token);

/* Error model: no CUDA failure may take the caller's process down.  A failed runtime call (or an unsupported request)
 * records a message (artDevLastError), prints it once to stderr -- as the reference does for its own init errors,
 * resampler.c:127-135 -- and unwinds to the extern "C" entry point, which reports failure to art_context.c. */
struct ArtError { int code; };
[[noreturn]] void artRaiseCuda (cudaError_t code, const char *file, int line, const char *expr);
[[noreturn]] void artRaise (const char *fmt, ...);

#define ART_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            artRaiseCuda (e_, __FILE__, __LINE__, #expr);                                     \
    } while (0)

/* every extern "C" entry point of the CUDA translation units is bracketed by these */
#define ART_GUARD_BEGIN try {
#define ART_GUARD_END(onError) } catch (const ArtError &) { return onError; } catch (const std::exception &e_) { artNote (e_.what ()); return onError; }
void artNote (const char *what);
