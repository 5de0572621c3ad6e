/*
 * art_sinc_generic.cu -- the any-ratio windowed-sinc kernel (sm_100a).
 *
 * Replaces the inner loop of resampleProcess / resampleProcessInterleaved
 * (resampler.c:523-526, :640-643), subsample_interpolate / subsample_no_interpolate
 * (:1135-1157), their _precise twins (:1159-1181) and apply_filter (:1033-1057).
 *
 * One CTA owns a tile of NB consecutive output frames of one job for a group of Cg
 * channels:
 *   1. every thread derives the position of its outputs from the closed form in
 *      art_plan.h (binary64, same operation order as the reference) and records
 *      (window start, filter row, interpolation weight);
 *   2. the tile's outputs are grouped by filter row (shared-memory counting sort) --
 *      a row pair is 2*T floats that no neighbouring output shares when the phase
 *      step is large, so grouping is what turns 3 KB of filter traffic per output
 *      into 3 KB per row per tile;
 *   3. the input window of the tile (history ++ input block) is staged in shared
 *      memory with coalesced loads, channel-interleaved in groups of CV so that one
 *      LDS.64 / LDS.128 brings a sample of every channel of the group;
 *   4. a warp takes a run of up to 8 outputs that share a row pair (run table built
 *      with the sort): lanes split the taps, each lane keeps 8 x CV x 2 accumulators
 *      in registers, the row pair is read once per run through L1, the samples come
 *      from shared memory (conflict free: 32 consecutive vectors per load, immediate
 *      offsets from one pointer per output); runs of <= 4, 2, 1 outputs use
 *      narrower register tiles instead of idle slots;
 *   5. the partial sums are combined with a transposing shuffle reduction (one
 *      shuffle per value instead of five) and written to global memory.
 *
 * Arithmetic: float FMA accumulation, float lerp of the two row sums (the reference
 * lerps in double after two float sums; the difference is below 1 ulp of the sums).
 * ART_MODE_PRECISE accumulates in double like apply_filter_precise.
 */
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "art_kernels.cuh"
#include "art_device.h"

#define ART_G_THREADS 256
#define ART_G_WARPS   (ART_G_THREADS / 32)
#define ART_RUN       8

#define KEY_PASS0(F) ((F) + 1)      /* pass-through of the centre sample (resampler.c:1141-1142) */
#define KEY_PASS1(F) ((F) + 2)      /* ... of the one after it (row index == F)                  */
#define NUM_KEYS(F)  ((F) + 3)

template <int CV> struct ArtVec;
#if ART_WIDE        /* double samples: one LDS.64 / LDS.128 brings one or two channels */
typedef double artweight_t;
template <> struct ArtVec<1> { typedef double  type; __device__ static double get (const double  &x, int)   { return x; } };
template <> struct ArtVec<2> { typedef double2 type; __device__ static double get (const double2 &x, int v) { return v ? x.y : x.x; } };
#else
typedef float artweight_t;
template <> struct ArtVec<1> { typedef float  type; __device__ static float get (const float  &x, int)   { return x; } };
template <> struct ArtVec<2> { typedef float2 type; __device__ static float get (const float2 &x, int v) { return v ? x.y : x.x; } };
template <> struct ArtVec<4> { typedef float4 type; __device__ static float get (const float4 &x, int v) { return v == 0 ? x.x : v == 1 ? x.y : v == 2 ? x.z : x.w; } };
#endif

struct ArtTileCtx {
    const artsample_t *xs;           // staged window, [group][Wp][CV]
    const int *srel;                 // region index of the first tap, per tile-local output
    const artweight_t *wgt;          // interpolation weight, per tile-local output
    const unsigned short *order;     // tile-local output indices grouped by filter row
    long long sFirst;                // region index staged at xs[.][0]
    unsigned int n0;                 // first output frame of the tile
    int c0, nc;                      // first channel of the CTA, channels it owns
    int Wp, Tp, half;
};

/* SLOTS outputs that share the row pair `kk`, all channel groups of the CTA. */
template <bool INTERP, typename AccT, int CV, int SLOTS>
__device__ __forceinline__ void art_run_tile (const ArtTileCtx &t, const ArtJob &job, const artsample_t *__restrict__ bank,
                                              int kk, int e0, int len, int lane)
{
    typedef typename ArtVec<CV>::type VecT;
    const artsample_t *__restrict__ rowA = bank + (size_t) kk * t.Tp + lane;
    const artsample_t *__restrict__ rowB = rowA + t.Tp;
    const int NI = t.Tp >> 5;

    int off[SLOTS];
    artweight_t f[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        const int i = t.order[e0 + min (j, len - 1)];            // idle slots shadow the last real entry
        off[j] = (int) ((long long) t.srel[i] - t.sFirst) + lane;
        f[j] = t.wgt[i];
    }

    for (int cg = 0; cg < t.nc; cg += CV) {
        AccT a0[SLOTS][CV], a1[SLOTS][CV];
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
#pragma unroll
            for (int v = 0; v < CV; ++v) { a0[j][v] = 0; a1[j][v] = 0; }

        const VecT *xp[SLOTS];
        const VecT *plane = reinterpret_cast<const VecT *> (t.xs) + (size_t) (cg / CV) * t.Wp;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            xp[j] = plane + off[j];

#pragma unroll 4
        for (int i = 0; i < NI; ++i) {
            const AccT ca = __ldg (rowA + 32 * i);
            const AccT cb = INTERP ? (AccT) __ldg (rowB + 32 * i) : (AccT) 0;
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                const VecT xv = xp[j][32 * i];
#pragma unroll
                for (int v = 0; v < CV; ++v) {
                    const AccT x = ArtVec<CV>::get (xv, v);
                    a0[j][v] = fma (ca, x, a0[j][v]);
                    if (INTERP) a1[j][v] = fma (cb, x, a1[j][v]);
                }
            }
        }

        AccT vals[SLOTS * CV];
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
#pragma unroll
            for (int v = 0; v < CV; ++v)
                vals[j * CV + v] = INTERP ? fma ((AccT) f[j], a1[j][v] - a0[j][v], a0[j][v]) : a0[j][v];

        const AccT total = art_transpose_reduce<SLOTS * CV, AccT> (vals, lane);
        constexpr int LPV = 32 / (SLOTS * CV);                   // lanes holding the same value
        const int q = lane / LPV, j = q / CV, v = q - j * CV;
        if ((lane % LPV) == 0 && j < len && cg + v < t.nc)
            *art_out_ptr (job, t.c0 + cg + v, (long long) t.n0 + t.order[e0 + j]) = (artsample_t) total;
    }
}

#if !ART_WIDE       /* the packed-FP32 forms belong to the float path */
/* float32 version of art_run_tile built on packed FFMA2.  Interpolated: one accumulator pair per
 * (output, channel) holds (row A sum, row B sum), the coefficient pair (A[k], B[k]) is packed once per
 * tap and the sample is a scalar operand -- half the FMA issue slots of the scalar form.  Not
 * interpolated: pairs run over adjacent channels instead. */
template <bool INTERP, int CV, int SLOTS>
__device__ __forceinline__ void art_run_tile_f32 (const ArtTileCtx &t, const ArtJob &job, const float *__restrict__ bank,
                                                  int kk, int e0, int len, int lane)
{
    typedef typename ArtVec<CV>::type VecT;
    constexpr int NP = INTERP ? CV : (CV >= 2 ? CV / 2 : 1);    // accumulator pairs per output
    const float *__restrict__ rowA = bank + (size_t) kk * t.Tp + lane;
    const float *__restrict__ rowB = rowA + t.Tp;
    const int NI = t.Tp >> 5;

    int off[SLOTS];
    float f[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        const int i = t.order[e0 + min (j, len - 1)];            // idle slots shadow the last real entry
        off[j] = (int) ((long long) t.srel[i] - t.sFirst) + lane;
        f[j] = t.wgt[i];
    }

    for (int cg = 0; cg < t.nc; cg += CV) {
        unsigned long long acc2[SLOTS][NP];
        float acc1[SLOTS];                                        // only for !INTERP && CV == 1
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            acc1[j] = 0.0f;
#pragma unroll
            for (int v = 0; v < NP; ++v) acc2[j][v] = 0ull;
        }

        const VecT *xp[SLOTS];
        const VecT *plane = reinterpret_cast<const VecT *> (t.xs) + (size_t) (cg / CV) * t.Wp;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j)
            xp[j] = plane + off[j];

#pragma unroll 4
        for (int i = 0; i < NI; ++i) {
            const float ca = __ldg (rowA + 32 * i);
            const float cb = INTERP ? __ldg (rowB + 32 * i) : ca;
            const unsigned long long c2 = art_pack2 (ca, cb);     // (A, B) when interpolating, (A, A) otherwise
#pragma unroll
            for (int j = 0; j < SLOTS; ++j) {
                const VecT xv = xp[j][32 * i];
                if (INTERP) {
#pragma unroll
                    for (int v = 0; v < CV; ++v) {
                        const float x = ArtVec<CV>::get (xv, v);
                        art_ffma2 (acc2[j][v], c2, art_pack2 (x, x));
                    }
                }
                else if (CV >= 2) {
#pragma unroll
                    for (int v = 0; v < CV; v += 2)
                        art_ffma2 (acc2[j][v / 2], art_pack2 (ArtVec<CV>::get (xv, v), ArtVec<CV>::get (xv, v + 1)), c2);
                }
                else
                    acc1[j] = fmaf (ca, ArtVec<CV>::get (xv, 0), acc1[j]);
            }
        }

        float vals[SLOTS * CV];
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            if (INTERP) {
#pragma unroll
                for (int v = 0; v < CV; ++v) {
                    float a0, a1;
                    art_unpack2 (acc2[j][v], a0, a1);
                    vals[j * CV + v] = fmaf (f[j], a1 - a0, a0);
                }
            }
            else if (CV >= 2) {
#pragma unroll
                for (int v = 0; v < CV; v += 2)
                    art_unpack2 (acc2[j][v / 2], vals[j * CV + v], vals[j * CV + v + 1]);
            }
            else
                vals[j * CV] = acc1[j];
        }

        const float total = art_transpose_reduce<SLOTS * CV, float> (vals, lane);
        constexpr int LPV = 32 / (SLOTS * CV);                   // lanes holding the same value
        const int q = lane / LPV, j = q / CV, v = q - j * CV;
        if ((lane % LPV) == 0 && j < len && cg + v < t.nc)
            *art_out_ptr (job, t.c0 + cg + v, (long long) t.n0 + t.order[e0 + j]) = total;
    }
}

/* Near-unity ratios (ASRC): runs of CONSECUTIVE outputs share one filter-row pair and their windows start at consecutive
 * samples, so output r of a run is sum_k h[k] x[s + r + k]: a plain FIR over a sliding window that can be register-blocked.
 * The tile is cut into mini-runs of up to four such outputs.  A warp takes eight of them: lane (q, g) works on mini-run g and
 * the q-th quarter of the taps, keeps a four-sample window of channel vectors in registers and loads ONE new vector per tap
 * for 4 outputs x CV channels x 2 rows of multiply-adds (the row-sorted form above loads one vector per output and tap and is
 * bound by shared-memory bandwidth at twice the FMA time).  FFMA2 pairs run over adjacent channels -- the vector load delivers
 * them as aligned register pairs and the coefficient is the scalar operand, so no instruction is spent on packing.  The staged
 * window is skewed by one slot every eight so that the eight lanes of a quarter, whose windows start four samples apart, fall
 * on different banks. */
#define ART_SKEW(p) ((p) + ((p) >> 3))
template <int CV> struct ArtPairs;
template <> struct ArtPairs<2> { typedef unsigned long long type;
    __device__ static unsigned long long get (const unsigned long long &x, int)   { return x; } };
template <> struct ArtPairs<4> { typedef ulonglong2 type;
    __device__ static unsigned long long get (const ulonglong2 &x, int p) { return p ? x.y : x.x; } };

template <int CV>
__device__ __forceinline__ void art_run_unity (const ArtTileCtx &t, const ArtJob &job, const float *__restrict__ bank,
                                               int kk, int e0, int len, int lane)
{
    const int q = lane >> 3;
    const int Tq = t.Tp >> 2;                                          // taps per quarter: a multiple of 8, rows are zero-padded to Tp
    const float *__restrict__ rowA = bank + (size_t) kk * t.Tp + q * Tq;
    const float *__restrict__ rowB = rowA + t.Tp;
    const int p0 = (int) ((long long) t.srel[e0] - t.sFirst) + q * Tq; // window position of (output 0 of the mini-run, first tap of the quarter)
    const float fOut = t.wgt[e0 + min (q, max (len, 1) - 1)];          // this lane finishes output q of its mini-run

    if constexpr (CV >= 2) {
        typedef typename ArtPairs<CV>::type PairT;
        constexpr int NP = CV / 2;
        for (int cg = 0; cg < t.nc; cg += CV) {
            unsigned long long accA[4][NP], accB[4][NP];               // per output: channel pairs of the row A / row B sums
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int p = 0; p < NP; ++p) { accA[r][p] = 0ull; accB[r][p] = 0ull; }
            const PairT *plane = reinterpret_cast<const PairT *> (t.xs) + (size_t) (cg / CV) * t.Wp;
            /* skewed position of p0 + 8 * kb + d = skew (p0) + 9 * kb + d + ((p0 % 8 + d) / 8): the lane-dependent part is a constant
             * per d, so the window walks with ONE add per load (base + offset) instead of re-deriving the skew for every tap */
            const PairT *wbase = plane + ART_SKEW (p0);
            const int r8 = p0 & 7;
            int woff[11];
#pragma unroll
            for (int d = 0; d < 11; ++d) woff[d] = d + ((r8 + d) >> 3);
            PairT w0 = wbase[woff[0]], w1 = wbase[woff[1]], w2 = wbase[woff[2]], w3;
#define ART_UNITY_TAP(ca, cb, x0, x1, x2, x3)                                                   \
            do {                                                                                \
                const unsigned long long ca2 = art_pack2 (ca, ca), cb2 = art_pack2 (cb, cb);    \
                _Pragma ("unroll")                                                              \
                for (int p = 0; p < NP; ++p) {                                                  \
                    art_ffma2 (accA[0][p], ArtPairs<CV>::get (x0, p), ca2);                     \
                    art_ffma2 (accB[0][p], ArtPairs<CV>::get (x0, p), cb2);                     \
                    art_ffma2 (accA[1][p], ArtPairs<CV>::get (x1, p), ca2);                     \
                    art_ffma2 (accB[1][p], ArtPairs<CV>::get (x1, p), cb2);                     \
                    art_ffma2 (accA[2][p], ArtPairs<CV>::get (x2, p), ca2);                     \
                    art_ffma2 (accB[2][p], ArtPairs<CV>::get (x2, p), cb2);                     \
                    art_ffma2 (accA[3][p], ArtPairs<CV>::get (x3, p), ca2);                     \
                    art_ffma2 (accB[3][p], ArtPairs<CV>::get (x3, p), cb2);                     \
                }                                                                               \
            } while (0)
            for (int k = 0; k < Tq; k += 8, wbase += 9) {              // the window rotates through the four registers: no moves
                const float4 a4 = __ldg (reinterpret_cast<const float4 *> (rowA + k));
                const float4 b4 = __ldg (reinterpret_cast<const float4 *> (rowB + k));
                const float4 a8 = __ldg (reinterpret_cast<const float4 *> (rowA + k + 4));
                const float4 b8 = __ldg (reinterpret_cast<const float4 *> (rowB + k + 4));
                w3 = wbase[woff[3]];  ART_UNITY_TAP (a4.x, b4.x, w0, w1, w2, w3);
                w0 = wbase[woff[4]];  ART_UNITY_TAP (a4.y, b4.y, w1, w2, w3, w0);
                w1 = wbase[woff[5]];  ART_UNITY_TAP (a4.z, b4.z, w2, w3, w0, w1);
                w2 = wbase[woff[6]];  ART_UNITY_TAP (a4.w, b4.w, w3, w0, w1, w2);
                w3 = wbase[woff[7]];  ART_UNITY_TAP (a8.x, b8.x, w0, w1, w2, w3);
                w0 = wbase[woff[8]];  ART_UNITY_TAP (a8.y, b8.y, w1, w2, w3, w0);
                w1 = wbase[woff[9]];  ART_UNITY_TAP (a8.z, b8.z, w2, w3, w0, w1);
                w2 = wbase[woff[10]]; ART_UNITY_TAP (a8.w, b8.w, w3, w0, w1, w2);
            }
#undef ART_UNITY_TAP
            // sum the four tap quarters; the halving exchanges leave output q (all channels, both rows) in lane (q, g)
            float val[4][2 * CV];                                      // [output][A ch0 .. A ch(CV-1), B ch0 ..]
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    art_unpack2 (accA[r][p], val[r][2 * p], val[r][2 * p + 1]);
                    art_unpack2 (accB[r][p], val[r][CV + 2 * p], val[r][CV + 2 * p + 1]);
                }
            float half2[2][2 * CV], fin[2 * CV];
            const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int e = 0; e < 2 * CV; ++e) {
                    const float send = up16 ? val[rr][e] : val[rr + 2][e], keep = up16 ? val[rr + 2][e] : val[rr][e];
                    half2[rr][e] = keep + __shfl_xor_sync (0xffffffffu, send, 16);
                }
#pragma unroll
            for (int e = 0; e < 2 * CV; ++e) {
                const float send = up8 ? half2[0][e] : half2[1][e], keep = up8 ? half2[1][e] : half2[0][e];
                fin[e] = keep + __shfl_xor_sync (0xffffffffu, send, 8);
            }
            if (q < len) {
                float o[CV];
#pragma unroll
                for (int v = 0; v < CV; ++v) o[v] = fmaf (fOut, fin[CV + v] - fin[v], fin[v]);
                // the CV channels of a frame of an interleaved block: one vector store
                float *dst = art_out_ptr (job, t.c0 + cg, (long long) t.n0 + e0 + q);
                if (!job.outPlanes && job.outCS == 1 && cg + CV <= t.nc && (reinterpret_cast<uintptr_t> (dst) % (sizeof (float) * CV)) == 0) {
                    if constexpr (CV == 4) *reinterpret_cast<float4 *> (dst) = make_float4 (o[0], o[1], o[2], o[3]);
                    else                   *reinterpret_cast<float2 *> (dst) = make_float2 (o[0], o[1]);
                }
                else {
#pragma unroll
                    for (int v = 0; v < CV; ++v)
                        if (cg + v < t.nc)
                            *art_out_ptr (job, t.c0 + cg + v, (long long) t.n0 + e0 + q) = o[v];
                }
            }
        }
    }
    else {
        // one channel: the pair is (row A, row B) of one output and the sample is the scalar operand
        for (int cg = 0; cg < t.nc; ++cg) {
            unsigned long long acc[4] = { 0ull, 0ull, 0ull, 0ull };
            const float *plane = t.xs + (size_t) cg * t.Wp;
            float w0 = plane[ART_SKEW (p0)], w1 = plane[ART_SKEW (p0 + 1)], w2 = plane[ART_SKEW (p0 + 2)], w3;
#define ART_UNITY_TAP(ca, cb, x0, x1, x2, x3)                                                   \
            do {                                                                                \
                const unsigned long long c2 = art_pack2 (ca, cb);                               \
                art_ffma2 (acc[0], c2, art_pack2 (x0, x0));                                     \
                art_ffma2 (acc[1], c2, art_pack2 (x1, x1));                                     \
                art_ffma2 (acc[2], c2, art_pack2 (x2, x2));                                     \
                art_ffma2 (acc[3], c2, art_pack2 (x3, x3));                                     \
            } while (0)
            for (int k = 0; k < Tq; k += 4) {
                const float4 a4 = __ldg (reinterpret_cast<const float4 *> (rowA + k));
                const float4 b4 = __ldg (reinterpret_cast<const float4 *> (rowB + k));
                w3 = plane[ART_SKEW (p0 + k + 3)]; ART_UNITY_TAP (a4.x, b4.x, w0, w1, w2, w3);
                w0 = plane[ART_SKEW (p0 + k + 4)]; ART_UNITY_TAP (a4.y, b4.y, w1, w2, w3, w0);
                w1 = plane[ART_SKEW (p0 + k + 5)]; ART_UNITY_TAP (a4.z, b4.z, w2, w3, w0, w1);
                w2 = plane[ART_SKEW (p0 + k + 6)]; ART_UNITY_TAP (a4.w, b4.w, w3, w0, w1, w2);
            }
#undef ART_UNITY_TAP
            float val[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) art_unpack2 (acc[r], val[r][0], val[r][1]);
            float half2[2][2], fin[2];
            const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float send = up16 ? val[rr][e] : val[rr + 2][e], keep = up16 ? val[rr + 2][e] : val[rr][e];
                    half2[rr][e] = keep + __shfl_xor_sync (0xffffffffu, send, 16);
                }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float send = up8 ? half2[0][e] : half2[1][e], keep = up8 ? half2[1][e] : half2[0][e];
                fin[e] = keep + __shfl_xor_sync (0xffffffffu, send, 8);
            }
            if (q < len)
                *art_out_ptr (job, t.c0 + cg, (long long) t.n0 + e0 + q) = fmaf (fOut, fin[1] - fin[0], fin[0]);
        }
    }
}

#endif

template <bool INTERP, bool PRECISE, int CV, int SLOTS>
__device__ __forceinline__ void art_run_any (const ArtTileCtx &t, const ArtJob &job, const artsample_t *__restrict__ bank,
                                             int kk, int e0, int len, int lane)
{
#if ART_WIDE
    art_run_tile<INTERP, double, CV, SLOTS> (t, job, bank, kk, e0, len, lane);
#else
    if (PRECISE) art_run_tile<INTERP, double, CV, SLOTS> (t, job, bank, kk, e0, len, lane);
    else         art_run_tile_f32<INTERP, CV, SLOTS> (t, job, bank, kk, e0, len, lane);
#endif
}

#ifndef ART_SKEW
#define ART_SKEW(p) (p)
#endif

/* UNITY: the near-unity form (art_run_unity) as a kernel of its own: it needs ~85 registers where the row-sorted form takes
 * 128, so -- for one or two channels per vector; four need the registers -- three CTAs fit an SM instead of two and the serial
 * phases of a tile (positions, staging, the mini-run scan) of one CTA overlap the multiply-adds of the others */
template <bool INTERP, bool PRECISE, int CV, bool UNITY>
__global__ void __launch_bounds__ (ART_G_THREADS, (UNITY && CV <= 2) ? 3 : 2)
art_sinc_generic_kernel (const ArtClass k, const __grid_constant__ ArtJob single, const ArtJob *__restrict__ jobs)
{
    extern __shared__ __align__ (16) unsigned char smem_raw[];
    const int nkeys = NUM_KEYS (k.F);
    const int nkeysPad = (nkeys + 2 + 3) & ~3;
    const int maxRuns = k.NB / ART_RUN + nkeys + 8;

    artsample_t *xs = reinterpret_cast<artsample_t *> (smem_raw);       // [Cg/CV][Wp][CV]
    artweight_t *wgt = reinterpret_cast<artweight_t *> (xs + (size_t) k.Cg * k.Wp);   // [NB] interpolation weight
    int *srel = reinterpret_cast<int *> (wgt + k.NB);                   // [NB] region index of the first tap
    int *binEnd = srel + k.NB;                                          // [nkeysPad] counts -> starts -> ends
    int *chunk0 = binEnd + nkeysPad;                                    // [nkeysPad] first run index per key
    unsigned int *runTab = reinterpret_cast<unsigned int *> (chunk0 + nkeysPad);      // [maxRuns] key<<20 | start<<4 | len-1
    unsigned short *key = reinterpret_cast<unsigned short *> (runTab + maxRuns);      // [NB]
    unsigned short *order = key + k.NB;                                 // [NB] tile-local output index, grouped by key

    __shared__ long long sh_first, sh_last;
    __shared__ double sh_base0;
    __shared__ int sh_w0;
    __shared__ int sh_scan[ART_G_WARPS];
    __shared__ int sh_scan2[ART_G_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // one job travels by value in the parameter space (no descriptor upload on the latency path);
    // batches and block sequences come as an array in global memory
    // channel group varies fastest across the grid, so the CTAs reading the same frames run together (L2)
    const int groups = (k.C + k.Cg - 1) / k.Cg;
    const int tileIdx = blockIdx.x / groups, cgroup = blockIdx.x - tileIdx * groups;
    const ArtJob &job = jobs ? jobs[k.numJobs > 1 ? art_find_job_warp (jobs, k.numJobs, tileIdx) : 0] : single;
    const unsigned int t0 = (unsigned int) (tileIdx - job.tile0) * (unsigned int) k.NB;
    if (t0 >= job.outputs)
        return;
    const int cnt = (int) min ((unsigned int) k.NB, job.outputs - t0);
    const unsigned int n0 = job.nStart + t0;                      // call-relative index of the tile's first output
    const int c0 = cgroup * k.Cg;
    const int nc = min (k.Cg, k.C - c0);
    const int T = k.T, Tref = k.Tref, half = Tref / 2 + k.lead, F = k.F;          // T taps per row; positions run on Tref

    ArtLoopState st;
    st.P = job.P; st.ratio = job.ratio; st.I = job.I; st.T = Tref;
    const long long D = 15LL * Tref;

    for (int i = tid; i < nkeysPad; i += ART_G_THREADS) {
        binEnd[i] = 0;
        chunk0[i] = 0;
    }
    if (tid == 0) {                         // the rounding chain up to the tile's first output, once
        int w;
        (void) art_output_pos (&st, n0, &w);
        sh_w0 = w;
        sh_base0 = art_ring_base (st.P, Tref, w);
    }
    __syncthreads ();
    const int w0 = sh_w0;
    const double base0 = sh_base0;

    /* ---- 1. positions ------------------------------------------------------------------ */
    for (int i = tid; i < cnt; i += ART_G_THREADS) {
        int w;
        const double pos = art_output_pos_from (&st, n0 + i, w0, base0, &w);
        const double whole = floor (pos);
        const double fr = pos - whole;
        // region index of the first tap: ring index (whole - half + 1), un-compacted, minus origin
        const long long s = (long long) whole - half + 1 + (long long) w * D - job.origin;
        int kk;
        artweight_t f = 0;
        if (INTERP) {
            double ph = fr * F;                                   // resampler.c:1149-1152
            int row = (int) floor (ph);
            ph -= row;
            if (row >= F) { row = F - 1; ph = 1.0; }             // fr*F rounded up to F: same point on the bank
            kk = row;
            f = (artweight_t) ph;
        }
        else {
            int row = (int) floor (fr * F + 0.5);                 // resampler.c:1137
            if (!(k.mode & ART_MODE_LOWPASS) && row % F == 0)     // resampler.c:1141-1142
                kk = row ? KEY_PASS1 (F) : KEY_PASS0 (F);
            else
                kk = row;
        }
        if (i == 0) sh_first = s;
        if (i == cnt - 1) sh_last = s;
        srel[i] = (int) s;                 // region index of the first tap (frame counts are ints)
        wgt[i] = f;
        key[i] = (unsigned short) kk;
        atomicAdd (&binEnd[kk], 1);
    }
    __syncthreads ();

    const long long sFirst = sh_first;
    constexpr bool unity = !ART_WIDE && INTERP && !PRECISE && UNITY;  // consecutive-output chunks instead of row-sorted runs (see art_run_unity)
    // samples of window the tile touches (the unity form reads up to 3 + 32 positions past a chunk's last window: zero taps, but staged)
    const int span = (int) (sh_last - sFirst) + k.Tp + (unity ? 40 : 0);
    if ((unity ? ART_SKEW (span) + 1 : span) > k.Wp) {
        if (tid == 0)
            printf ("libresampler_b200: tile window %d exceeds plane %d (ratio %g)\n", span, k.Wp, job.ratio);
        __trap ();
    }

    /* ---- 3. stage the window (issued before the scan so the loads overlap it) ------------ */
    {
        // xs[(group * Wp + j) * CV + v] = sample j of channel group*CV + v
        const bool interleavedSrc = (job.inPlanes == nullptr) && (job.inCS == 1);
        typedef typename ArtVec<CV>::type VecT;
        // whole channel vectors straight from an interleaved block: one LDG.64/128 and one STS.64/128 per CV samples
        const bool vectorSrc = CV > 1 && interleavedSrc && nc == k.Cg && (job.inFS % CV) == 0 &&
                               (reinterpret_cast<uintptr_t> (job.in + c0) % (sizeof (artsample_t) * CV)) == 0;
        if (vectorSrc) {
            const int Q = k.Cg / CV, total = span * Q;
            const long long lo = -job.prevAvail, hi = job.inValid;
            VecT *xv = reinterpret_cast<VecT *> (xs);
            for (int e = tid; e < total; e += ART_G_THREADS) {
                const int j = e / Q, cq = e - j * Q;
                const int js = unity ? ART_SKEW (j) : j;
                const long long idx = sFirst + j;
                VecT v;
                if (idx >= lo && idx < hi)
                    v = __ldg (reinterpret_cast<const VecT *> (job.in + idx * job.inFS + c0 + cq * CV));
                else {
                    alignas (16) artsample_t tmp[CV];
#pragma unroll
                    for (int u = 0; u < CV; ++u) tmp[u] = art_fetch (job, T, c0 + cq * CV + u, idx);
                    v = *reinterpret_cast<VecT *> (tmp);
                }
                xv[(size_t) cq * k.Wp + js] = v;
            }
        }
        else if (interleavedSrc) {
            const int total = span * k.Cg;
            for (int e = tid; e < total; e += ART_G_THREADS) {
                const int j = e / k.Cg, cc = e - j * k.Cg;
                const int js = unity ? ART_SKEW (j) : j;
                xs[((cc / CV) * k.Wp + js) * CV + (cc % CV)] = cc < nc ? art_fetch (job, T, c0 + cc, sFirst + j) : (artsample_t) 0;
            }
        }
        else {
            for (int cc = 0; cc < k.Cg; ++cc)
                for (int j = tid; j < span; j += ART_G_THREADS) {
                    const int js = unity ? ART_SKEW (j) : j;
                    xs[((cc / CV) * k.Wp + js) * CV + (cc % CV)] = cc < nc ? art_fetch (job, T, c0 + cc, sFirst + j) : (artsample_t) 0;
                }
        }
    }

    /* ---- 2u. near-unity ratios: cut the tile into mini-runs of <= 4 consecutive outputs with one row pair and consecutive
     *          windows, in natural order (no sort) -------------------------------------------------------------------------- */
#if !ART_WIDE
    if constexpr (unity) {
        const int per = (cnt + ART_G_THREADS - 1) / ART_G_THREADS;
        const int i0 = min (tid * per, cnt), i1 = min (i0 + per, cnt);
        auto boundary = [&] (int i) -> bool { return i == 0 || key[i] != key[i - 1] || srel[i] != srel[i - 1] + 1; };
        // the last boundary at or before every thread's range (block-wide running maximum)
        int lastB = -1;
        for (int i = i0; i < i1; ++i) if (boundary (i)) lastB = i;
        int inc = lastB;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync (0xffffffffu, inc, o);
            if (lane >= o) inc = max (inc, a);
        }
        if (lane == 31) sh_scan[warp] = inc;
        __syncthreads ();
        int carry = __shfl_up_sync (0xffffffffu, inc, 1);
        if (lane == 0) carry = -1;
        for (int w = 0; w < warp; ++w) carry = max (carry, sh_scan[w]);
        // mini-run starts: every boundary and every 4th output after it
        int nat = carry, mine = 0;
        for (int i = i0; i < i1; ++i) {
            if (boundary (i)) nat = i;
            mine += ((i - nat) & 3) == 0;
        }
        int incC = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync (0xffffffffu, incC, o);
            if (lane >= o) incC += a;
        }
        if (lane == 31) sh_scan2[warp] = incC;
        __syncthreads ();
        int at = incC - mine;
        for (int w = 0; w < warp; ++w) at += sh_scan2[w];
        nat = carry;
        for (int i = i0; i < i1; ++i) {
            if (boundary (i)) nat = i;
            if (((i - nat) & 3) == 0) order[at++] = (unsigned short) i;
        }
        int totalMini = 0;
        for (int w = 0; w < ART_G_WARPS; ++w) totalMini += sh_scan2[w];
        __syncthreads ();

        ArtTileCtx t;
        t.xs = xs; t.srel = srel; t.wgt = wgt; t.order = order;
        t.sFirst = sFirst; t.n0 = n0; t.c0 = c0; t.nc = nc; t.Wp = k.Wp; t.Tp = k.Tp; t.half = half;
        for (int m0 = warp * 8; m0 < totalMini; m0 += ART_G_WARPS * 8) {
            const int m = m0 + (lane & 7);                        // lane group g = lane & 7 takes mini-run m0 + g
            const bool live = m < totalMini;
            const int e0 = order[live ? m : totalMini - 1];
            const int e1 = m + 1 < totalMini ? order[m + 1] : cnt;
            art_run_unity<CV> (t, job, k.bank, key[e0], e0, live ? e1 - e0 : 0, lane);
        }
        return;
    }
#endif

    /* ---- 2. group by filter row ---------------------------------------------------------- */
    int totalRuns;
    {
        // exclusive scans of counts (-> starts) and of ceil(count / RUN) (-> first run), in one pass
        const int per = (nkeys + ART_G_THREADS - 1) / ART_G_THREADS;
        const int b0 = tid * per, b1 = min (b0 + per, nkeys);
        int sumC = 0, sumR = 0;
        for (int b = b0; b < b1; ++b) {
            sumC += binEnd[b];
            sumR += (binEnd[b] + ART_RUN - 1) / ART_RUN;
        }
        int incC = sumC, incR = sumR;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int a = __shfl_up_sync (0xffffffffu, incC, o);
            int r = __shfl_up_sync (0xffffffffu, incR, o);
            if (lane >= o) { incC += a; incR += r; }
        }
        if (lane == 31) { sh_scan[warp] = incC; sh_scan2[warp] = incR; }
        __syncthreads ();
        int baseC = 0, baseR = 0;
        for (int w = 0; w < warp; ++w) { baseC += sh_scan[w]; baseR += sh_scan2[w]; }
        int runC = baseC + incC - sumC, runR = baseR + incR - sumR;
        for (int b = b0; b < b1; ++b) {
            const int c = binEnd[b];
            binEnd[b] = runC;          // start of the bin; the scatter below advances it to the end
            chunk0[b] = runR;
            // the run table: every run knows its row, where its outputs sit in `order`, and how many
            for (int r = 0, left = c; left > 0; ++r, left -= ART_RUN)
                runTab[runR + r] = ((unsigned int) b << 20) | ((unsigned int) (runC + r * ART_RUN) << 4) |
                                   (unsigned int) (min (left, ART_RUN) - 1);
            runC += c;
            runR += (c + ART_RUN - 1) / ART_RUN;
        }
        int tr = 0;
        for (int w = 0; w < ART_G_WARPS; ++w) tr += sh_scan2[w];
        totalRuns = tr;
        __syncthreads ();
        for (int i = tid; i < cnt; i += ART_G_THREADS) {
            const int slot = atomicAdd (&binEnd[key[i]], 1);
            order[slot] = (unsigned short) i;
        }
    }
    __syncthreads ();

    /* ---- 4. convolve: one warp per run ---------------------------------------------------- */
    ArtTileCtx t;
    t.xs = xs; t.srel = srel; t.wgt = wgt; t.order = order;
    t.sFirst = sFirst; t.n0 = n0; t.c0 = c0; t.nc = nc; t.Wp = k.Wp; t.Tp = k.Tp; t.half = half;

    for (int run = warp; run < totalRuns; run += ART_G_WARPS) {
        const unsigned int packed = runTab[run];
        const int kk = (int) (packed >> 20), e0 = (int) ((packed >> 4) & 0xffff), len = (int) (packed & 15) + 1;

        if (kk > F) {
            /* pass-through: the stored sample itself */
            const int shift = (kk == KEY_PASS1 (F)) ? 1 : 0;
            for (int q = lane; q < len * nc; q += 32) {
                const int j = q / nc, cc = q - j * nc;
                const int i = order[e0 + j];
                const int at = (int) ((long long) srel[i] - sFirst) + half - 1 + shift;
                *art_out_ptr (job, c0 + cc, (long long) n0 + i) = xs[((cc / CV) * k.Wp + at) * CV + (cc % CV)];
            }
        }
        else if (len > 4) art_run_any<INTERP, PRECISE, CV, 8> (t, job, k.bank, kk, e0, len, lane);
        else if (len > 2) art_run_any<INTERP, PRECISE, CV, 4> (t, job, k.bank, kk, e0, len, lane);
        else if (len > 1) art_run_any<INTERP, PRECISE, CV, 2> (t, job, k.bank, kk, e0, len, lane);
        else              art_run_any<INTERP, PRECISE, CV, 1> (t, job, k.bank, kk, e0, len, lane);
    }
}

/* ---- history: the newest T consumed samples of every channel ------------------------------- */
__global__ void art_history_kernel (const ArtClass k, const __grid_constant__ ArtJob single, const ArtJob *__restrict__ jobs, int numJobs)
{
    const int total = k.C * k.T;
    for (int jb = blockIdx.y; jb < numJobs; jb += gridDim.y) {            // grid.y is capped at 65535
        const ArtJob &job = jobs ? jobs[jb] : single;
        if (!job.histOut)
            continue;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int c = e / k.T, i = e - c * k.T;
            job.histOut[e] = art_fetch (job, k.T, c, job.consumed - k.T + i);
        }
    }
}

/* ---- host side ----------------------------------------------------------------------------- */

static size_t generic_smem (const ArtClass &k)
{
    const int nkeys = NUM_KEYS (k.F);
    const int nkeysPad = (nkeys + 2 + 3) & ~3;
    const int maxRuns = k.NB / ART_RUN + nkeys + 8;
    return (size_t) k.Cg * k.Wp * sizeof (artsample_t) + (size_t) k.NB * (sizeof (artweight_t) + 4 + 2 + 2) + (size_t) nkeysPad * 8 + (size_t) maxRuns * 4 + 16;
}

static int plane_floats (int NB, double ratio, int Tp)
{
    // NB consecutive outputs span at most (NB-1)/ratio + 2 input frames; + the padded tap count
    double span = (double) (NB - 1) / ratio + 2.0 + Tp;
    if (span > 1.0e9) span = 1.0e9;
    return (((int) span) + 31) & ~31;
}

/* plane pitch of the near-unity form: the skewed layout (one slot in eight), its over-read of 40 positions, and a pitch of
 * 4 mod 8 vectors so that the two planes an 8-channel CTA stages side by side start on different banks */
static int unity_plane (int Wp)
{
    return ((((Wp + 40) * 9 / 8 + 40) + 7) & ~7) + 4;
}

void artPlanGenericGeometry (ArtClass &k, double minRatio, unsigned int maxOutputs, int smCount, ArtLaunchGeom &g)
{
    const size_t budget = (k.unity && k.C < 4 ? 74 : 110) * 1024;     // two CTAs per SM inside the 227 KB carve-out (three of the near-unity kernel at CV <= 2)
    const int C = k.C;
    int cv = C >= 4 ? 4 : (C >= 2 ? 2 : 1);
    if ((k.mode & ART_MODE_PRECISE) || ART_WIDE) cv = C >= 2 ? 2 : 1;
    const int maxCg = ((C < 8 ? C : 8) + cv - 1) / cv * cv;

    // small calls: shorter tiles so that the grid still covers the GPU
    int nbCap = 4096;
    while (nbCap > 256 &&
           (unsigned long long) nbCap * smCount * 2 > (unsigned long long) maxOutputs * ((C + maxCg - 1) / maxCg))
        nbCap >>= 1;

    // A row pair is fetched once per run: what amortises it is (outputs per row in a tile) x
    // (channels per CTA).  Maximise that, then the tile length.
    double bestScore = -1.0;
    int bestNB = 0, bestCg = 0;
    for (int NB = nbCap; NB >= 1; NB >>= 1)
        for (int Cg = maxCg; Cg >= cv; Cg -= cv) {
            ArtClass t = k;
            t.NB = NB; t.Cg = Cg; t.Wp = plane_floats (NB, minRatio, k.Tp);
            if (k.unity) t.Wp = unity_plane (t.Wp);
            if (t.Wp >= (1 << 26) || generic_smem (t) > budget)
                continue;
            double perRow = (double) NB / (k.F + 1);
            if (perRow < 1.0) perRow = 1.0;
            if (perRow > ART_RUN) perRow = ART_RUN + (perRow - ART_RUN) * 0.05;   // beyond a full run only L1 traffic improves
            const double score = perRow * Cg + 1e-6 * NB;
            if (score > bestScore) { bestScore = score; bestNB = NB; bestCg = Cg; }
        }
    if (!bestNB) {
        artRaise ("ratio %g needs a %d-float window per output; unsupported", minRatio, plane_floats (1, minRatio, k.Tp));
    }
    k.NB = bestNB;
    k.Cg = bestCg;
    k.Wp = plane_floats (bestNB, minRatio, k.Tp);
    if (k.unity) k.Wp = unity_plane (k.Wp);
    g.CV = cv;
    g.smemBytes = generic_smem (k);
}

template <bool INTERP, bool PRECISE, int CV, bool UNITY = false>
static void launch_one (const ArtClass &k, const ArtLaunchGeom &g, const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    auto kern = art_sinc_generic_kernel<INTERP, PRECISE, CV, UNITY>;
    static size_t configured[16] = { 0 };      // per device: the opt-in is per (function, context)
    int device = 0;
    ART_CUDA_CHECK (cudaGetDevice (&device));
    if (g.smemBytes > configured[device & 15]) {
        // the planner keeps tiles below 110 KB (two CTAs per SM); opt in to exactly that much
        ART_CUDA_CHECK (cudaFuncSetAttribute (kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        configured[device & 15] = 112 * 1024;
    }
    const unsigned int grid = (unsigned int) g.totalTiles * (unsigned int) ((k.C + k.Cg - 1) / k.Cg);
    void *prof;
    artProfileBegin (stream, &prof);
    kern<<<grid, ART_G_THREADS, g.smemBytes, stream>>> (k, single, d_jobs);
    artProfileEnd (stream, prof);
    ART_CUDA_CHECK (cudaGetLastError ());
    ++g_artLaunches;
}

void artLaunchGeneric (const ArtClass &k, const ArtLaunchGeom &g, const ArtJob &single, const ArtJob *d_jobs, cudaStream_t stream)
{
    if (g.totalTiles <= 0)
        return;
    const bool interp = k.mode & ART_MODE_INTERP, precise = k.mode & ART_MODE_PRECISE;
#define ART_DISPATCH(CVV)                                                                 \
    do {                                                                                  \
        if (interp && !precise && k.unity) launch_one<true, false, CVV, true> (k, g, single, d_jobs, stream); \
        else if (interp && !precise)  launch_one<true, false, CVV> (k, g, single, d_jobs, stream); \
        else if (!interp && !precise) launch_one<false, false, CVV> (k, g, single, d_jobs, stream); \
        else if (interp)              launch_one<true, true, CVV> (k, g, single, d_jobs, stream);  \
        else                          launch_one<false, true, CVV> (k, g, single, d_jobs, stream); \
    } while (0)
#if ART_WIDE        /* double samples, double accumulation: the "precise" instances are the only ones */
    if (g.CV == 2) { if (interp) launch_one<true, true, 2> (k, g, single, d_jobs, stream); else launch_one<false, true, 2> (k, g, single, d_jobs, stream); }
    else           { if (interp) launch_one<true, true, 1> (k, g, single, d_jobs, stream); else launch_one<false, true, 1> (k, g, single, d_jobs, stream); }
    (void) precise;
#else
    if (g.CV == 4) ART_DISPATCH (4);
    else if (g.CV == 2) ART_DISPATCH (2);
    else ART_DISPATCH (1);
#endif
#undef ART_DISPATCH
}

void artLaunchHistory (const ArtClass &k, const ArtJob &single, const ArtJob *d_jobs, int numJobs, cudaStream_t stream)
{
    const int total = k.C * k.T;
    if (!d_jobs) numJobs = 1;
    dim3 grid ((total + 255) / 256 > 64 ? 64 : (total + 255) / 256, numJobs < 65535 ? numJobs : 65535);
    art_history_kernel<<<grid, 256, 0, stream>>> (k, single, d_jobs, numJobs);
    ART_CUDA_CHECK (cudaGetLastError ());
    ++g_artLaunches;
}
