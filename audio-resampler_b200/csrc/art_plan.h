/*
 * art_plan.h -- closed-form restatement of the reference's streaming control loop.
 *
 * The reference decides, one frame at a time, whether to pull an input frame into
 * its 16*T ring or to emit an output frame (resampler.c:494-529 / :611-646 /
 * :793-825).  Output n of a call is evaluated at
 *
 *        pos_n = (P - w*D) + (double) n / ratio            (resampler.c:526,643,822)
 *
 * where P is outputOffset at call entry, D = 15*T is what a ring compaction
 * subtracts (resampler.c:501-502) and w is how many compactions happened before
 * the output was emitted.  The reference subtracts D from outputOffset once per
 * compaction and only adds n/ratio back at the end of the call (:531), so inside
 * a long downsampling call "P - w*D" runs far negative and the individual
 * subtractions round; art_ring_base() reproduces that chain of roundings without
 * iterating w times.  The roundings decide counts, so they are kept exactly.
 *
 * Everything is a pure function of (P, I, T, ratio, n): the host uses it to get
 * input_used / output_generated / the new state in O(log N) per call, and every
 * CUDA thread uses the same inline code to find its own output position.  Shared
 * between C (host) and CUDA (device): only floor() and integer bit tricks.
 */
#ifndef ART_PLAN_H
#define ART_PLAN_H

#include <math.h>

#ifdef __CUDACC__
#define ART_HD __host__ __device__ __forceinline__
#else
#define ART_HD static inline
#endif

typedef struct {
    double P;          /* outputOffset at loop entry (after any flush adjustment)     */
    double ratio;      /* effective ratio (fixedRatio already substituted)            */
    int    I;          /* inputIndex at loop entry (after any flush adjustment)       */
    int    T;          /* numTaps                                                     */
} ArtLoopState;

/* "offset2" after n outputs: starts at literal 0.0 and becomes n / ratio afterwards
 * (resampler.c:489,526) -- the distinction matters for ratio == 0 */
ART_HD double art_step (unsigned int n, double ratio)
{
    return n ? (double) n / ratio : 0.0;
}

#define ART_NEVER (1LL << 62)

/* 2^(53 + e) where 2^e is the weight of the lowest set mantissa bit of x (x != 0): every
 * multiple of 2^e below this magnitude is exactly representable in binary64. */
ART_HD double art_exact_bound (double x)
{
    union { double d; unsigned long long u; } v;
    unsigned long long m;
    int E, tz;
    v.d = x;
    E = (int) ((v.u >> 52) & 0x7ff);
    m = v.u & 0xfffffffffffffULL;
    if (E) m |= 1ULL << 52; else E = 1;                /* subnormal */
#ifdef __CUDA_ARCH__
    tz = __ffsll ((long long) m) - 1;                   /* m != 0 because x != 0 */
#else
    tz = __builtin_ctzll (m);
#endif
    v.u = (unsigned long long) (E - 1075 + tz + 53 + 1023) << 52;    /* 2^(53+e), e = E-1075+tz */
    return v.d;
}

/* outputOffset after w ring compactions: fl(fl(fl(P - D) - D) ... - D), w times
 * (resampler.c:501, :618, :798).  P - k*D is exact while it stays below the bound above,
 * so whole runs of exact subtractions are taken in one step and only the (at most ~50)
 * subtractions that actually round are performed one by one. */
ART_HD double art_ring_base (double P, int T, long long w)
{
    const double D = 15.0 * T;
    double x = P;
    while (w > 0) {
        double room;
        long long k;
        if (x == 0.0)
            return x - (double) w * D;                  /* integers below 2^53: exact */
        room = floor ((x + art_exact_bound (x)) / D) - 2.0;     /* conservative count of exact steps */
        k = room < 1.0 ? 0 : (room > (double) w ? w : (long long) room);
        if (k > 0) {
            x -= (double) k * D;                        /* exact by construction */
            w -= k;
        }
        else {
            x -= D;                                     /* may round, exactly as the reference's does */
            w -= 1;
        }
    }
    return x;
}

/* number of ring compactions performed while pulling the first u inputs of the call
 * (compaction happens right before a pull that finds inputIndex == 16*T,
 * resampler.c:497 / :614 / :796) */
ART_HD int art_wraps (int I, long long u, int T)
{
    const long long NS = 16LL * T, D = 15LL * T;
    long long over = (long long) I + u - NS;
    return over <= 0 ? 0 : (int) ((over + D - 1) / D);
}

/* would the loop emit output n once u inputs have been pulled?  (negation of the
 * test at resampler.c:495 / :612 / :794) */
ART_HD int art_can_emit (const ArtLoopState *s, unsigned int n, long long u)
{
    const long long D = 15LL * s->T;
    const int w = art_wraps (s->I, u, s->T);
    const double pos = art_ring_base (s->P, s->T, w) + art_step (n, s->ratio);
    const double have = (double) ((long long) s->I + u - w * D - s->T / 2);
    return pos < have;
}

/* inputs that must have been pulled in this call before output n can be emitted */
ART_HD long long art_inputs_before (const ArtLoopState *s, unsigned int n)
{
    const double t = s->P + art_step (n, s->ratio);
    long long u;
    if (!(t < 4.0e18))                      /* inf / NaN position: never reachable */
        return ART_NEVER;
    u = (long long) floor (t) + s->T / 2 - s->I - 2;
    if (u < 0) u = 0;
    while (!art_can_emit (s, n, u))
        ++u;
    return u;
}

/* inputs pulled before output n, fast: away from an integer boundary the count follows from
 * floor(P + n/ratio) alone (the roundings of the chain move the sum by < 1e-9); next to one
 * the exact predicate decides */
ART_HD long long art_inputs_before_fast (const ArtLoopState *s, unsigned int n)
{
    const double t = s->P + art_step (n, s->ratio);
    const double k = floor (t);
    const double fr = t - k;
    long long u;
    if (fr > 1e-6 && fr < 1.0 - 1e-6 && t < 4.0e18) {
        u = (long long) k + s->T / 2 - s->I + 1;
        return u < 0 ? 0 : u;
    }
    return art_inputs_before (s, n);
}

/* position of output n in ring coordinates and the compaction count it is rounded under */
ART_HD double art_output_pos (const ArtLoopState *s, unsigned int n, int *wraps_out)
{
    const int w = art_wraps (s->I, art_inputs_before_fast (s, n), s->T);
    *wraps_out = w;
    return art_ring_base (s->P, s->T, w) + art_step (n, s->ratio);
}

/* same, given the compaction count w0 and ring base of an earlier output of the call (the
 * first output of a tile): the chain is extended by the few single subtractions that
 * separate the two instead of being rebuilt */
ART_HD double art_output_pos_from (const ArtLoopState *s, unsigned int n, int w0, double base0, int *wraps_out)
{
    const int w = art_wraps (s->I, art_inputs_before_fast (s, n), s->T);
    const double D = 15.0 * s->T;
    double x = base0;
    int k;
    *wraps_out = w;
    if (w < w0 || w - w0 > 64)
        return art_ring_base (s->P, s->T, w) + art_step (n, s->ratio);
    for (k = w0; k < w; ++k)
        x -= D;
    return x + art_step (n, s->ratio);
}

typedef struct {
    unsigned int outputs;      /* output_generated                                    */
    unsigned int inputs;       /* input_used                                          */
    double       P_after;      /* outputOffset on return, before any snap             */
    int          I_after;      /* inputIndex on return                                */
} ArtLoopPlan;

/* The whole loop in closed form: how many outputs fit, how many inputs they need. */
ART_HD ArtLoopPlan art_plan_loop (const ArtLoopState *s, int numIn, int numOut)
{
    ArtLoopPlan p;
    const long long D = 15LL * s->T;
    unsigned int lo = 0, hi = numOut > 0 ? (unsigned int) numOut : 0;
    if (numIn < 0) numIn = 0;

    /* largest N <= numOut with N == 0 or inputs_before(N-1) <= numIn (monotone in N).  The answer is
     * within a couple of units of (I + numIn - T/2 - P) * ratio, so bracket it from that estimate with
     * doubling steps and bisect only inside the bracket: ~4 predicate evaluations instead of ~log2(numOut). */
#define ART_OK(N) ((N) == 0 || art_inputs_before_fast (s, (N) - 1) <= numIn)
    if (hi > 0 && ART_OK (hi))
        lo = hi;
    else if (hi > 0) {
        double est = ((double) s->I + (double) numIn - (double) (s->T / 2) - s->P) * s->ratio;
        unsigned int g, stepw = 2;
        if (!(est > 0.0)) est = 0.0;
        if (est > (double) hi) est = (double) hi;
        g = (unsigned int) est;
        if (ART_OK (g)) {
            /* invariant: ok(lo), !ok(hi) */
            lo = g;
            while (lo + stepw < hi && ART_OK (lo + stepw)) { lo += stepw; stepw <<= 1; }
            if (lo + stepw < hi) hi = lo + stepw;
        }
        else {
            hi = g;
            while (hi > stepw && !ART_OK (hi - stepw)) { hi -= stepw; stepw <<= 1; }
            lo = hi > stepw ? hi - stepw : 0;
        }
        while (hi - lo > 1) {
            unsigned int mid = lo + (hi - lo) / 2;
            if (ART_OK (mid)) lo = mid; else hi = mid;
        }
    }
#undef ART_OK

    p.outputs = lo;
    if (numOut <= 0)
        p.inputs = 0;
    else if (lo == (unsigned int) numOut)
        p.inputs = (unsigned int) art_inputs_before_fast (s, lo - 1);   /* output space ran out */
    else
        p.inputs = (unsigned int) numIn;                            /* input ran out        */

    {
        const int w = art_wraps (s->I, p.inputs, s->T);
        const double step = art_step (lo, s->ratio);      /* resampler.c:526/:531 */
        p.P_after = art_ring_base (s->P, s->T, w) + step;
        p.I_after = (int) ((long long) s->I + p.inputs - w * D);
    }
    return p;
}

#endif
