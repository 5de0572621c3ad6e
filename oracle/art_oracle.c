/*
 * art_oracle.c -- TEST INFRASTRUCTURE ONLY (see art_oracle.h).
 *
 * Scalar CPU restatement of the audio-resampler hot path.  The arithmetic
 * (operation order, float/double types, the ring-compaction bookkeeping that
 * decides how positions are rounded) follows the reference exactly so that
 * input_used / output_generated / position are bit-identical to it; the
 * code structure is this project's own (one strided core, contiguous bank
 * and ring storage).  Build with -O2 -ffp-contract=off: the reference's
 * -fassociative-math only changes the float summation order, which parity
 * tolerates.
 */
#include "art_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ bank */

/* One windowed-sinc row centred `fraction` of a sample past tap taps/2-1.
 * Restates init_filter, resampler.c:1090-1133. */
static void build_row (const OracleResampler *r, osample_t *row, double *scratch, double fraction)
{
    static const double bh[4] = { 0.35875, 0.48829, 0.14128, 0.01168 };   /* resampler.c:1093-1096 */
    const int taps = r->taps, half = taps / 2;
    double total = 0.0;

    for (int t = 0; t < taps; ++t) {
        double dist = fabs ((half - 1) + fraction - t) * M_PI;             /* :1106 */
        double wpos = dist / half;                                          /* :1107 */
        double v = 1.0;

        if (dist != 0.0) {                                                  /* :1110-1117 */
            v = sin (dist * r->lowpass_ratio) / (dist * r->lowpass_ratio);
            if (r->flags & ORC_BLACKMAN_HARRIS)
                v *= bh[0] + bh[1] * cos (wpos) + bh[2] * cos (2 * wpos) + bh[3] * cos (3 * wpos);
            else
                v *= 0.5 * (1.0 + cos (wpos));
        }

        scratch[t] = v;
        total += v;                                                         /* :1121 */
    }

    /* unity DC gain, then round to float from the centre outwards while
     * carrying the rounding error to the next tap visited (:1126-1132).
     * Visiting order: half, half-1, half+1, half-2, ..., taps-1, 0. */
    const double gain = 1.0 / total;
    double carried = 0.0;

    for (int k = 0; k < half; ++k) {
        const int order[2] = { half + k, half - 1 - k };
        for (int j = 0; j < 2; ++j) {
            int t = order[j];
            scratch[t] *= gain;
            row[t] = (osample_t) (scratch[t] - carried);
            carried += row[t] - scratch[t];
        }
    }
}

/* resampleInit, resampler.c:115-199 */
OracleResampler *oracle_init (int channels, int taps, int phases, double lowpass_ratio, int flags)
{
    if (lowpass_ratio > 0.0 && lowpass_ratio < 1.0)                         /* :120-125 */
        flags |= ORC_LOWPASS;
    else {
        flags &= ~ORC_LOWPASS;
        lowpass_ratio = 1.0;
    }

    if ((taps & 3) || taps <= 0 || taps > 1024)                             /* :127-130 */
        return NULL;
    if (phases < 1 || phases > 1024)                                        /* :132-135 */
        return NULL;

    OracleResampler *r = calloc (1, sizeof *r);
    r->channels = channels;
    r->taps = taps;
    r->phases = phases;
    r->flags = flags;
    r->lowpass_ratio = lowpass_ratio;
    r->ring_len = 16 * taps;                                                /* :139 */
    r->bank = calloc ((size_t) (phases + 1) * taps, sizeof (osample_t));
    r->ring = calloc ((size_t) (channels > 0 ? channels : 1) * r->ring_len, sizeof (osample_t));

    double *scratch = malloc (sizeof (double) * taps);
    for (int p = 0; p < phases; ++p)                                        /* :149-155 */
        build_row (r, r->bank + (size_t) p * taps, scratch, (double) p / phases);
    free (scratch);

    /* the extra row is row 0 delayed by one tap (:156-159) ... */
    osample_t *last = r->bank + (size_t) phases * taps;
    for (int t = 0; t < taps; ++t)
        last[(t + 1) % taps] = r->bank[t];
    /* ... and the two window outliers are cleared (:167-168) */
    r->bank[taps - 1] = 0.0f;
    last[0] = 0.0f;

    r->read_pos = taps / 2;                                                 /* :176-177 */
    r->write_index = taps;
    if (r->flags & ORC_EXTRAPOLATE) r->flags |= ORC_PREFILL_PENDING;                /* :179-182 */
    return r;
}

static unsigned long common_divisor (unsigned long a, unsigned long b)     /* resampler.c:999-1008 */
{
    while (b) { unsigned long t = a % b; a = b; b = t; }
    return a;
}

/* resampleFixedRatioInit, resampler.c:310-356 */
OracleResampler *oracle_fixed_ratio_init (int channels, int taps, int max_phases, double src_rate,
                                          double dst_rate, int lowpass_hz, int flags)
{
    double lowpass = lowpass_hz / (dst_rate / 2.0);                         /* :312 */
    double ratio = dst_rate / src_rate;                                     /* :313 */

    if (lowpass_hz > dst_rate / 2.0)                                        /* :316-319 */
        return NULL;

    if (src_rate == floor (src_rate) && dst_rate == floor (dst_rate) && !(flags & ORC_NO_REDUCTION)) {
        unsigned long need = (unsigned long) dst_rate /                     /* :324 */
                             common_divisor ((unsigned long) src_rate, (unsigned long) dst_rate);
        if (need <= (unsigned long) max_phases) {                           /* :326-334 */
            flags &= ~ORC_INTERPOLATE;
            max_phases = (int) need;
            if (max_phases & (max_phases - 1))
                flags |= ORC_SNAP;
        }
    }

    if (!lowpass_hz && (flags & ORC_LOWPASS) && dst_rate < src_rate) {      /* :340-348 */
        lowpass = 1.0 - (7.5 / taps / ratio);
        if (lowpass < 0.8) lowpass = 0.8;
        if (lowpass < ratio) lowpass = ratio;
    }

    OracleResampler *r = oracle_init (channels, taps, max_phases, lowpass * ratio, flags | ORC_FIXED_RATIO);
    if (r)
        r->fixed_ratio = dst_rate / src_rate;                               /* :353 */
    return r;
}

void oracle_free (OracleResampler *r)
{
    if (!r) return;
    free (r->bank);
    free (r->ring);
    free (r);
}

/* resampleReset, resampler.c:383-397 */
void oracle_reset (OracleResampler *r)
{
    memset (r->ring, 0, sizeof (osample_t) * (size_t) r->channels * r->ring_len);
    r->read_pos = r->taps / 2;
    r->write_index = r->taps;
    r->flags &= ~ORC_FLUSHED;
    if (r->flags & ORC_EXTRAPOLATE) r->flags |= ORC_PREFILL_PENDING;                /* :393-394 */
}

/* resampleAdvancePosition, resampler.c:927-935 */
void oracle_advance (OracleResampler *r, double delta)
{
    if (delta < 0.0)
        return;
    if (!(r->flags & ORC_INTERPOLATE) && floor (delta) != delta)
        return;
    r->read_pos += delta;
}

/* resampleGetPosition, resampler.c:965-968 */
double oracle_position (const OracleResampler *r)
{
    return r->read_pos + (r->taps / 2.0) - r->write_index;
}

const osample_t *oracle_bank_row (const OracleResampler *r, int row)
{
    return r->bank + (size_t) row * r->taps;
}

/* ------------------------------------------------------------ convolution */

/* apply_filter "version 2", resampler.c:1033-1044: float accumulator, pairs
 * taken from both ends towards the middle. */
static double dot_outside_in (const osample_t *coef, const osample_t *x, int taps)
{
    osample_t acc = 0.0f;
    for (int lo = 0, hi = taps - 1; lo < hi; ++lo, --hi)
        acc += (coef[lo] * x[lo]) + (coef[hi] * x[hi]);
    return acc;
}

/* apply_filter_precise, resampler.c:1049-1057 */
static double dot_double (const osample_t *coef, const osample_t *x, int taps)
{
    double acc = 0.0;
    for (int t = 0; t < taps; ++t)
        acc += (double) coef[t] * x[t];
    return acc;
}

/* subsample_interpolate / subsample_no_interpolate and their _precise twins,
 * resampler.c:1135-1181.  `line` is one channel's ring, `where` the read
 * position in ring coordinates. */
static double sample_at (const OracleResampler *r, const osample_t *line, double where)
{
    double (*dot) (const osample_t *, const osample_t *, int) =
        (r->flags & ORC_EXTENDED_MATH) ? dot_double : dot_outside_in;      /* :191-196 */
    const int taps = r->taps;
    const double whole = floor (where);

    if (r->flags & ORC_INTERPOLATE) {
        double frac = (where - whole) * r->phases;                          /* :1149-1150 */
        int row = (int) floor (frac);
        frac -= row;                                                        /* :1152 */
        const osample_t *x = line + (int) whole - taps / 2 + 1;                 /* :1153 */
        return dot (r->bank + (size_t) row * taps, x, taps) * (1.0 - frac) +
               dot (r->bank + (size_t) (row + 1) * taps, x, taps) * frac;   /* :1155-1156 */
    }

    int row = (int) floor ((where - whole) * r->phases + 0.5);              /* :1137 */
    const osample_t *centre = line + (int) whole;
    if (!(r->flags & ORC_LOWPASS) && row % r->phases == 0)                  /* :1141-1142 */
        return centre[row / r->phases];
    return dot (r->bank + (size_t) row * taps, centre - taps / 2 + 1, taps);   /* :1144 */
}

/* ------------------------------------------------------------- streaming */


/* ---- endpoint extrapolation (extrapolator.c) ---------------------------------------------------------
 * A 4-coefficient linear predictor is fitted to the known samples by coordinate descent with a halving
 * step (extrapolator.c:93-170), made stable through its reflection coefficients (:174-194, :244-283),
 * replaced by "repeat the last sample" or by silence when those predict better (:213-222), and then run
 * forward to synthesise the missing samples (extrapolator.c:22-43). */
#define ORC_LPC_ORDER 4
#define ORC_LPC_MAX_ROUNDS 100000

static double predictor_output (const float *coef, const osample_t *window)        /* sum_c coef[N-1-c] * window[c] */
{
    double acc = 0.0;
    for (int c = 0; c < ORC_LPC_ORDER; ++c)
        acc += coef[ORC_LPC_ORDER - c - 1] * window[c];
    return acc;
}

static void to_reflection (const double *lpc, double *refl)                     /* extrapolator.c:244-272 */
{
    double cur[ORC_LPC_ORDER], nxt[ORC_LPC_ORDER];
    memcpy (cur, lpc, sizeof cur);
    for (int m = ORC_LPC_ORDER - 1; m >= 0; --m) {
        refl[m] = cur[m];
        double den = 1.0 - refl[m] * refl[m];
        if (fabs (den) < 1e-6) {
            refl[m] = refl[m] < 0.0 ? -0.9999995 : 0.9999995;
            den = 1.0 - refl[m] * refl[m];
        }
        for (int i = 0; i < m; ++i)
            nxt[i] = (cur[i] - refl[m] * cur[m - i - 1]) / den;
        for (int i = 0; i < m; ++i)
            cur[i] = nxt[i];
    }
}

static void from_reflection (const double *refl, double *lpc)                   /* extrapolator.c:276-283 */
{
    for (int i = 0; i < ORC_LPC_ORDER; ++i) {
        lpc[i] = refl[i];
        for (int j = 0; j < i / 2; ++j) {
            double keep = lpc[j];
            lpc[j] += refl[i] * lpc[i - 1 - j];
            lpc[i - 1 - j] += refl[i] * keep;
        }
        if (i & 1)
            lpc[i >> 1] += lpc[i >> 1] * refl[i];
    }
}

static void fit_predictor (const osample_t *x, int count, float *coef)              /* extrapolator.c:93-240 */
{
    const int evals = count - ORC_LPC_ORDER;
    double signal_energy = 0.0, delta_energy = 0.0, best, step = 3.0 / 16.0;
    double *resid = malloc (sizeof (double) * (evals > 0 ? evals : 1));
    int rounds = 0, moved = 0;

    memset (coef, 0, sizeof (float) * ORC_LPC_ORDER);
    for (int i = 0; i < evals; ++i) {                                            /* :101-107 */
        const osample_t now = x[i + ORC_LPC_ORDER], before = x[i + ORC_LPC_ORDER - 1];
        delta_energy += (now - before) * (now - before);
        signal_energy += now * now;
    }
    if (signal_energy == 0.0) { free (resid); return; }                          /* :109-112 */
    best = signal_energy;

    while (best > 0.0 && rounds < ORC_LPC_MAX_ROUNDS) {                          /* :118 */
        int which;
        for (int k = 0; k < evals; ++k)                                          /* :121-129 */
            resid[k] = predictor_output (coef, x + k) + x[k + ORC_LPC_ORDER];
        for (which = 0; rounds++, which < ORC_LPC_ORDER; which++) {              /* :131 */
            double down = 0.0, up = 0.0;
            for (int k = 0; k < evals; ++k) {
                const double nudge = x[k + ORC_LPC_ORDER - which - 1] * step;
                down += (resid[k] - nudge) * (resid[k] - nudge);
                up += (resid[k] + nudge) * (resid[k] + nudge);
            }
            if (down < best || up < best) {                                      /* :141-153 */
                if (down < up) { best = down; coef[which] -= step; }
                else           { best = up;   coef[which] += step; }
                moved++;
                break;
            }
        }
        if (which == ORC_LPC_ORDER) {                                            /* :158-163 */
            if (step > 3.0 / (1 << 22)) step *= 0.5;
            else break;
        }
    }
    free (resid);

    if (moved) {                                                                 /* :170-194 */
        double wide[ORC_LPC_ORDER], refl[ORC_LPC_ORDER];
        int clipped = 0;
        for (int i = 0; i < ORC_LPC_ORDER; ++i) wide[i] = coef[i];
        to_reflection (wide, refl);
        for (int i = 0; i < ORC_LPC_ORDER; ++i)
            if (fabs (refl[i]) > 0.9999) { refl[i] = refl[i] < 0.0 ? -0.9999 : 0.9999; clipped++; }
        if (clipped) {
            from_reflection (refl, wide);
            for (int i = 0; i < ORC_LPC_ORDER; ++i) coef[i] = wide[i];
        }
    }

    double err = 0.0;                                                            /* :198-209 */
    for (int k = 0; k < evals; ++k) {
        const double e = predictor_output (coef, x + k) + x[k + ORC_LPC_ORDER];
        err += e * e;
    }
    if (delta_energy < err && delta_energy < signal_energy) {                    /* :213-222 */
        memset (coef, 0, sizeof (float) * ORC_LPC_ORDER);
        coef[0] = -1.0f;
    }
    else if (signal_energy <= err)
        memset (coef, 0, sizeof (float) * ORC_LPC_ORDER);
}

/* extrapolate_forward, extrapolator.c:22-43: x[0..known) are given, x[known..known+more) are written */
void oracle_extend_forward (osample_t *x, int known, int more)
{
    float coef[ORC_LPC_ORDER];
    memset (x + known, 0, sizeof (osample_t) * more);
    fit_predictor (x, known, coef);
    for (int i = 0; i < more; ++i)
        x[known + i] = (osample_t) -predictor_output (coef, x + known - ORC_LPC_ORDER + i);
}

/* extrapolate_reverse, extrapolator.c:49-65: end[-1], end[-2] ... end[-known] are given (newest first when read
 * backwards), end[-known-1] ... end[-known-more] are written */
void oracle_extend_backward (osample_t *end, int known, int more)
{
    osample_t *flip = calloc ((size_t) known + more, sizeof (osample_t));
    for (int i = 0; i < known; ++i) flip[i] = end[-1 - i];
    oracle_extend_forward (flip, known, more);
    for (int i = known; i < known + more; ++i) end[-1 - i] = flip[i];
    free (flip);
}

/* Ring compaction: keep the newest `taps` samples, shift both cursors
 * (resampler.c:497-503 / :614-620 / :667-673). */
static void compact_ring (OracleResampler *r)
{
    const int drop = r->ring_len - r->taps;
    for (int c = 0; c < r->channels; ++c) {
        osample_t *line = r->ring + (size_t) c * r->ring_len;
        memmove (line, line + drop, sizeof (osample_t) * r->taps);
    }
    r->read_pos -= drop;
    r->write_index -= drop;
}

/* postfillAllChannels, resampler.c:663-685 */
static void append_silence (OracleResampler *r)
{
    const int half = r->taps / 2;
    if (r->ring_len - r->write_index < half)
        compact_ring (r);
    for (int c = 0; c < r->channels; ++c) {
        osample_t *line = r->ring + (size_t) c * r->ring_len;
        memset (line + r->write_index, 0, sizeof (osample_t) * (r->ring_len - r->write_index));
        if (r->flags & ORC_EXTRAPOLATE)                                     /* :677-680 */
            oracle_extend_forward (line + r->write_index - half, half, half);
    }
    r->flags |= ORC_FLUSHED;
    r->write_index += half;
}

/* The control loop shared by resampleProcess (resampler.c:433-541) and
 * resampleProcessInterleaved (:550-658).  in_base[c]/out_base[c] point at
 * frame 0 of channel c; consecutive frames are *_stride floats apart. */
static OracleResult run (OracleResampler *r, const osample_t *const *in_base, int in_stride, int n_in,
                         osample_t *const *out_base, int out_stride, int n_out, double ratio)
{
    OracleResult res = { 0, 0 };
    const int half = r->taps / 2;
    double step = 0.0;                                   /* "offset2" */

    if (r->flags & ORC_FIXED_RATIO) ratio = r->fixed_ratio;                 /* :435-436 */
    if (r->flags & ORC_FLUSHED) n_in = 0;                                   /* :438-439 */
    if (n_in < 0) append_silence (r);                                       /* :491-492 */

    while (n_out > 0) {
        if (r->read_pos + step >= r->write_index - half) {                  /* :495 */
            if (n_in <= 0)
                break;
            if (r->write_index == r->ring_len)                              /* :497 */
                compact_ring (r);
            for (int c = 0; c < r->channels; ++c)
                r->ring[(size_t) c * r->ring_len + r->write_index] =
                    in_base[c][(size_t) res.input_used * in_stride];
            r->write_index++;
            res.input_used++;
            n_in--;
        }
        else {
            if (r->flags & ORC_PREFILL_PENDING) {                                   /* :516-522, prefillAllChannels :691-698 */
                const int have = r->write_index - r->taps;
                r->flags &= ~ORC_PREFILL_PENDING;
                if (have >= 8)
                    for (int c = 0; c < r->channels; ++c)
                        oracle_extend_backward (r->ring + (size_t) c * r->ring_len + r->write_index, have, r->taps - have);
            }
            for (int c = 0; c < r->channels; ++c)
                out_base[c][(size_t) res.output_generated * out_stride] =
                    (osample_t) sample_at (r, r->ring + (size_t) c * r->ring_len, r->read_pos + step);
            step = ++res.output_generated / ratio;                          /* :526 */
            n_out--;
        }
    }

    r->read_pos += step;                                                    /* :531 */
    if (r->flags & ORC_SNAP) {                                              /* :533-535 */
        double whole = floor (r->read_pos);
        r->read_pos = whole + floor ((r->read_pos - whole) * r->phases + 0.5) / r->phases;
    }
    return res;
}

OracleResult oracle_process_interleaved (OracleResampler *r, const osample_t *in, int n_in,
                                         osample_t *out, int n_out, double ratio)
{
    const int C = r->channels;
    const osample_t **ib = malloc (sizeof *ib * (C > 0 ? C : 1));
    osample_t **ob = malloc (sizeof *ob * (C > 0 ? C : 1));
    for (int c = 0; c < C; ++c) { ib[c] = in ? in + c : NULL; ob[c] = out + c; }
    OracleResult res = run (r, ib, C, n_in, ob, C, n_out, ratio);
    free (ib); free (ob);
    return res;
}

OracleResult oracle_process_planar (OracleResampler *r, const osample_t *const *in, int n_in,
                                    osample_t *const *out, int n_out, double ratio)
{
    const int C = r->channels;
    const osample_t **ib = malloc (sizeof *ib * (C > 0 ? C : 1));
    for (int c = 0; c < C; ++c) ib[c] = in ? in[c] : NULL;
    OracleResult res = run (r, ib, 1, n_in, out, 1, n_out, ratio);
    free (ib);
    return res;
}

/* resampleProcessAndFlushInterleaved, resampler.c:741-758 */
OracleResult oracle_process_flush_interleaved (OracleResampler *r, const osample_t *in, int n_in,
                                               osample_t *out, int n_out, double ratio)
{
    OracleResult res = oracle_process_interleaved (r, in, n_in, out, n_out, ratio);
    if ((n_in -= res.input_used) != 0 || (n_out -= res.output_generated) == 0)
        return res;
    OracleResult tail = oracle_process_interleaved (r, NULL, -1,
        out + (size_t) res.output_generated * r->channels, n_out, ratio);
    res.output_generated += tail.output_generated;
    return res;
}

/* resampleProcessAndFlush, resampler.c:712-739 */
OracleResult oracle_process_flush_planar (OracleResampler *r, const osample_t *const *in, int n_in,
                                          osample_t *const *out, int n_out, double ratio)
{
    OracleResult res = oracle_process_planar (r, in, n_in, out, n_out, ratio);
    if ((n_in -= res.input_used) != 0 || (n_out -= res.output_generated) == 0)
        return res;
    osample_t **shifted = malloc (sizeof *shifted * r->channels);
    for (int c = 0; c < r->channels; ++c) shifted[c] = out[c] + res.output_generated;
    OracleResult tail = oracle_process_planar (r, NULL, -1, shifted, n_out, ratio);
    free (shifted);
    res.output_generated += tail.output_generated;
    return res;
}

/* resampleGetRequiredSamples, resampler.c:853-880 (accumulates 1/ratio) */
unsigned int oracle_required_input (const OracleResampler *r, int n_out, double ratio)
{
    const int half = r->taps / 2, drop = r->ring_len - r->taps;
    int wi = r->write_index;
    double pos = r->read_pos;
    unsigned int used = 0;

    if (r->flags & ORC_FIXED_RATIO) ratio = r->fixed_ratio;
    while (n_out > 0) {
        if (pos >= wi - half) {
            if (wi == r->ring_len) { pos -= drop; wi -= drop; }
            wi++; used++;
        }
        else { pos += 1.0 / ratio; n_out--; }
    }
    return used;
}

/* resampleGetExpectedOutput, resampler.c:882-918 */
unsigned int oracle_expected_output (const OracleResampler *r, int n_in, double ratio)
{
    const int half = r->taps / 2, drop = r->ring_len - r->taps;
    int wi = r->write_index;
    double pos = r->read_pos;
    unsigned int made = 0;

    if (r->flags & ORC_FIXED_RATIO) ratio = r->fixed_ratio;
    if (r->flags & ORC_FLUSHED) n_in = 0;
    else if (n_in < 0) wi += half;

    for (;;) {
        if (pos >= wi - half) {
            if (n_in <= 0) break;
            if (wi == r->ring_len) { pos -= drop; wi -= drop; }
            wi++; n_in--;
        }
        else { pos += 1.0 / ratio; made++; }
    }
    return made;
}

/* ----------------------------------------------------------------- biquad */

static void biquad_common (OracleBiquadCoeffs *c, double frequency, int highpass)
{
    /* biquad.c:18-30 and :34-46 */
    double q = sqrt (0.5), k = tan (M_PI * frequency);
    double norm = 1.0 / (1.0 + k / q + k * k);
    memset (c, 0, sizeof *c);
    if (highpass) {
        c->a0 = norm;
        c->a1 = -2.0 * norm;
    }
    else {
        c->a0 = k * k * norm;
        c->a1 = 2 * c->a0;            /* note: doubles the already-rounded float a0 */
    }
    c->a2 = c->a0;
    c->b1 = 2.0 * (k * k - 1.0) * norm;
    c->b2 = (1.0 - k / q + k * k) * norm;
}

void oracle_biquad_lowpass (OracleBiquadCoeffs *c, double f)  { biquad_common (c, f, 0); }
void oracle_biquad_highpass (OracleBiquadCoeffs *c, double f) { biquad_common (c, f, 1); }

/* biquad_init, biquad.c:51-74 */
void oracle_biquad_init (OracleBiquad *q, const OracleBiquadCoeffs *c, double gain)
{
    const osample_t fwd[5] = { c->a0, c->a1, c->a2, c->a3, c->a4 };
    const osample_t bwd[5] = { 0.0f, c->b1, c->b2, c->b3, c->b4 };
    memset (q, 0, sizeof *q);
    for (int i = 0; i < 5; ++i) {
        q->a[i] = fwd[i] * gain;
        q->b[i] = bwd[i];
    }
    q->order = 1;
    for (int i = 2; i <= 4; ++i)
        if (fwd[i] != 0.0f || bwd[i] != 0.0f)
            q->order = i;
}

/* biquad_apply_buffer, biquad.c:106-163: direct form I in float, the newest
 * history entry sits at cursor & 3.  The sum is formed left to right exactly
 * as the reference's expression is written. */
void oracle_biquad_run (OracleBiquad *q, osample_t *buf, int count, int stride)
{
    int cur = q->cursor;
    while (count--) {
        osample_t acc = *buf * q->a[0];
        for (int d = 1; d <= q->order; ++d) {
            int slot = (cur - (d - 1)) & 3;
            acc = acc + (q->xh[slot] * q->a[d]) - (q->b[d] * q->yh[slot]);
        }
        ++cur;
        q->xh[cur & 3] = *buf;
        *buf = q->yh[cur & 3] = acc;
        buf += stride;
    }
    q->cursor = cur;
}

/* ------------------------------------------------------------------ noise */

/* fill_buffer_with_noise, artest.c:744-754 */
void oracle_noise (unsigned long long *state, osample_t *dst, int count)
{
    unsigned long long s = *state;
    while (count--) {
        for (int k = 0; k < 3; ++k)
            s = ((s << 4) - s) ^ 1;
        *dst++ = (float) ((int) (s >> 32) / 4294967296.0);
    }
    *state = s;
}

/* ------------------------------------------------------- float <-> integer stage */
/* decimator.c: what sits either side of the resampling path in art.c (:996 floatIntegersLE in front,
 * :1066 decimateProcessInterleavedLE behind).  One strided per-channel core serves the planar and the
 * interleaved entry points; channel state lives in one struct per channel. */

/* biquad_apply_sample, biquad.c:78-102 */
static osample_t shaper_step (OracleBiquad *q, osample_t in)
{
    osample_t acc = in * q->a[0];
    int cur = q->cursor & 3;
    for (int d = q->order; d >= 1; --d) {
        int slot = (cur - (d - 1)) & 3;
        acc += (q->xh[slot] * q->a[d]) - (q->b[d] * q->yh[slot]);
    }
    q->cursor = cur = (cur + 1) & 3;
    q->xh[cur] = in;
    q->yh[cur] = acc;
    return acc;
}

/* shaper_init, decimator.c:383-402: N(z) -> the decoupled H(z) form */
static void shaper_from_nz (OracleBiquad *q, const double n[9])
{
    OracleBiquadCoeffs c;
    memset (&c, 0, sizeof c);
    c.a0 = n[5] - n[1]; c.a1 = n[6] - n[2]; c.a2 = n[7] - n[3]; c.a3 = n[8] - n[4];
    c.b1 = n[5]; c.b2 = n[6]; c.b3 = n[7]; c.b4 = n[8];
    oracle_biquad_init (q, &c, 1.0);
}

/* the generator step of tpdf_dither / decimateInit's seeding, decimator.c:47-49, :361-365 */
static unsigned int lcg15 (unsigned int x) { return ((x << 4) - x) ^ 1u; }

OracleDecimator *oracle_decimate_init (int channels, int bits, int bytes, double gain, int rate, int flags)
{
    static const double ath[5][9] = {            /* decimator.c:68-77, Gesemann's ATH curves */
        { 1.0, -0.780459, +0.569358, -0.348221, +0.466316, +0.950797, +0.282052, +0.004337, +1.76209e-5 },
        { 1.0, -1.1474, 0.5383, -0.3530, 0.3475, 1.0587, 0.0676, -0.6054, -0.2738 },
        { 1.0, -1.3344, 0.7455, -0.4602, 0.4363, 0.9030, 0.0116, -0.5853, -0.2571 },
        { 1.0, -2.150679, +2.1402057, -1.042712, +0.206838, +0.67433, +1.017047, +0.4028633, +0.098656 },
        { 1.0, -2.16994, +2.01986, -0.894857, +0.1557738, +0.517789, +1.1062189, +0.4825786, +0.244994 } };
    static const int ath_rate[5] = { 32000, 44100, 48000, 88200, 96000 };
    static const double order1[9] = { 1, -1, 0, 0, 0, 0, 0, 0, 0 }, order2[9] = { 1, -2, 1, 0, 0, 0, 0, 0, 0 },
                        order3[9] = { 1, -3, 3, -1, 0, 0, 0, 0, 0 };
    OracleDecimator *d = calloc (1, sizeof *d);
    d->channels = channels; d->bits = bits; d->bytes = bytes; d->gain = gain; d->flags = flags;
    d->lane = calloc (channels, sizeof *d->lane);
    if (flags & 0x7) {                            /* DITHER_ENABLED: per-channel seeds are consecutive BYTES of one generator (:41-52) */
        unsigned int r = 0x31415926u;
        for (int c = 0; c < channels; ++c) {
            unsigned int seed = 0;
            for (int b = 0; b < 4; ++b) {
                seed |= (r >> 24) << (8 * b);     /* little-endian assembly of the byte stream */
                r = lcg15 (lcg15 (lcg15 (r)));
            }
            d->lane[c].rng = seed;
        }
        d->dither = (flags & 0x1) ? -1 : ((flags & 0x4) ? 1 : 0);       /* highpass, lowpass, flat (:54-59) */
    }
    if (flags & 0xf00)                            /* SHAPING_ENABLED (:62-90) */
        for (int c = 0; c < channels; ++c) {
            const double *nz = order1;
            if (flags & 0x800) {
                for (int k = 0; k < 5; ++k) if (rate == ath_rate[k]) nz = ath[k];
            }
            else if (flags & 0x100) nz = order1;
            else if (flags & 0x200) nz = order2;
            else if (flags & 0x400) nz = order3;
            shaper_from_nz (&d->lane[c].shaper, nz);
        }
    return d;
}

void oracle_decimate_free (OracleDecimator *d)
{
    if (d) { free (d->lane); free (d); }
}

/* tpdf_dither, decimator.c:361-373 */
static double tpdf (unsigned int *gen, int type)
{
    unsigned int r = lcg15 (lcg15 (*gen));
    unsigned int first = type ? (*gen ^ (unsigned int) (type >> 31)) : ~r;
    r = lcg15 (lcg15 (lcg15 (r)));
    *gen = r;
    return (((first >> 1) + (r >> 1)) / 2147483648.0) - 1.0;
}

/* one channel: decimator.c:170-199 (= :243-272, :301-333) */
static int decimate_channel (OracleDecimator *d, int c, const osample_t *in, int in_stride, int frames, unsigned char *out, int out_stride_bytes)
{
    const osample_t scaler = (1 << d->bits) / 2.0 * d->gain;
    const int pad = d->bytes - ((d->bits + 7) / 8);
    const int bias = (d->bits <= 8) * 128, top = (1 << (d->bits - 1)) - 1, bottom = ~top, shl = (24 - d->bits) % 8;
    int clips = 0;
    for (int i = 0; i < frames; ++i, in += in_stride, out += out_stride_bytes) {
        const osample_t dith = (d->flags & 0x7) ? tpdf (&d->lane[c].rng, d->dither) : 0.0;
        const osample_t code = (*in * scaler) - d->lane[c].feedback;
        int v = floor (code + dith + 0.5);
        unsigned char *o = out;
        if (d->flags & 0xf00)
            d->lane[c].feedback = shaper_step (&d->lane[c].shaper, v - code);
        if (v > top) { v = top; ++clips; }
        else if (v < bottom) { v = bottom; ++clips; }
        for (int j = 0; j < pad; ++j) *o++ = 0;
        v = ((unsigned int) v << shl) + bias;
        *o++ = v;
        if (d->bits > 8) { *o++ = v >> 8; if (d->bits > 16) *o++ = v >> 16; }
    }
    return clips;
}

int oracle_decimate_interleaved (OracleDecimator *d, const osample_t *in, int frames, unsigned char *out)
{
    int clips = 0;
    for (int c = 0; c < d->channels; ++c)
        clips += decimate_channel (d, c, in + c, d->channels, frames, out + c * d->bytes, d->channels * d->bytes);
    return clips;
}

int oracle_decimate_planar (OracleDecimator *d, const osample_t *const *in, int frames, unsigned char *const *out)
{
    int clips = 0;
    for (int c = 0; c < d->channels; ++c)
        clips += decimate_channel (d, c, in[c], 1, frames, out[c], d->bytes);
    return clips;
}

/* floatIntegersLE, decimator.c:416-450 */
void oracle_float_integers (const unsigned char *in, double gain, int bits, int bytes, int stride, osample_t *out, int count)
{
    const int used = (bits + 7) / 8;
    const osample_t g = bits <= 8 ? gain / 128.0 : (bits <= 16 ? gain / 32768.0 : gain / 8388608.0);
    in += bytes - used;
    for (int i = 0; i < count; ++i, in += (size_t) stride * bytes) {
        int v;
        if (bits <= 8) v = (int) in[0] - 128;
        else if (bits <= 16) v = (short) (in[0] | (in[1] << 8));
        else if (bits <= 24) v = in[0] | (in[1] << 8) | ((int) (signed char) in[2] << 16);
        else continue;                            /* the reference does nothing above 24 bits */
        *out++ = v * g;
    }
}
