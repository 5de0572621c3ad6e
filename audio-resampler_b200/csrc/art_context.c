/*
 * art_context.c -- host side (plain C) of libresampler_b200.so: the reference's public API
 * (include/resampler.h, include/resampler_b200.h), filter-bank design, and the scalar
 * streaming state.  Samples never pass through this file's arithmetic: every sample is
 * produced by the CUDA kernels behind art_device.h.
 */
#include "../../include/resampler_b200.h"
#include "art_device.h"
#include "art_extrapolate.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------ filter bank */

/* Row `fraction` of the bank: a sinc centred fraction of a sample past tap T/2-1, scaled by
 * the lowpass ratio, under a 4-term Blackman-Harris or Hann window, normalised to unity DC
 * gain and rounded to float with the rounding error carried centre-outwards.
 * Follows init_filter, reference resampler.c:1090-1133 (window constants :1093-1096). */
static void design_row (artsample_t *row, double *work, int taps, double fraction, double lowpass, int blackmanHarris)
{
    const int half = taps / 2;
    double sum = 0.0, scale, carry = 0.0;
    int t, k;

    for (t = 0; t < taps; ++t) {
        const double dist = fabs ((half - 1) + fraction - t) * M_PI;
        const double wang = dist / half;
        double v = 1.0;

        if (dist != 0.0) {
            v = sin (dist * lowpass) / (dist * lowpass);
            v *= blackmanHarris
                 ? 0.35875 + 0.48829 * cos (wang) + 0.14128 * cos (2 * wang) + 0.01168 * cos (3 * wang)
                 : 0.5 * (1.0 + cos (wang));
        }
        sum += work[t] = v;
    }

    scale = 1.0 / sum;
    for (k = 0; k < half; ++k) {            /* visit half, half-1, half+1, half-2, ... taps-1, 0 */
        int pass;
        for (pass = 0; pass < 2; ++pass) {
            t = pass ? half - 1 - k : half + k;
            work[t] *= scale;
            row[t] = (artsample_t) (work[t] - carry);
            carry += row[t] - work[t];
        }
    }
}

static artsample_t **design_bank (int taps, int filters, double lowpass, int flags)
{
    artsample_t **rows = calloc ((size_t) filters + 1, sizeof *rows);
    double *work = malloc (sizeof (double) * taps);
    int r, t;

    for (r = 0; r <= filters; ++r)
        rows[r] = calloc (taps, sizeof (artsample_t));
    for (r = 0; r < filters; ++r)                                   /* resampler.c:149-155 */
        design_row (rows[r], work, taps, (double) r / filters, lowpass, flags & BLACKMAN_HARRIS);
    for (t = 0; t < taps; ++t)                                      /* resampler.c:156-159 */
        rows[filters][(t + 1) % taps] = rows[0][t];
    rows[0][taps - 1] = 0;                                       /* resampler.c:167-168 */
    rows[filters][0] = 0.0f;
    free (work);
    return rows;
}

/* ------------------------------------------------------------------------------ init */

static int device_mode (int flags)
{
    int mode = 0;
    if (flags & SUBSAMPLE_INTERPOLATE)   mode |= ART_MODE_INTERP;
    if (flags & INCLUDE_LOWPASS)         mode |= ART_MODE_LOWPASS;
    if (flags & EXTEND_CONVOLUTION_MATH) mode |= ART_MODE_PRECISE;   /* resampler.c:191-196 */
    return mode;
}

Resample *resampleInit (int numChannels, int numTaps, int numFilters, double lowpassRatio, int flags)
{
    Resample *cxt;
    int mode = 0;

    if (lowpassRatio > 0.0 && lowpassRatio < 1.0)                   /* resampler.c:120-125 */
        flags |= INCLUDE_LOWPASS;
    else {
        flags &= ~INCLUDE_LOWPASS;
        lowpassRatio = 1.0;
    }

    if ((numTaps & 3) || numTaps <= 0 || numTaps > 1024) {          /* resampler.c:127-130 */
        fprintf (stderr, "must 4-1024 filter taps, and a multiple of 4!\n");
        return NULL;
    }
    if (numFilters < 1 || numFilters > 1024) {                      /* resampler.c:132-135 */
        fprintf (stderr, "must be 1-1024 filters!\n");
        return NULL;
    }
    if (numChannels < 1) {
        fprintf (stderr, "must be at least 1 channel!\n");
        return NULL;
    }
    flags &= ~EXTRAPOLATE_PREFILL;
    if (flags & EXTRAPOLATE_ENDPOINTS)                              /* resampler.c:179-182 */
        flags |= EXTRAPOLATE_PREFILL;

    cxt = calloc (1, sizeof *cxt);
    cxt->lowpassRatio = lowpassRatio;
    cxt->numChannels = numChannels;
    cxt->numSamples = numTaps * 16;                                 /* kept for struct compatibility (:139) */
    cxt->numFilters = numFilters;
    cxt->numTaps = numTaps;
    cxt->flags = flags;
    cxt->filters = design_bank (numTaps, numFilters, lowpassRatio, flags);
    cxt->outputOffset = numTaps / 2;                                /* resampler.c:176-177 */
    cxt->inputIndex = numTaps;

    mode = device_mode (flags);
    cxt->device = artDevCreate (numChannels, numTaps, 0, numFilters, mode, (const artsample_t *const *) cxt->filters);
    if (!cxt->device) {
        resampleFree (cxt);
        return NULL;
    }
    return cxt;
}

static unsigned long gcd_ul (unsigned long a, unsigned long b)     /* resampler.c:999-1008 */
{
    while (b) { unsigned long r = a % b; a = b; b = r; }
    return a;
}

Resample *resampleFixedRatioInit (int numChannels, int numTaps, int maxFilters, double sourceRate, double destinRate, int lowpassFreq, int flags)
{
    double lowpassRatio = lowpassFreq / (destinRate / 2.0);         /* resampler.c:312-313 */
    const double resampleRatio = destinRate / sourceRate;
    Resample *cxt;

    if (lowpassFreq > destinRate / 2.0) {                           /* resampler.c:316-319 */
        fprintf (stderr, "lowpass frequency must be lower than destination Nyquist!\n");
        return NULL;
    }

    /* integer rates whose reduced numerator fits the filter budget need no interpolation (:323-335) */
    if (sourceRate == floor (sourceRate) && destinRate == floor (destinRate) && !(flags & NO_FILTER_REDUCTION)) {
        const unsigned long exact = (unsigned long) destinRate / gcd_ul ((unsigned long) sourceRate, (unsigned long) destinRate);
        if (exact <= (unsigned long) maxFilters) {
            flags &= ~SUBSAMPLE_INTERPOLATE;
            maxFilters = (int) exact;
            if (maxFilters & (maxFilters - 1))
                flags |= RESAMPLER_SNAP_OFFSET;
        }
    }

    /* automatic lowpass for downsampling (:340-348) */
    if (!lowpassFreq && (flags & INCLUDE_LOWPASS) && destinRate < sourceRate) {
        lowpassRatio = 1.0 - (7.5 / numTaps / resampleRatio);
        if (lowpassRatio < 0.8) lowpassRatio = 0.8;
        if (lowpassRatio < resampleRatio) lowpassRatio = resampleRatio;
    }

    cxt = resampleInit (numChannels, numTaps, maxFilters, lowpassRatio * resampleRatio, flags | RESAMPLE_FIXED_RATIO);
    if (cxt)
        cxt->fixedRatio = destinRate / sourceRate;                  /* resampler.c:353 */
    return cxt;
}

double resampleGetLowpassRatio (Resample *cxt) { return cxt->lowpassRatio; }             /* resampler.c:365 */
int resampleGetNumFilters (Resample *cxt) { return cxt->numFilters; }                    /* resampler.c:370 */
int resampleInterpolationUsed (Resample *cxt) { return cxt->flags & SUBSAMPLE_INTERPOLATE; }   /* :375 */

void resampleReset (Resample *cxt)                                  /* resampler.c:383-397 */
{
    artDevReset (cxt->device);
    cxt->outputOffset = cxt->numTaps / 2;
    cxt->inputIndex = cxt->numTaps;
    cxt->flags &= ~RESAMPLER_FLUSHED;
    if (cxt->flags & EXTRAPOLATE_ENDPOINTS)                         /* resampler.c:393-394 */
        cxt->flags |= EXTRAPOLATE_PREFILL;
}

void resampleFree (Resample *cxt)                                   /* resampler.c:973-995 */
{
    int r;
    if (!cxt)
        return;
    if (cxt->device)
        artDevDestroy (cxt->device);
    if (cxt->plainDevice)
        artDevDestroy (cxt->plainDevice);
    free (cxt->prefilterTaps);
    if (cxt->filters) {
        for (r = 0; r <= cxt->numFilters; ++r)
            free (cxt->filters[r]);
        free (cxt->filters);
    }
    free (cxt);
}

/* ------------------------------------------------------------------- fused pre-filter */

/* include/resampler_b200.h: fold a cascade of biquad sections (biquad.c:51-74 coefficients, :106-163 recursion) into the bank */
int resampleB200AttachPrefilter (Resample *cxt, const Biquad *sections, int numSections)
{
    enum { MAXLEN = 1024 };
    double *g, *tmp, total = 0.0, tail;
    artsample_t **rows;
    ArtDev *fresh;
    int s, n, d, r, m, k, len, lead, T, Tk, mode = ART_MODE_LOWPASS;        /* never the pass-through shortcut (resampler.c:1141-1142) */

    if (!cxt || !sections || numSections < 1 || numSections > 16)
        return -1;
    T = cxt->numTaps;
    if (cxt->prefilterLead) {
        fprintf (stderr, "libresampler_b200: a pre-filter is already attached\n");
        return -1;
    }
    if (cxt->flags & EXTRAPOLATE_ENDPOINTS) {
        fprintf (stderr, "libresampler_b200: a fused pre-filter cannot be combined with EXTRAPOLATE_ENDPOINTS (the reference extrapolates the filtered signal)\n");
        return -1;
    }
    if (cxt->inputIndex != T || (cxt->flags & RESAMPLER_FLUSHED)) {
        fprintf (stderr, "libresampler_b200: attach the pre-filter before the first input (or right after resampleReset)\n");
        return -1;
    }
    for (s = 0; s < numSections; ++s)
        for (d = 0; d < 4; ++d)
            if (sections[s].x[d] != 0.0f || sections[s].y[d] != 0.0f) {
                fprintf (stderr, "libresampler_b200: pre-filter sections must be in their initial (zero) state\n");
                return -1;
            }

    /* impulse response of the cascade, in double, from the float coefficients the sections hold */
    g = calloc (MAXLEN, sizeof *g);
    tmp = calloc (MAXLEN, sizeof *tmp);
    g[0] = 1.0;
    for (s = 0; s < numSections; ++s) {
        const Biquad *q = &sections[s];
        memcpy (tmp, g, sizeof *g * MAXLEN);
        for (n = 0; n < MAXLEN; ++n) {
            double y = (double) q->a[0] * tmp[n];
            for (d = 1; d <= q->order && d <= 4; ++d)
                if (n >= d)
                    y += (double) q->a[d] * tmp[n - d] - (double) q->b[d] * g[n - d];
            g[n] = y;
        }
    }
    for (n = 0; n < MAXLEN; ++n) total += fabs (g[n]);
    /* shortest length (a multiple of 32: rows stay 128-byte aligned) that leaves less than 1e-9 of the response's mass behind */
    for (len = 32, tail = total; len < MAXLEN; len += 32) {
        tail = 0.0;
        for (n = len; n < MAXLEN; ++n) tail += fabs (g[n]);
        if (tail <= 1e-9 * total) break;
    }
    free (tmp);
    if (!(total > 0.0) || len >= MAXLEN || T + len > 1024) {
        fprintf (stderr, "libresampler_b200: the pre-filter's impulse response is too long to fold into %d taps (needs %d more)\n", T, len);
        free (g);
        return -1;
    }
    lead = len;                                 /* h'[m] = sum_k g[k] * row[m + k], m = -(len - 1) .. T - 1, plus one zero tap in front */
    Tk = T + lead;

    rows = calloc ((size_t) cxt->numFilters + 1, sizeof *rows);
    for (r = 0; r <= cxt->numFilters; ++r) {
        const artsample_t *row = cxt->filters[r];
        rows[r] = calloc (Tk, sizeof (artsample_t));
        for (m = -lead + 1; m < T; ++m) {
            double acc = 0.0;
            for (k = m < 0 ? -m : 0; k < len && m + k < T; ++k)
                acc += g[k] * (double) row[m + k];
            rows[r][m + lead] = (artsample_t) acc;
        }
    }
    mode |= device_mode (cxt->flags);
    fresh = artDevCreate (cxt->numChannels, Tk, lead, cxt->numFilters, mode, (const artsample_t *const *) rows);
    for (r = 0; r <= cxt->numFilters; ++r) free (rows[r]);
    free (rows);
    if (!fresh) {
        free (g);
        return -1;
    }
    artDevDestroy (cxt->device);
    cxt->device = fresh;
    cxt->prefilterLead = lead;
    cxt->prefilterTaps = g;
    return 0;
}

/* ------------------------------------------------------------------- the control loop */

/* Everything resampleProcess* decides before touching a sample (resampler.c:435-439, :491-492,
 * :494-535), evaluated in closed form by art_plan.h.  Updates the context's scalar state and
 * describes the call for the device. */
static ResampleResult plan_call (Resample *cxt, int numInputFrames, int numOutputFrames, double ratio, ArtCallPlan *call)
{
    const int T = cxt->numTaps, half = T / 2, NS = 16 * T, D = 15 * T;
    ResampleResult res;
    ArtLoopPlan lp;

    if (cxt->flags & RESAMPLE_FIXED_RATIO)      /* resampler.c:435-436 */
        ratio = cxt->fixedRatio;
    if (cxt->flags & RESAMPLER_FLUSHED)         /* resampler.c:438-439 */
        numInputFrames = 0;

    call->pre = 0;
    if (numInputFrames < 0) {                   /* flush: postfillAllChannels, resampler.c:663-685 */
        if (NS - cxt->inputIndex < half) {
            cxt->outputOffset -= D;
            cxt->inputIndex -= D;
        }
        cxt->flags |= RESAMPLER_FLUSHED;
        cxt->inputIndex += half;
        call->pre = half;
        numInputFrames = 0;
    }

    call->st.P = cxt->outputOffset;
    call->st.I = cxt->inputIndex;
    call->st.T = T;
    call->st.ratio = ratio;
    lp = art_plan_loop (&call->st, numInputFrames, numOutputFrames);

    cxt->outputOffset = lp.P_after;
    cxt->inputIndex = lp.I_after;
    if (cxt->flags & RESAMPLER_SNAP_OFFSET) {   /* resampler.c:533-535 */
        const double whole = floor (cxt->outputOffset);
        cxt->outputOffset = whole + floor ((cxt->outputOffset - whole) * cxt->numFilters + 0.5) / cxt->numFilters;
    }

    call->outputs = lp.outputs;
    call->inValid = (int) lp.inputs;            /* later frames are neither read nor consumed */
    call->consumed = call->pre + lp.inputs;
    res.input_used = lp.inputs;
    res.output_generated = lp.outputs;
    return res;
}

/* The four single-call entry points differ only in where the samples live. */
enum { IO_HOST_INTERLEAVED, IO_HOST_PLANAR, IO_DEVICE_INTERLEAVED, IO_DEVICE_PLANAR };

/* all device work reports failure as non-zero (art_device.h): the message has been printed and is kept for
 * resampleB200LastError(); the caller of the public API sees a call that consumed and produced nothing */
static int run_call (Resample *cxt, int io, const ArtCallPlan *call, const void *in, void *out, void *stream)
{
    switch (io) {
        case IO_HOST_INTERLEAVED:   return artDevRunHostInterleaved (cxt->device, call, (const artsample_t *) in, (artsample_t *) out);
        case IO_HOST_PLANAR:        return artDevRunHostPlanar (cxt->device, call, (const artsample_t *const *) in, (artsample_t *const *) out);
        case IO_DEVICE_INTERLEAVED: return artDevRunDeviceInterleaved (cxt->device, call, (const artsample_t *) in, (artsample_t *) out, stream);
        default:                    return artDevRunDevicePlanar (cxt->device, call, (const artsample_t *const *) in, (artsample_t *const *) out, stream);
    }
}

/* the scalar stream state a failed call has to put back */
typedef struct { double outputOffset; int inputIndex, flags; } ArtSaved;
static ArtSaved save_state (const Resample *cxt) { ArtSaved s; s.outputOffset = cxt->outputOffset; s.inputIndex = cxt->inputIndex; s.flags = cxt->flags; return s; }
static void restore_state (Resample *cxt, const ArtSaved *s) { cxt->outputOffset = s->outputOffset; cxt->inputIndex = s->inputIndex; cxt->flags = s->flags; }

/* EXTRAPOLATE_ENDPOINTS (resampler.c:516-522 / :663-698, extrapolator.c): at the end of a stream the T/2 frames the flush
 * appends are predicted from the last T/2 frames instead of being silence; at its start, when the first output is about
 * to be produced, the zero history in front of the first sample is replaced by a backward prediction from the samples
 * received so far.  Both are a few hundred samples of serial work per stream, done on the host (art_extrapolate.c) around
 * small synchronous transfers; the call itself then runs as any other. */
static int run_call_with_endpoints (Resample *cxt, int io, ArtCallPlan *call, const void *in, void *out, void *stream, int flushing)
{
    int failed = 0;
    const int T = cxt->numTaps, half = T / 2, C = cxt->numChannels;
    const int onDevice = io == IO_DEVICE_INTERLEAVED || io == IO_DEVICE_PLANAR;
    const int planar = io == IO_HOST_PLANAR || io == IO_DEVICE_PLANAR;
    void *st = onDevice ? stream : NULL;
    artsample_t *hist = NULL, *tail = NULL;           /* [C][T] history at call entry; [C][half] predicted flush block (planar) */
    const artsample_t **tailPlanes = NULL;
    artsample_t *tailInterleaved = NULL;
    int c, i;

    if (flushing || ((cxt->flags & EXTRAPOLATE_PREFILL) && call->outputs)) {
        hist = malloc (sizeof (artsample_t) * (size_t) C * T);
        if (artDevGetHistoryOn (cxt->device, hist, st)) {
            free (hist);
            return -1;
        }
    }

    if (flushing) {                             /* postfillAllChannels, resampler.c:663-685 */
        artsample_t *work = malloc (sizeof (artsample_t) * (size_t) T);
        tail = malloc (sizeof (artsample_t) * (size_t) C * half);
        for (c = 0; c < C; ++c) {
            memcpy (work, hist + (size_t) c * T + half, sizeof (artsample_t) * half);
            artExtendForward (work, half, half);
            memcpy (tail + (size_t) c * half, work + half, sizeof (artsample_t) * half);
        }
        free (work);
        call->inValid = call->pre;              /* the region's first T/2 frames now hold data */
    }

    if ((cxt->flags & EXTRAPOLATE_PREFILL) && call->outputs) {          /* prefillAllChannels, resampler.c:691-698 */
        const long long u0 = art_inputs_before (&call->st, 0);          /* frames pulled in this call before its first output */
        const long long n0 = (long long) call->st.I - call->pre - T;    /* frames received by earlier calls */
        const long long have = n0 + call->pre + u0;                     /* inputIndex - numTaps at that moment */
        cxt->flags &= ~EXTRAPOLATE_PREFILL;
        if (have >= 8 && have < T && n0 >= 0 && u0 <= call->inValid - (flushing ? call->pre : 0) + 0LL) {
            artsample_t *line = malloc (sizeof (artsample_t) * (size_t) T);         /* the ring's first T slots: [have, T) predicted, then ... */
            artsample_t *first = u0 ? malloc (sizeof (artsample_t) * (size_t) u0 * C) : NULL;   /* the call's first u0 input frames */
            if (u0) {
                if (!onDevice && !planar) memcpy (first, in, sizeof (artsample_t) * (size_t) u0 * C);
                else if (!onDevice)
                    for (c = 0; c < C; ++c) for (i = 0; i < u0; ++i) first[(size_t) i * C + c] = ((const artsample_t *const *) in)[c][i];
                else if (!planar) failed |= artDevFetch (cxt->device, (const artsample_t *) in, (size_t) u0 * C, first, st);
                else {
                    artsample_t *plane = malloc (sizeof (artsample_t) * (size_t) u0);
                    for (c = 0; c < C; ++c) {
                        failed |= artDevFetch (cxt->device, ((const artsample_t *const *) in)[c], (size_t) u0, plane, st);
                        for (i = 0; i < u0; ++i) first[(size_t) i * C + c] = plane[i];
                    }
                    free (plane);
                }
            }
            for (c = 0; c < C; ++c) {
                /* newest `have` samples, oldest first, at the end of `line`: earlier calls, the flush block, this call */
                artsample_t *seq = line + T - have;
                long long at = 0;
                for (i = 0; i < n0; ++i) seq[at++] = hist[(size_t) c * T + T - n0 + i];
                for (i = 0; i < call->pre; ++i) seq[at++] = tail ? tail[(size_t) c * half + i] : 0.0f;
                for (i = 0; i < u0; ++i) seq[at++] = first[(size_t) i * C + c];
                artExtendBackward (line + T, (int) have, (int) (T - have));
                /* line[k], k < T - have, is ring slot have + k: the slots in front of the first real sample.  The history
                 * at call entry is ring [n0, n0 + T), so they land at history index have - n0 + k */
                failed |= artDevPatchHistory (cxt->device, c, (int) (have - n0), (int) (T - have), line, st);
            }
            free (first);
            free (line);
        }
    }

    if (failed)
        ;
    else if (!flushing)
        failed = run_call (cxt, io, call, in, out, stream);
    else if (!onDevice) {
        if (planar) {
            tailPlanes = malloc (sizeof *tailPlanes * C);
            for (c = 0; c < C; ++c) tailPlanes[c] = tail + (size_t) c * half;
            failed = run_call (cxt, io, call, tailPlanes, out, stream);
        }
        else {
            tailInterleaved = malloc (sizeof (artsample_t) * (size_t) C * half);
            for (c = 0; c < C; ++c) for (i = 0; i < half; ++i) tailInterleaved[(size_t) i * C + c] = tail[(size_t) c * half + i];
            failed = run_call (cxt, io, call, tailInterleaved, out, stream);
        }
    }
    else if (planar) {
        const artsample_t *d_tail = artDevStage (cxt->device, tail, (size_t) C * half, st);
        if (!d_tail)
            failed = -1;
        else {
            tailPlanes = malloc (sizeof *tailPlanes * C);
            for (c = 0; c < C; ++c) tailPlanes[c] = d_tail + (size_t) c * half;
            failed = run_call (cxt, io, call, tailPlanes, out, stream);
        }
    }
    else {
        const artsample_t *d_tail;
        tailInterleaved = malloc (sizeof (artsample_t) * (size_t) C * half);
        for (c = 0; c < C; ++c) for (i = 0; i < half; ++i) tailInterleaved[(size_t) i * C + c] = tail[(size_t) c * half + i];
        d_tail = artDevStage (cxt->device, tailInterleaved, (size_t) C * half, st);
        failed = d_tail ? run_call (cxt, io, call, d_tail, out, stream) : -1;
    }
    free (tailPlanes);
    free (tailInterleaved);
    free (tail);
    free (hist);
    return failed;
}

/* Flush of a stream with a folded-in pre-filter.  The reference pads the FILTERED signal with T/2 zeros (the caller's biquads never
 * see the padding, art.c:1011-1017 / resampler.c:663-685); padding the raw signal instead would let the cascade ring on into the
 * last outputs.  So, for this one call per stream: take the raw history (T + lead frames), filter its newest T frames on the host,
 * and run the flush on a device context that holds the UNFUSED bank with that filtered history. */
static int run_flush_prefiltered (Resample *cxt, int io, const ArtCallPlan *call, void *out, void *stream)
{
    const int T = cxt->numTaps, lead = cxt->prefilterLead, Tk = T + lead, C = cxt->numChannels;
    const int onDevice = io == IO_DEVICE_INTERLEAVED || io == IO_DEVICE_PLANAR;
    artsample_t *raw = malloc (sizeof (artsample_t) * (size_t) C * Tk), *filtered = malloc (sizeof (artsample_t) * (size_t) C * T);
    ArtDev *fused = cxt->device;
    int c, i, k, failed;

    failed = artDevGetHistoryOn (fused, raw, onDevice ? stream : NULL);
    if (!failed && !cxt->plainDevice) {
        cxt->plainDevice = artDevCreate (C, T, 0, cxt->numFilters, device_mode (cxt->flags), (const artsample_t *const *) cxt->filters);
        failed = cxt->plainDevice == NULL;
    }
    if (!failed) {
        for (c = 0; c < C; ++c)
            for (i = 0; i < T; ++i) {
                double acc = 0.0;
                for (k = 0; k < lead; ++k)
                    acc += cxt->prefilterTaps[k] * (double) raw[(size_t) c * Tk + lead + i - k];
                filtered[(size_t) c * T + i] = (artsample_t) acc;
            }
        failed = artDevSetHistory (cxt->plainDevice, filtered);
    }
    if (!failed) {                                  /* (artDevSetHistory is synchronous: the history is in place for any stream) */
        cxt->device = cxt->plainDevice;
        failed = run_call (cxt, io, call, NULL, out, stream);
        cxt->device = fused;
    }
    free (raw);
    free (filtered);
    return failed;
}

static ResampleResult process_one (Resample *cxt, int io, const void *in, int numInputFrames, void *out, int numOutputFrames, double ratio, void *stream)
{
    const int flushing = numInputFrames < 0 && !(cxt->flags & RESAMPLER_FLUSHED);
    const ArtSaved before = save_state (cxt);
    ArtCallPlan call;
    ResampleResult res = plan_call (cxt, numInputFrames, numOutputFrames, ratio, &call);
    int failed;
    if (!(call.outputs || call.consumed))
        return res;
    if ((cxt->flags & EXTRAPOLATE_ENDPOINTS) && (flushing || (cxt->flags & EXTRAPOLATE_PREFILL)))
        failed = run_call_with_endpoints (cxt, io, &call, in, out, stream, flushing);
    else if (flushing && cxt->prefilterLead)
        failed = run_flush_prefiltered (cxt, io, &call, out, stream);
    else
        failed = run_call (cxt, io, &call, in, out, stream);
    if (failed) {                               /* nothing consumed, nothing produced, state as before the call */
        restore_state (cxt, &before);
        res.input_used = res.output_generated = 0;
    }
    return res;
}

ResampleResult resampleProcessInterleaved (Resample *cxt, const artsample_t *input, int numInputFrames, artsample_t *output, int numOutputFrames, double ratio)
{
    return process_one (cxt, IO_HOST_INTERLEAVED, input, numInputFrames, output, numOutputFrames, ratio, NULL);
}

ResampleResult resampleProcess (Resample *cxt, const artsample_t *const *input, int numInputFrames, artsample_t *const *output, int numOutputFrames, double ratio)
{
    return process_one (cxt, IO_HOST_PLANAR, input, numInputFrames, (void *) output, numOutputFrames, ratio, NULL);
}

ResampleResult resampleProcessInterleavedDevice (Resample *cxt, const artsample_t *d_input, int numInputFrames, artsample_t *d_output, int numOutputFrames, double ratio, void *stream)
{
    return process_one (cxt, IO_DEVICE_INTERLEAVED, d_input, numInputFrames, d_output, numOutputFrames, ratio, stream);
}

ResampleResult resampleProcessDevice (Resample *cxt, const artsample_t *const *d_input, int numInputFrames, artsample_t *const *d_output, int numOutputFrames, double ratio, void *stream)
{
    return process_one (cxt, IO_DEVICE_PLANAR, d_input, numInputFrames, (void *) d_output, numOutputFrames, ratio, stream);
}

/* resampler.c:741-758 */
ResampleResult resampleProcessAndFlushInterleaved (Resample *cxt, const artsample_t *input, int numInputFrames, artsample_t *output, int numOutputFrames, double ratio)
{
    ResampleResult res = resampleProcessInterleaved (cxt, input, numInputFrames, output, numOutputFrames, ratio), tail;

    if ((numInputFrames -= res.input_used) != 0 || (numOutputFrames -= res.output_generated) == 0)
        return res;
    tail = resampleProcessInterleaved (cxt, NULL, -1, output + (size_t) res.output_generated * cxt->numChannels, numOutputFrames, ratio);
    res.output_generated += tail.output_generated;
    return res;
}

/* resampler.c:712-739 */
ResampleResult resampleProcessAndFlush (Resample *cxt, const artsample_t *const *input, int numInputFrames, artsample_t *const *output, int numOutputFrames, double ratio)
{
    ResampleResult res = resampleProcess (cxt, input, numInputFrames, output, numOutputFrames, ratio), tail;
    artsample_t **rest;
    int c;

    if ((numInputFrames -= res.input_used) != 0 || (numOutputFrames -= res.output_generated) == 0)
        return res;
    rest = malloc (sizeof *rest * cxt->numChannels);
    for (c = 0; c < cxt->numChannels; ++c)
        rest[c] = output[c] + res.output_generated;
    tail = resampleProcess (cxt, NULL, -1, (artsample_t *const *) rest, numOutputFrames, ratio);
    free (rest);
    res.output_generated += tail.output_generated;
    return res;
}

/* ---------------------------------------------------------------- batched extensions */

/* does this call have to go through run_call_with_endpoints? */
static int needs_endpoint_work (const Resample *cxt, int numInputFrames)
{
    if (cxt->prefilterLead && numInputFrames < 0 && !(cxt->flags & RESAMPLER_FLUSHED))
        return 1;                               /* the flush of a pre-filtered stream runs on its own (run_flush_prefiltered) */
    if (!(cxt->flags & EXTRAPOLATE_ENDPOINTS))
        return 0;
    return (cxt->flags & EXTRAPOLATE_PREFILL) || (numInputFrames < 0 && !(cxt->flags & RESAMPLER_FLUSHED));
}


void resampleBatchProcessInterleavedDevice (Resample *const *cxts, int numContexts,
                                            const artsample_t *const *d_inputs, const int *numInputFrames,
                                            artsample_t *const *d_outputs, const int *numOutputFrames,
                                            const double *ratios, ResampleResult *results, void *stream)
{
    ArtCallPlan *calls;
    ArtDev **devs;
    const artsample_t **din;
    artsample_t **dout;
    ArtSaved *saved;
    int *owner;
    int i, n = 0, live = 0;

    if (numContexts <= 0)
        return;
    calls = malloc (sizeof *calls * numContexts);
    devs = malloc (sizeof *devs * numContexts);
    din = malloc (sizeof *din * numContexts);
    dout = malloc (sizeof *dout * numContexts);
    saved = malloc (sizeof *saved * numContexts);
    owner = malloc (sizeof *owner * numContexts);
    for (i = 0; i < numContexts; ++i) {
        ResampleResult r;
        if (needs_endpoint_work (cxts[i], numInputFrames[i])) {     /* stream start / end with extrapolation: on its own */
            r = process_one (cxts[i], IO_DEVICE_INTERLEAVED, d_inputs ? d_inputs[i] : NULL, numInputFrames[i], d_outputs[i],
                             numOutputFrames[i], ratios ? ratios[i] : 0.0, stream);
            if (results) results[i] = r;
            continue;
        }
        saved[n] = save_state (cxts[i]);
        owner[n] = i;
        r = plan_call (cxts[i], numInputFrames[i], numOutputFrames[i], ratios ? ratios[i] : 0.0, &calls[n]);
        devs[n] = cxts[i]->device;
        din[n] = d_inputs ? d_inputs[i] : NULL;
        dout[n] = d_outputs[i];
        if (results) results[i] = r;
        live |= calls[n].outputs || calls[n].consumed;
        ++n;
    }
    if (live && artDevRunBatchInterleaved (devs, calls, n, din, dout, stream))
        for (i = 0; i < n; ++i) {               /* the launch failed as a whole */
            restore_state (cxts[owner[i]], &saved[i]);
            if (results) results[owner[i]].input_used = results[owner[i]].output_generated = 0;
        }
    free (calls);
    free (devs);
    free (din);
    free (dout);
    free (saved);
    free (owner);
}

void resampleBatchProcessInterleaved (Resample *const *cxts, int numContexts,
                                      const artsample_t *const *inputs, const int *numInputFrames,
                                      artsample_t *const *outputs, const int *numOutputFrames,
                                      const double *ratios, ResampleResult *results)
{
    ArtCallPlan *calls;
    ArtDev **devs;
    const artsample_t **hin;
    artsample_t **hout;
    ArtSaved *saved;
    int *owner;
    int i, n = 0;

    if (numContexts <= 0)
        return;
    calls = malloc (sizeof *calls * numContexts);
    devs = malloc (sizeof *devs * numContexts);
    hin = malloc (sizeof *hin * numContexts);
    hout = malloc (sizeof *hout * numContexts);
    saved = malloc (sizeof *saved * numContexts);
    owner = malloc (sizeof *owner * numContexts);
    for (i = 0; i < numContexts; ++i) {
        ResampleResult r;
        if (needs_endpoint_work (cxts[i], numInputFrames[i])) {
            r = process_one (cxts[i], IO_HOST_INTERLEAVED, inputs ? inputs[i] : NULL, numInputFrames[i], outputs[i],
                             numOutputFrames[i], ratios ? ratios[i] : 0.0, NULL);
            if (results) results[i] = r;
            continue;
        }
        saved[n] = save_state (cxts[i]);
        owner[n] = i;
        r = plan_call (cxts[i], numInputFrames[i], numOutputFrames[i], ratios ? ratios[i] : 0.0, &calls[n]);
        devs[n] = cxts[i]->device;
        hin[n] = inputs ? inputs[i] : NULL;
        hout[n] = outputs[i];
        if (results) results[i] = r;
        ++n;
    }
    if (n && artDevRunHostBatchInterleaved (devs, calls, n, hin, hout))
        for (i = 0; i < n; ++i) {
            restore_state (cxts[owner[i]], &saved[i]);
            if (results) results[owner[i]].input_used = results[owner[i]].output_generated = 0;
        }
    free (calls);
    free (devs);
    free (hin);
    free (hout);
    free (saved);
    free (owner);
}

int resampleProcessBlocksInterleavedDevice (Resample *cxt, const artsample_t *d_input, const int *blockFrames,
                                            const double *ratios, int numBlocks,
                                            artsample_t *d_output, int outputCapacityFrames,
                                            ResampleResult *results, double *positions, void *stream)
{
    ArtCallPlan *calls;
    long long *inOff, *outOff, inAt = 0, outAt = 0, inFirst = 0;
    int b, done = 0, first = 0;
    ArtSaved atLaunch;

    if (numBlocks <= 0)
        return 0;
    calls = malloc (sizeof *calls * numBlocks);
    inOff = malloc (sizeof *inOff * numBlocks);
    outOff = malloc (sizeof *outOff * numBlocks);

    /* with endpoint extrapolation the blocks up to the first output go one at a time (the backward prediction needs
     * their samples on the host); once it is done the rest is one launch */
    while (done < numBlocks && (cxt->flags & EXTRAPOLATE_ENDPOINTS) && (cxt->flags & EXTRAPOLATE_PREFILL)) {
        Resample probe = *cxt;
        ArtCallPlan dry;
        long long room = (long long) outputCapacityFrames - outAt;
        ResampleResult r;
        if (blockFrames[done] < 0 || room <= 0)
            break;
        r = plan_call (&probe, blockFrames[done], room > 0x7fffffff ? 0x7fffffff : (int) room, ratios[done], &dry);
        if ((int) r.input_used != blockFrames[done])
            break;
        r = process_one (cxt, IO_DEVICE_INTERLEAVED, d_input + inAt * cxt->numChannels, blockFrames[done],
                         d_output + outAt * cxt->numChannels, room > 0x7fffffff ? 0x7fffffff : (int) room, ratios[done], stream);
        inAt += r.input_used;
        outAt += r.output_generated;
        if (results) results[done] = r;
        if (positions) positions[done] = resampleGetPosition (cxt);
        ++done;
    }
    first = done;
    inFirst = inAt;
    atLaunch = save_state (cxt);
    for (b = first; b < numBlocks; ++b) {
        /* dry-run on a copy of the scalar state: a block that cannot finish is not started */
        Resample probe = *cxt;
        long long room = (long long) outputCapacityFrames - outAt;
        ResampleResult r;

        if (blockFrames[b] < 0 || room <= 0)
            break;
        r = plan_call (&probe, blockFrames[b], room > 0x7fffffff ? 0x7fffffff : (int) room, ratios[b], &calls[b]);
        if ((int) r.input_used != blockFrames[b])
            break;
        cxt->outputOffset = probe.outputOffset;
        cxt->inputIndex = probe.inputIndex;
        cxt->flags = probe.flags;
        inOff[b] = inAt - inFirst;       /* relative to the first block of the launch: what precedes it is in the history */
        outOff[b] = outAt;
        inAt += r.input_used;
        outAt += r.output_generated;
        if (results) results[b] = r;
        if (positions) positions[b] = resampleGetPosition (cxt);
        ++done;
    }
    if (done > first && artDevRunBlocksInterleaved (cxt->device, calls + first, done - first, inOff + first, outOff + first,
                                                    d_input + inFirst * cxt->numChannels, d_output, stream)) {
        restore_state (cxt, &atLaunch);         /* the launch failed: the blocks it covered did not happen */
        done = first;
    }
    free (calls);
    free (inOff);
    free (outOff);
    return done;
}

/* ------------------------------------------------------------------ dry runs, position */

/* resampler.c:853-880: note that these accumulate 1/ratio where the real loop divides */
unsigned int resampleGetRequiredSamples (Resample *cxt, int numOutputFrames, double ratio)
{
    const int half = cxt->numTaps / 2, NS = 16 * cxt->numTaps, D = 15 * cxt->numTaps;
    int index = cxt->inputIndex;
    double offset = cxt->outputOffset;
    unsigned int used = 0;

    if (cxt->flags & RESAMPLE_FIXED_RATIO)
        ratio = cxt->fixedRatio;
    while (numOutputFrames > 0)
        if (offset >= index - half) {
            if (index == NS) { offset -= D; index -= D; }
            ++index;
            ++used;
        }
        else {
            offset += 1.0 / ratio;
            --numOutputFrames;
        }
    return used;
}

/* resampler.c:882-918 */
unsigned int resampleGetExpectedOutput (Resample *cxt, int numInputFrames, double ratio)
{
    const int half = cxt->numTaps / 2, NS = 16 * cxt->numTaps, D = 15 * cxt->numTaps;
    int index = cxt->inputIndex;
    double offset = cxt->outputOffset;
    unsigned int made = 0;

    if (cxt->flags & RESAMPLE_FIXED_RATIO)
        ratio = cxt->fixedRatio;
    if (cxt->flags & RESAMPLER_FLUSHED)
        numInputFrames = 0;
    else if (numInputFrames < 0)
        index += half;

    for (;;)
        if (offset >= index - half) {
            if (numInputFrames <= 0)
                break;
            if (index == NS) { offset -= D; index -= D; }
            ++index;
            --numInputFrames;
        }
        else {
            offset += 1.0 / ratio;
            ++made;
        }
    return made;
}

void resampleAdvancePosition (Resample *cxt, double delta)         /* resampler.c:927-935 */
{
    if (delta < 0.0)
        fprintf (stderr, "resampleAdvancePosition() can only advance forward!\n");
    else if (!(cxt->flags & SUBSAMPLE_INTERPOLATE) && floor (delta) != delta)
        fprintf (stderr, "resampleAdvancePosition() cannot advance partial samples without interpolation!\n");
    else
        cxt->outputOffset += delta;
}

double resampleGetPosition (Resample *cxt)                          /* resampler.c:965-968 */
{
    return cxt->outputOffset + (cxt->numTaps / 2.0) - cxt->inputIndex;
}

/* ----------------------------------------------------------------------- device helpers */

int resampleB200SetDevice (int device) { return artDevSelect (device); }
int resampleB200GetDeviceCount (void) { return artDevCount (); }

void resampleB200Synchronize (Resample *cxt) { artDevSynchronize (cxt->device); }
const char *resampleB200LastError (int clear) { return artDevLastError (clear); }
unsigned long long resampleB200KernelLaunches (void) { return artDevLaunchCount (); }
void resampleB200PathCounts (unsigned long long *generic, unsigned long long *periodic) { artDevPathCounts (generic, periodic); }
unsigned long long resampleB200TensorLaunches (void) { return artDevTensorLaunches (); }
void resampleB200SetTensorPath (int mode) { artDevSetTensorMode (mode); }
void resampleB200SetTensorDigits (int digits) { artDevSetTensorDigits (digits); }
void resampleB200ProfileEnable (int on) { artDevProfileEnable (on); }
unsigned long long resampleB200ProfileCollect (double *totalMs) { return artDevProfileCollect (totalMs); }
