/*
 * art_context.c -- host side (plain C) of libresampler_b200.so: the reference's public API
 * (include/resampler.h, include/resampler_b200.h), filter-bank design, and the scalar
 * streaming state.  Samples never pass through this file's arithmetic: every sample is
 * produced by the CUDA kernels behind art_device.h.
 */
#include "../../include/resampler_b200.h"
#include "art_device.h"

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
static void design_row (float *row, double *work, int taps, double fraction, double lowpass, int blackmanHarris)
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
            row[t] = (float) (work[t] - carry);
            carry += row[t] - work[t];
        }
    }
}

static float **design_bank (int taps, int filters, double lowpass, int flags)
{
    float **rows = calloc ((size_t) filters + 1, sizeof *rows);
    double *work = malloc (sizeof (double) * taps);
    int r, t;

    for (r = 0; r <= filters; ++r)
        rows[r] = calloc (taps, sizeof (float));
    for (r = 0; r < filters; ++r)                                   /* resampler.c:149-155 */
        design_row (rows[r], work, taps, (double) r / filters, lowpass, flags & BLACKMAN_HARRIS);
    for (t = 0; t < taps; ++t)                                      /* resampler.c:156-159 */
        rows[filters][(t + 1) % taps] = rows[0][t];
    rows[0][taps - 1] = 0.0f;                                       /* resampler.c:167-168 */
    rows[filters][0] = 0.0f;
    free (work);
    return rows;
}

/* ------------------------------------------------------------------------------ init */

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
    if (flags & EXTRAPOLATE_ENDPOINTS) {
        static int warned;
        if (!warned++)
            fprintf (stderr, "libresampler_b200: EXTRAPOLATE_ENDPOINTS is not implemented yet; endpoints are zero-extended\n");
        flags &= ~EXTRAPOLATE_ENDPOINTS;
    }

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

    if (flags & SUBSAMPLE_INTERPOLATE)   mode |= ART_MODE_INTERP;
    if (flags & INCLUDE_LOWPASS)         mode |= ART_MODE_LOWPASS;
    if (flags & EXTEND_CONVOLUTION_MATH) mode |= ART_MODE_PRECISE;   /* resampler.c:191-196 */

    cxt->device = artDevCreate (numChannels, numTaps, numFilters, mode, (const float *const *) cxt->filters);
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
}

void resampleFree (Resample *cxt)                                   /* resampler.c:973-995 */
{
    int r;
    if (!cxt)
        return;
    if (cxt->device)
        artDevDestroy (cxt->device);
    if (cxt->filters) {
        for (r = 0; r <= cxt->numFilters; ++r)
            free (cxt->filters[r]);
        free (cxt->filters);
    }
    free (cxt);
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

ResampleResult resampleProcessInterleaved (Resample *cxt, const artsample_t *input, int numInputFrames, artsample_t *output, int numOutputFrames, double ratio)
{
    ArtCallPlan call;
    ResampleResult res = plan_call (cxt, numInputFrames, numOutputFrames, ratio, &call);
    if (call.outputs || call.consumed)
        artDevRunHostInterleaved (cxt->device, &call, input, output);
    return res;
}

ResampleResult resampleProcess (Resample *cxt, const artsample_t *const *input, int numInputFrames, artsample_t *const *output, int numOutputFrames, double ratio)
{
    ArtCallPlan call;
    ResampleResult res = plan_call (cxt, numInputFrames, numOutputFrames, ratio, &call);
    if (call.outputs || call.consumed)
        artDevRunHostPlanar (cxt->device, &call, input, output);
    return res;
}

ResampleResult resampleProcessInterleavedDevice (Resample *cxt, const float *d_input, int numInputFrames, float *d_output, int numOutputFrames, double ratio, void *stream)
{
    ArtCallPlan call;
    ResampleResult res = plan_call (cxt, numInputFrames, numOutputFrames, ratio, &call);
    if (call.outputs || call.consumed)
        artDevRunDeviceInterleaved (cxt->device, &call, d_input, d_output, stream);
    return res;
}

ResampleResult resampleProcessDevice (Resample *cxt, const float *const *d_input, int numInputFrames, float *const *d_output, int numOutputFrames, double ratio, void *stream)
{
    ArtCallPlan call;
    ResampleResult res = plan_call (cxt, numInputFrames, numOutputFrames, ratio, &call);
    if (call.outputs || call.consumed)
        artDevRunDevicePlanar (cxt->device, &call, d_input, d_output, stream);
    return res;
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

void resampleBatchProcessInterleavedDevice (Resample *const *cxts, int numContexts,
                                            const float *const *d_inputs, const int *numInputFrames,
                                            float *const *d_outputs, const int *numOutputFrames,
                                            const double *ratios, ResampleResult *results, void *stream)
{
    ArtCallPlan *calls;
    ArtDev **devs;
    int i, live = 0;

    if (numContexts <= 0)
        return;
    calls = malloc (sizeof *calls * numContexts);
    devs = malloc (sizeof *devs * numContexts);
    for (i = 0; i < numContexts; ++i) {
        ResampleResult r = plan_call (cxts[i], numInputFrames[i], numOutputFrames[i], ratios ? ratios[i] : 0.0, &calls[i]);
        devs[i] = cxts[i]->device;
        if (results) results[i] = r;
        live |= calls[i].outputs || calls[i].consumed;
    }
    if (live)
        artDevRunBatchInterleaved (devs, calls, numContexts, d_inputs, d_outputs, stream);
    free (calls);
    free (devs);
}

void resampleBatchProcessInterleaved (Resample *const *cxts, int numContexts,
                                      const float *const *inputs, const int *numInputFrames,
                                      float *const *outputs, const int *numOutputFrames,
                                      const double *ratios, ResampleResult *results)
{
    ArtCallPlan *calls;
    ArtDev **devs;
    int i;

    if (numContexts <= 0)
        return;
    calls = malloc (sizeof *calls * numContexts);
    devs = malloc (sizeof *devs * numContexts);
    for (i = 0; i < numContexts; ++i) {
        ResampleResult r = plan_call (cxts[i], numInputFrames[i], numOutputFrames[i], ratios ? ratios[i] : 0.0, &calls[i]);
        devs[i] = cxts[i]->device;
        if (results) results[i] = r;
    }
    artDevRunHostBatchInterleaved (devs, calls, numContexts, inputs, outputs);
    free (calls);
    free (devs);
}

int resampleProcessBlocksInterleavedDevice (Resample *cxt, const float *d_input, const int *blockFrames,
                                            const double *ratios, int numBlocks,
                                            float *d_output, int outputCapacityFrames,
                                            ResampleResult *results, double *positions, void *stream)
{
    ArtCallPlan *calls;
    long long *inOff, *outOff, inAt = 0, outAt = 0;
    int b, done = 0;

    if (numBlocks <= 0)
        return 0;
    calls = malloc (sizeof *calls * numBlocks);
    inOff = malloc (sizeof *inOff * numBlocks);
    outOff = malloc (sizeof *outOff * numBlocks);

    for (b = 0; b < numBlocks; ++b) {
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
        inOff[b] = inAt;
        outOff[b] = outAt;
        inAt += r.input_used;
        outAt += r.output_generated;
        if (results) results[b] = r;
        if (positions) positions[b] = resampleGetPosition (cxt);
        ++done;
    }
    if (done)
        artDevRunBlocksInterleaved (cxt->device, calls, done, inOff, outOff, d_input, d_output, stream);
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
unsigned long long resampleB200KernelLaunches (void) { return artDevLaunchCount (); }
void resampleB200PathCounts (unsigned long long *generic, unsigned long long *periodic) { artDevPathCounts (generic, periodic); }
unsigned long long resampleB200TensorLaunches (void) { return artDevTensorLaunches (); }
void resampleB200SetTensorPath (int mode) { artDevSetTensorMode (mode); }
void resampleB200ProfileEnable (int on) { artDevProfileEnable (on); }
unsigned long long resampleB200ProfileCollect (double *totalMs) { return artDevProfileCollect (totalMs); }
