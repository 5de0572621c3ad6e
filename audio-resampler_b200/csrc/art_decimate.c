/*
 * art_decimate.c -- host side (plain C) of the float <-> integer stages (include/decimator.h): context set-up and
 * the per-channel state; every sample is converted by the CUDA kernels of art_decimate.cu.
 */
#include "../../include/decimator.h"
#include "art_device.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* N(z) of a noise shaper -> the decoupled H(z) form the feedback loop runs (reference decimator.c:383-402) */
static void shaper_from_nz (Biquad *f, const double *nz)
{
    BiquadCoefficients c;
    memset (&c, 0, sizeof c);
    if (nz[0] != 1.0) {
        fprintf (stderr, "shaper_init() error: a0 = %g, should be one!\n", nz[0]);
        exit (1);
    }
    c.a0 = nz[5] - nz[1]; c.a1 = nz[6] - nz[2]; c.a2 = nz[7] - nz[3]; c.a3 = nz[8] - nz[4];
    c.b1 = nz[5]; c.b2 = nz[6]; c.b3 = nz[7]; c.b4 = nz[8];
    biquad_init (f, &c, 1.0);
}

static uint32_t lcg15 (uint32_t x) { return ((x << 4) - x) ^ 1u; }

/* reference decimator.c:29-100 */
Decimate *decimateInit (int numChannels, int outputBits, int outputBytes, double outputGain, int sampleRate, int flags)
{
    /* {a0..a4, b1..b4} of N(z): the ATH curves by Sebastian Gesemann per sample rate (decimator.c:68-77), then orders 1-3 */
    static const struct { int rate; double nz[9]; } ath[] = {
        { 32000, { 1.0, -0.780459, +0.569358, -0.348221, +0.466316, +0.950797, +0.282052, +0.004337, +1.76209e-5 } },
        { 44100, { 1.0, -1.1474, 0.5383, -0.3530, 0.3475, 1.0587, 0.0676, -0.6054, -0.2738 } },
        { 48000, { 1.0, -1.3344, 0.7455, -0.4602, 0.4363, 0.9030, 0.0116, -0.5853, -0.2571 } },
        { 88200, { 1.0, -2.150679, +2.1402057, -1.042712, +0.206838, +0.67433, +1.017047, +0.4028633, +0.098656 } },
        { 96000, { 1.0, -2.16994, +2.01986, -0.894857, +0.1557738, +0.517789, +1.1062189, +0.4825786, +0.244994 } } };
    static const double order[3][9] = { { 1, -1, 0, 0, 0, 0, 0, 0, 0 }, { 1, -2, 1, 0, 0, 0, 0, 0, 0 }, { 1, -3, 3, -1, 0, 0, 0, 0, 0 } };
    Decimate *cxt = calloc (1, sizeof *cxt);
    int ch;

    cxt->numChannels = numChannels;
    cxt->outputBytes = outputBytes;
    cxt->outputBits = outputBits;
    cxt->outputGain = outputGain;
    cxt->flags = flags;
    cxt->feedback = calloc (numChannels, sizeof (artsample_t));

    if (flags & DITHER_ENABLED) {               /* the channels' seeds are consecutive bytes of one generator's output */
        unsigned char *seed = malloc (sizeof (uint32_t) * numChannels);
        uint32_t r = 0x31415926;
        size_t b;
        cxt->tpdf_generators = (uint32_t *) seed;
        for (b = 0; b < sizeof (uint32_t) * numChannels; ++b) {
            seed[b] = r >> 24;
            r = lcg15 (lcg15 (lcg15 (r)));
        }
        cxt->dither_type = (flags & DITHER_HIGHPASS) ? -1 : ((flags & DITHER_LOWPASS) ? 1 : 0);
    }

    if (flags & SHAPING_ENABLED) {
        const double *nz = order[0];            /* also the ATH fallback for non-standard rates */
        size_t k;
        if (flags & SHAPING_ATH_CURVE) {
            for (k = 0; k < sizeof ath / sizeof ath[0]; ++k)
                if (ath[k].rate == sampleRate) nz = ath[k].nz;
        }
        else if (flags & SHAPING_1ST_ORDER) nz = order[0];
        else if (flags & SHAPING_2ND_ORDER) nz = order[1];
        else if (flags & SHAPING_3RD_ORDER) nz = order[2];
        cxt->noise_shapers = calloc (numChannels, sizeof (Biquad));
        for (ch = 0; ch < numChannels; ++ch)
            shaper_from_nz (cxt->noise_shapers + ch, nz);
    }
    return cxt;
}

void decimateFree (Decimate *cxt)               /* reference decimator.c:341-356 */
{
    if (!cxt)
        return;
    free (cxt->tpdf_generators);
    free (cxt->noise_shapers);
    free (cxt->feedback);
    free (cxt);
}

/* ---- one launch over any number of contexts: gather the channels' state, run, scatter it back ---- */

static int run_batch (Decimate *const *cxts, int n, const void *const *in, int planarIn, const int *frames, void *const *out, int planarOut,
                      int *clips, int onDevice, void *stream)
{
    ArtDecLane *lanes;
    int total = 0, i, c, at = 0, sum = 0;

    for (i = 0; i < n; ++i) total += cxts[i]->numChannels;
    if (total <= 0)
        return 0;
    lanes = calloc (total, sizeof *lanes);
    for (i = 0; i < n; ++i) {
        const Decimate *d = cxts[i];
        const int used = (d->outputBits + 7) / 8;
        for (c = 0; c < d->numChannels; ++c, ++at) {
            ArtDecLane *l = &lanes[at];
            l->context = i;
            l->frames = frames[i] > 0 ? frames[i] : 0;
            l->bits = d->outputBits; l->bytes = d->outputBytes; l->pad = d->outputBytes - used;
            l->scaler = (artsample_t) ((1 << d->outputBits) / 2.0 * d->outputGain);
            l->dither = (d->flags & DITHER_ENABLED) ? 1 : 0;
            l->ditherType = d->dither_type;
            l->shaping = (d->flags & SHAPING_ENABLED) ? 1 : 0;
            l->rng = d->tpdf_generators ? d->tpdf_generators[c] : 0;
            l->feedback = d->feedback[c];
            if (planarIn)  { l->in = ((const artsample_t *const *) in[i])[c]; l->inStride = 1; }
            else           { l->in = (const artsample_t *) in[i] + c; l->inStride = d->numChannels; }
            if (planarOut) { l->out = ((unsigned char *const *) out[i])[c]; l->outStride = d->outputBytes; }
            else           { l->out = (unsigned char *) out[i] + (size_t) c * d->outputBytes; l->outStride = d->numChannels * d->outputBytes; }
            if (d->noise_shapers) {
                const Biquad *q = &d->noise_shapers[c];
                int k;
                memcpy (l->a, q->a, sizeof l->a);
                memcpy (l->b, q->b, sizeof l->b);
                for (k = 0; k < 4; ++k) {       /* newest first (the ring is addressed by index & 3, biquad.c:81) */
                    l->x[k] = q->x[(q->index - k) & 3];
                    l->y[k] = q->y[(q->index - k) & 3];
                }
                l->order = q->order;
            }
        }
    }
    if (artDecimateRun (lanes, total, n, cxts[0]->numChannels, onDevice, stream)) {
        free (lanes);
        return 0;
    }
    for (i = 0, at = 0; i < n; ++i) {
        Decimate *d = cxts[i];
        int mine = 0;
        for (c = 0; c < d->numChannels; ++c, ++at) {
            const ArtDecLane *l = &lanes[at];
            mine += l->clips;
            d->feedback[c] = l->feedback;
            if (d->tpdf_generators) d->tpdf_generators[c] = l->rng;
            if (d->noise_shapers && l->frames > 0) {
                Biquad *q = &d->noise_shapers[c];
                int k;
                q->index = (q->index + l->frames) & 3;      /* biquad_apply_sample keeps the index masked (biquad.c:96) */
                for (k = 0; k < 4; ++k) {
                    q->x[(q->index - k) & 3] = l->x[k];
                    q->y[(q->index - k) & 3] = l->y[k];
                }
            }
        }
        if (clips) clips[i] = mine;
        sum += mine;
    }
    free (lanes);
    return sum;
}

/* reference decimator.c:205-291 */
int decimateProcessInterleavedLE (Decimate *cxt, const artsample_t *input, int numInputFrames, unsigned char *output)
{
    const void *in = input; void *out = output;
    return run_batch (&cxt, 1, &in, 0, &numInputFrames, &out, 0, NULL, 0, NULL);
}

/* reference decimator.c:112-199 */
int decimateProcessLE (Decimate *cxt, const artsample_t *const *input, int numInputFrames, unsigned char *const *output)
{
    const void *in = input; void *out = (void *) output;
    return run_batch (&cxt, 1, &in, 1, &numInputFrames, &out, 1, NULL, 0, NULL);
}

int decimateProcessInterleavedLEDevice (Decimate *cxt, const artsample_t *d_input, int numInputFrames, unsigned char *d_output, void *stream)
{
    const void *in = d_input; void *out = d_output;
    return run_batch (&cxt, 1, &in, 0, &numInputFrames, &out, 0, NULL, 1, stream);
}

int decimateBatchProcessInterleavedLE (Decimate *const *cxts, int numContexts, const artsample_t *const *inputs,
                                       const int *numInputFrames, unsigned char *const *outputs, int *clips)
{
    return numContexts > 0 ? run_batch (cxts, numContexts, (const void *const *) inputs, 0, numInputFrames, (void *const *) outputs, 0, clips, 0, NULL) : 0;
}

int decimateBatchProcessInterleavedLEDevice (Decimate *const *cxts, int numContexts, const artsample_t *const *d_inputs,
                                             const int *numInputFrames, unsigned char *const *d_outputs, int *clips, void *stream)
{
    return numContexts > 0 ? run_batch (cxts, numContexts, (const void *const *) d_inputs, 0, numInputFrames, (void *const *) d_outputs, 0, clips, 1, stream) : 0;
}

/* reference decimator.c:416-450 */
void floatIntegersLE (unsigned char *input, double inputGain, int inputBits, int inputBytes, int inputStride, artsample_t *output, int numSamples)
{
    artFloatIntegersRun (input, inputGain, inputBits, inputBytes, inputStride, output, numSamples, 0, NULL);
}

void floatIntegersLEDevice (const unsigned char *d_input, double inputGain, int inputBits, int inputBytes, int inputStride,
                            artsample_t *d_output, int numSamples, void *stream)
{
    artFloatIntegersRun (d_input, inputGain, inputBits, inputBytes, inputStride, d_output, numSamples, 1, stream);
}
