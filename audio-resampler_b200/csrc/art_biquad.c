/*
 * art_biquad.c -- host side (plain C) of the biquad API (include/biquad.h and the cascade
 * extension in include/resampler_b200.h).  Coefficient design and struct set-up are scalar
 * and stay here; every buffer is filtered on the GPU by art_biquad.cu.
 */
#include "../../include/resampler_b200.h"
#include "art_device.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* second-order Butterworth-Q designs, reference biquad.c:18-30 and :34-46.  Note that the
 * coefficients are rounded to float as they are stored, and a1 of the lowpass is twice the
 * ROUNDED a0 (biquad.c:26). */
void biquad_lowpass (BiquadCoefficients *filter, double frequency)
{
    const double Q = sqrt (0.5), K = tan (M_PI * frequency);
    const double norm = 1.0 / (1.0 + K / Q + K * K);

    memset (filter, 0, sizeof *filter);
    filter->a0 = K * K * norm;
    filter->a1 = 2 * filter->a0;
    filter->a2 = filter->a0;
    filter->b1 = 2.0 * (K * K - 1.0) * norm;
    filter->b2 = (1.0 - K / Q + K * K) * norm;
}

void biquad_highpass (BiquadCoefficients *filter, double frequency)
{
    const double Q = sqrt (0.5), K = tan (M_PI * frequency);
    const double norm = 1.0 / (1.0 + K / Q + K * K);

    memset (filter, 0, sizeof *filter);
    filter->a0 = norm;
    filter->a1 = -2.0 * norm;
    filter->a2 = filter->a0;
    filter->b1 = 2.0 * (K * K - 1.0) * norm;
    filter->b2 = (1.0 - K / Q + K * K) * norm;
}

/* reference biquad.c:51-74 */
void biquad_init (Biquad *f, const BiquadCoefficients *c, double gain)
{
    const artsample_t fwd[5] = { c->a0, c->a1, c->a2, c->a3, c->a4 };
    const artsample_t bwd[5] = { 0.0f, c->b1, c->b2, c->b3, c->b4 };
    int d;

    memset (f, 0, sizeof *f);
    for (d = 0; d < 5; ++d) {
        f->a[d] = fwd[d] * gain;
        f->b[d] = bwd[d];
    }
    f->order = 1;
    for (d = 2; d <= 4; ++d)
        if (fwd[d] != 0.0f || bwd[d] != 0.0f)
            f->order = d;
}

/* reference biquad.c:78-102.  One sample is one short dependent chain of multiply-adds --
 * there is nothing to parallelise, so this convenience call (unused on the resampling path)
 * is the only sample arithmetic the library does on the host. */
artsample_t biquad_apply_sample (Biquad *f, artsample_t input)
{
    artsample_t sum = input * f->a[0];
    int i = f->index & 3, d;

    for (d = f->order; d >= 1; --d)
        sum += (f->x[(i - (d - 1)) & 3] * f->a[d]) - (f->b[d] * f->y[(i - (d - 1)) & 3]);
    f->index = i = (i + 1) & 3;
    f->x[i] = input;
    f->y[i] = sum;
    return sum;
}

/* the reference keeps its delays in 4-entry rings addressed by index & 3 (biquad.c:108-119);
 * the device code wants them newest-first */
static void ring_to_stage (const Biquad *f, ArtBiquadStage *s)
{
    int d;
    memcpy (s->a, f->a, sizeof s->a);
    memcpy (s->b, f->b, sizeof s->b);
    for (d = 0; d < 4; ++d) {
        s->x[d] = f->x[(f->index - d) & 3];
        s->y[d] = f->y[(f->index - d) & 3];
    }
    s->order = f->order;
}

static void stage_to_ring (const ArtBiquadStage *s, Biquad *f, int advanced)
{
    int d;
    f->index += advanced;                          /* biquad.c:162 stores the unmasked counter */
    for (d = 0; d < 4; ++d) {
        f->x[(f->index - d) & 3] = s->x[d];
        f->y[(f->index - d) & 3] = s->y[d];
    }
}

static void run_cascade (Biquad *const *stages, int numStages, int numChannels, artsample_t *buffer, int numFrames, int stride, int onDevice, void *stream)
{
    ArtBiquadStage *flat;
    int s, c;

    if (numFrames <= 0 || numStages <= 0 || numChannels <= 0)
        return;
    flat = malloc (sizeof *flat * numStages * numChannels);
    for (s = 0; s < numStages; ++s)
        for (c = 0; c < numChannels; ++c)
            ring_to_stage (&stages[s][c], &flat[s * numChannels + c]);
    artBiquadRun (flat, numStages, numChannels, buffer, numFrames, stride, onDevice, stream);
    for (s = 0; s < numStages; ++s)
        for (c = 0; c < numChannels; ++c)
            stage_to_ring (&flat[s * numChannels + c], &stages[s][c], numFrames);
    free (flat);
}

/* reference biquad.c:106-163 */
void biquad_apply_buffer (Biquad *f, artsample_t *buffer, int num_samples, int stride)
{
    Biquad *one[1];
    one[0] = f;
    run_cascade (one, 1, 1, buffer, num_samples, stride, 0, NULL);
}

void biquad_apply_cascade_interleaved (Biquad *const *stages, int numStages, int numChannels, artsample_t *buffer, int numFrames)
{
    run_cascade (stages, numStages, numChannels, buffer, numFrames, numChannels, 0, NULL);
}

void biquad_apply_cascade_interleaved_device (Biquad *const *stages, int numStages, int numChannels, artsample_t *d_buffer, int numFrames, void *stream)
{
    run_cascade (stages, numStages, numChannels, d_buffer, numFrames, numChannels, 1, stream);
}
