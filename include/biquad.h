/*
 * biquad.h -- drop-in C API of libresampler_b200.so for the reference's biquad filters
 * (reference biquad.h:27-47, implementations biquad.c:18, :34, :51, :78, :106).
 *
 * biquad_lowpass / biquad_highpass / biquad_init are scalar coefficient set-up and run on
 * the host exactly as in the reference.  biquad_apply_buffer runs the recurrence on the
 * GPU (a chunked two-pass scan); the caller-owned `Biquad` keeps the same
 * layout and receives the same x[]/y[]/index state the reference would leave in it.
 * biquad_apply_sample is one multiply-add chain on one sample: it stays on the host and
 * is provided for source compatibility with decimator.c-style callers only.
 */
#ifndef ART_B200_BIQUAD_H
#define ART_B200_BIQUAD_H

#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

#ifndef ART_B200_RESAMPLER_H
#if defined(PATH_WIDTH) && (PATH_WIDTH==64)     /* reference biquad.h:21-25 */
typedef double artsample_t;
#else
typedef float artsample_t;
#endif
#endif

typedef struct {
    artsample_t a0, a1, a2, a3, a4, b1, b2, b3, b4;
} BiquadCoefficients;                           /* reference biquad.h:27-29 */

typedef struct {
    artsample_t a[5], b[5];                     /* coefficients */
    artsample_t x[4], y[4];                     /* delayed input/output, ring indexed by index & 3 */
    int order, index;
} Biquad;                                       /* reference biquad.h:31-35 */

#ifdef __cplusplus
extern "C" {
#endif

void biquad_init (Biquad *f, const BiquadCoefficients *coeffs, double gain);
void biquad_lowpass (BiquadCoefficients *filter, double frequency);
void biquad_highpass (BiquadCoefficients *filter, double frequency);
void biquad_apply_buffer (Biquad *f, artsample_t *buffer, int num_samples, int stride);
artsample_t biquad_apply_sample (Biquad *f, artsample_t input);

#ifdef __cplusplus
}
#endif
#endif
