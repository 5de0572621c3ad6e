/*
 * resampler.h -- drop-in C API of libresampler_b200.so for the windowed-sinc resampling
 * path of dbry/audio-resampler.
 *
 * Every prototype, flag value and the leading fields of `Resample` are those of the
 * reference's resampler.h (reference lines cited per item) so that a caller such as
 * art.c / artest.c recompiles against this header unchanged.  The implementation behind
 * it is CUDA (sm_100a); there is no CPU fallback: without a usable Blackwell GPU the init
 * functions print the CUDA error and return NULL.
 *
 * Like the reference, the header serves two builds: compiled with -DPATH_WIDTH=64 every sample
 * (buffers, filter bank, biquad and decimator state) is a double and the caller links
 * libresampler_b200_64.so (what the reference's art64 / artest64 targets do with its own
 * sources, Makefile:12-19); otherwise samples are floats and the library is libresampler_b200.so.
 */
#ifndef ART_B200_RESAMPLER_H
#define ART_B200_RESAMPLER_H

#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#if defined(PATH_WIDTH) && (PATH_WIDTH==64)     /* reference resampler.h:22-26 */
typedef double artsample_t;
#else
typedef float artsample_t;
#endif

/* reference resampler.h:28-38 */
#define SUBSAMPLE_INTERPOLATE   0x1
#define BLACKMAN_HARRIS         0x2
#define INCLUDE_LOWPASS         0x4
#define RESAMPLE_MULTITHREADED  0x8             /* accepted; channels are always processed in parallel on the GPU */
#define NO_FILTER_REDUCTION     0x10
#define RESAMPLE_FIXED_RATIO    0x20            /* internal use only, do not set */
#define EXTRAPOLATE_ENDPOINTS   0x40
#define EXTRAPOLATE_PREFILL     0x80            /* internal use only, do not set */
#define EXTEND_CONVOLUTION_MATH 0x100
#define RESAMPLER_FLUSHED       0x200           /* internal use only, do not set */
#define RESAMPLER_SNAP_OFFSET   0x400           /* internal use only, do not set */

/* reference resampler.h:40-42 */
typedef struct {
    unsigned int input_used, output_generated;
} ResampleResult;

/* reference resampler.h:44-58.  The leading fields keep the reference's names, types and
 * order (artest.c:660-669 reads cxt->numChannels).  `filters` holds the host copy of the
 * (numFilters + 1) x numTaps bank; `buffers` is NULL -- the sample history lives in HBM. */
typedef struct resample {
    int numChannels, numSamples, numFilters, numTaps, inputIndex, flags;
    double *tempFilter, outputOffset, fixedRatio, lowpassRatio;
    double (*subsample)(struct resample *cxt, artsample_t *source, double offset);   /* always NULL here */
    artsample_t **buffers, **filters;
    void *device;                               /* private: the CUDA side of the context */
    int prefilterLead;                          /* private: taps a folded-in pre-filter adds in front of every window (resampler_b200.h) */
    double *prefilterTaps;                      /* private: its impulse response (prefilterLead taps) */
    void *plainDevice;                          /* private: device context with the unfused bank, used for the flush of a pre-filtered stream */
} Resample;

#ifdef __cplusplus
extern "C" {
#endif

/* reference resampler.h:64-78, implementations resampler.c:115, :310, :433, :550, :712, :741,
 * :853, :882, :927, :365, :965, :370, :375, :383, :973 */
Resample *resampleInit (int numChannels, int numTaps, int numFilters, double lowpassRatio, int flags);
Resample *resampleFixedRatioInit (int numChannels, int numTaps, int maxFilters, double sourceRate, double destinRate, int lowpassFreq, int flags);
ResampleResult resampleProcess (Resample *cxt, const artsample_t *const *input, int numInputFrames, artsample_t *const *output, int numOutputFrames, double ratio);
ResampleResult resampleProcessInterleaved (Resample *cxt, const artsample_t *input, int numInputFrames, artsample_t *output, int numOutputFrames, double ratio);
ResampleResult resampleProcessAndFlush (Resample *cxt, const artsample_t *const *input, int numInputFrames, artsample_t *const *output, int numOutputFrames, double ratio);
ResampleResult resampleProcessAndFlushInterleaved (Resample *cxt, const artsample_t *input, int numInputFrames, artsample_t *output, int numOutputFrames, double ratio);
unsigned int resampleGetRequiredSamples (Resample *cxt, int numOutputFrames, double ratio);
unsigned int resampleGetExpectedOutput (Resample *cxt, int numInputFrames, double ratio);
void resampleAdvancePosition (Resample *cxt, double delta);
double resampleGetLowpassRatio (Resample *cxt);
double resampleGetPosition (Resample *cxt);
int resampleGetNumFilters (Resample *cxt);
int resampleInterpolationUsed (Resample *cxt);
void resampleReset (Resample *cxt);
void resampleFree (Resample *cxt);

#ifdef __cplusplus
}
#endif
#endif
