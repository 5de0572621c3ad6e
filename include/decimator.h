/*
 * decimator.h -- float -> integer (dither, noise shaping) and integer -> float stages either side of the resampling
 * path: the drop-in surface of the reference's decimator.h (decimator.h:29-75), implemented by libresampler_b200.so.
 *
 * Same names, argument meaning and results as the reference: output bytes and clipped-sample counts are bit-identical
 * (tests/test_gpu_decimator.py).  The quantiser sits inside the noise-shaping feedback loop and the dither generator
 * is a serial recurrence, so a channel is inherently sequential: the GPU runs one thread per channel and earns its
 * keep on many channels / many contexts at once (decimateBatchProcessInterleavedLE below); without dither and shaping
 * every sample is independent and the work is spread over samples.
 */
#ifndef ART_B200_DECIMATOR_H
#define ART_B200_DECIMATOR_H

#include <stdint.h>
#include "biquad.h"

/* reference decimator.h:29-41 */
#define DITHER_HIGHPASS     0x1
#define DITHER_FLAT         0x2
#define DITHER_LOWPASS      0x4
#define DITHER_ENABLED      (DITHER_HIGHPASS | DITHER_FLAT | DITHER_LOWPASS)

#define SHAPING_1ST_ORDER   0x100
#define SHAPING_2ND_ORDER   0x200
#define SHAPING_3RD_ORDER   0x400
#define SHAPING_ATH_CURVE   0x800
#define SHAPING_ENABLED     (SHAPING_1ST_ORDER | SHAPING_2ND_ORDER | SHAPING_3RD_ORDER | SHAPING_ATH_CURVE)

#define DECIMATE_MULTITHREADED  0x1000          /* accepted; channels always run in parallel on the GPU */

/* reference decimator.h:43-61: the leading fields keep the reference's names, types and order */
typedef struct {
    int numChannels, outputBits, outputBytes, dither_type, flags;
    double outputGain;
    artsample_t *feedback;                      /* [numChannels] quantisation-noise feedback            */
    uint32_t *tpdf_generators;                  /* [numChannels] dither generator state                 */
    Biquad *noise_shapers;                      /* [numChannels] H(z) noise-shaping filters             */
} Decimate;

#ifdef __cplusplus
extern "C" {
#endif

/* reference decimator.h:67-72, implementations decimator.c:416, :29, :112, :205, :341 */
void floatIntegersLE (unsigned char *input, double inputGain, int inputBits, int inputBytes, int inputStride, artsample_t *output, int numSamples);
Decimate *decimateInit (int numChannels, int outputBits, int outputBytes, double outputGain, int sampleRate, int flags);
int decimateProcessLE (Decimate *cxt, const artsample_t *const *input, int numInputFrames, unsigned char *const *output);
int decimateProcessInterleavedLE (Decimate *cxt, const artsample_t *input, int numInputFrames, unsigned char *output);
void decimateFree (Decimate *cxt);

/* ---- extensions (not in the reference) --------------------------------------------------------------------------- */
/* device-pointer twins: buffers in GPU memory, enqueued on `stream` (a cudaStream_t as void*, NULL = default), the clipped
 * count and the channel state come back after a synchronisation of that stream (they are a few words) */
void floatIntegersLEDevice (const unsigned char *d_input, double inputGain, int inputBits, int inputBytes, int inputStride,
                            artsample_t *d_output, int numSamples, void *stream);
int decimateProcessInterleavedLEDevice (Decimate *cxt, const artsample_t *d_input, int numInputFrames, unsigned char *d_output, void *stream);
/* many contexts of any configuration in ONE launch: what gives a per-channel-serial stage its parallelism.  clips[i]
 * (may be NULL) receives the clipped-sample count of context i; returns their sum.  Host pointers. */
int decimateBatchProcessInterleavedLE (Decimate *const *cxts, int numContexts, const artsample_t *const *inputs,
                                       const int *numInputFrames, unsigned char *const *outputs, int *clips);
int decimateBatchProcessInterleavedLEDevice (Decimate *const *cxts, int numContexts, const artsample_t *const *d_inputs,
                                             const int *numInputFrames, unsigned char *const *d_outputs, int *clips, void *stream);

#ifdef __cplusplus
}
#endif
#endif
