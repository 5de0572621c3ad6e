/*
 * resampler_b200.h -- extension entry points of libresampler_b200.so (not in the reference).
 *
 * The reference API takes host pointers and returns synchronously, which on a GPU measures
 * PCIe.  These entry points keep the reference's semantics (same ResampleResult, same
 * position bookkeeping -- they share the planner with resampleProcess*, resampler.c:433-658)
 * but take DEVICE pointers, enqueue on a CUDA stream and return without synchronising.
 * `stream` is a cudaStream_t passed as void* (NULL = the context's private stream); no CUDA
 * or torch type appears in any signature.
 *
 * Stream discipline: the calls on ONE context must be ordered -- use one stream per context, or
 * synchronise between streams yourself (the context's history and its filter tables are written
 * by every call).  resampleReset() waits for the stream the context last ran on.
 */
#ifndef ART_B200_EXT_H
#define ART_B200_EXT_H

#include "resampler.h"
#include "biquad.h"

#ifdef __cplusplus
extern "C" {
#endif

/* which GPU a context lives on is the CUDA current device at init time; these helpers let a
 * non-CUDA host program pick it */
int  resampleB200SetDevice (int device);                 /* 0 on success */
int  resampleB200GetDeviceCount (void);
void resampleB200Synchronize (Resample *cxt);            /* wait for the context's private stream */
unsigned long long resampleB200KernelLaunches (void);    /* kernels launched by this library so far */
/* Error reporting.  The reference API has no error codes (resampler.h:64-78): init returns NULL after a message on stderr,
 * process calls "never fail".  Here a process call can fail (CUDA error, out of device memory, a batch that mixes
 * configurations): the library then prints the message once, leaves the context's position where it was, reports
 * input_used = output_generated = 0 -- and never aborts the caller's process.  LastError returns the message of the most
 * recent failure on the calling thread (NULL when there was none); clear != 0 forgets it. */
const char *resampleB200LastError (int clear);
/* how many convolution launches went to the any-ratio kernel and to the rational-ratio kernel */
void resampleB200PathCounts (unsigned long long *generic, unsigned long long *periodic);
/* The rational-ratio path has two forms: FFMA kernels, and a tensor-core (tcgen05) kernel used for interpolating contexts
 * when a launch holds enough work to fill the GPU.  mode 0: never use the tensor-core kernel, 1 (default): when the launch is
 * large enough, 2: whenever the configuration is eligible (tests), 3: as 1, and also for non-interpolating contexts
 * (resampleFixedRatioInit): their output then still matches the reference within 1e-6 of peak but is no longer bit-identical
 * across different call chunkings, which the FFMA form guarantees as the reference does.  The environment variable ART_B200_UMMA sets the initial
 * mode.  TensorLaunches counts its launches (PathCounts' `periodic` counts the FFMA form only). */
void resampleB200SetTensorPath (int mode);
unsigned long long resampleB200TensorLaunches (void);
/* Arithmetic of the tensor-core form.  A tile (128 MMA rows: 128 / c periods of c = 1, 2 or 4 channels, 0.1-0.4 s of signal) is
 * converted to fixed point relative to ITS maximum and cut into fp16 digits of 11 bits.
 *   3 digits (default): every sample keeps >= 22 significant bits of its own magnitude, whatever else the tile holds: the output is
 *                       within float rounding (observed <= 3e-7) of the reference relative to the LOCAL signal level, like the
 *                       reference's own float arithmetic and this library's FFMA kernels.  Six MMAs per 16 taps.
 *   2 digits:           samples are exact to 2^-24 of the TILE's maximum -- below the quantisation step of 24-bit PCM, but a passage
 *                       60 dB under the loudest sample of its tile is only accurate to ~5e-5 of its own level.  Five MMAs per 16
 *                       taps, ~15 % faster.
 * The environment variable ART_B200_DIGITS sets the initial value. */
void resampleB200SetTensorDigits (int digits);
/* Non-finite input (NaN, Inf): the reference's outputs are non-finite exactly where a filter window holds such a sample.  The
 * kernels here evaluate zero-padded windows (0 x NaN is NaN), so they return non-finite values for those outputs and for up to one
 * period of outputs (the ratio's numerator, 160 at 44.1k->48k) on either side; every other output is unaffected (a non-finite
 * sample is left out of a tile's block scaling).  tests/test_gpu_tensor_path.py pins this. */
/* measurement aid: when enabled, every convolution kernel launch is bracketed by CUDA events on its
 * own stream; Collect waits for them, returns how many launches were timed and their summed
 * duration in milliseconds, and clears the list */
void resampleB200ProfileEnable (int on);
unsigned long long resampleB200ProfileCollect (double *totalMs);

/* device-pointer twins of resampleProcessInterleaved (resampler.c:550) / resampleProcess (:433).
 * A flush is numInputFrames == -1, as in the reference. */
ResampleResult resampleProcessInterleavedDevice (Resample *cxt, const artsample_t *d_input, int numInputFrames,
                                                 artsample_t *d_output, int numOutputFrames, double ratio, void *stream);
ResampleResult resampleProcessDevice (Resample *cxt, const artsample_t *const *d_input, int numInputFrames,
                                      artsample_t *const *d_output, int numOutputFrames, double ratio, void *stream);

/* Many independent contexts of identical configuration (same channels/taps/filters/lowpass/flags,
 * same GPU) in ONE launch -- what workers.c's per-channel threads and a caller's per-stream loop
 * become on a GPU.  Element i of every array belongs to cxts[i].  results may be NULL. */
void resampleBatchProcessInterleavedDevice (Resample *const *cxts, int numContexts,
                                            const artsample_t *const *d_inputs, const int *numInputFrames,
                                            artsample_t *const *d_outputs, const int *numOutputFrames,
                                            const double *ratios, ResampleResult *results, void *stream);

/* The same for HOST buffers: uploads, kernels and downloads of successive contexts are pipelined on
 * three streams (PCIe in both directions overlaps the convolution); returns when all outputs are in
 * host memory.  Pinned host buffers are needed for the copies to be asynchronous. */
void resampleBatchProcessInterleaved (Resample *const *cxts, int numContexts,
                                      const artsample_t *const *inputs, const int *numInputFrames,
                                      artsample_t *const *outputs, const int *numOutputFrames,
                                      const double *ratios, ResampleResult *results);

/* ASRC: numBlocks consecutive blocks of ONE stream, block b holding blockFrames[b] input frames
 * and resampled at ratios[b], exactly as numBlocks successive resampleProcessInterleaved calls
 * would be (positions[b], if non-NULL, receives resampleGetPosition after block b).  Input blocks
 * are contiguous in d_input; outputs are packed contiguously into d_output.  Stops early -- and
 * returns the number of blocks completed -- when a block cannot consume all of its input within
 * outputCapacityFrames. */
int resampleProcessBlocksInterleavedDevice (Resample *cxt, const artsample_t *d_input, const int *blockFrames,
                                            const double *ratios, int numBlocks,
                                            artsample_t *d_output, int outputCapacityFrames,
                                            ResampleResult *results, double *positions, void *stream);

/* Fused pre-filter.  art.c runs a cascade of biquad lowpass sections over the input block in front of a downsampling resampler
 * (art.c:848-851, :1011-1017): three more passes over memory on a GPU.  A cascade whose impulse response has died out within a
 * few hundred samples -- every lowpass art.c designs: 32 taps reach 1e-10 at 0.45 * 44.1 / 96 -- is a short FIR filter, and
 * (resampling filter) o (FIR) is again a bank of FIR filters: this call convolves every row of the context's bank with the
 * cascade's impulse response (in double, from the float coefficients biquad_init stored) and from then on the context resamples
 * AND pre-filters in one pass -- zero extra bytes, a few per cent more taps.  The caller then skips its own biquad_apply_buffer
 * calls.  Counts and positions are unchanged (the control loop still runs on numTaps); samples agree with "reference biquads,
 * then reference resampler" within 1e-6 of peak (the reference's float32 recursion itself sits 1.6e-7 from the exact response).
 * `sections` are numSections initialised Biquad structs (biquad_init) in zero state, applied in order to every channel.
 * Returns 0 on success; non-zero (with a message, context unchanged) when the stream has already consumed input, when
 * EXTRAPOLATE_ENDPOINTS is set (the reference extrapolates the FILTERED signal), or when the response is too long
 * (numTaps + its length > 1024). */
int resampleB200AttachPrefilter (Resample *cxt, const Biquad *sections, int numSections);

/* The cascade art.c applies around the resampler (art.c:1011-1017, :1052-1058): numStages sets of
 * numChannels Biquads over one interleaved buffer, in ONE pass over memory.  stages[s] points at
 * the caller's array of numChannels Biquad structs for stage s (e.g. lowpass1, lowpass2). */
void biquad_apply_cascade_interleaved (Biquad *const *stages, int numStages, int numChannels,
                                       artsample_t *buffer, int numFrames);
void biquad_apply_cascade_interleaved_device (Biquad *const *stages, int numStages, int numChannels,
                                              artsample_t *d_buffer, int numFrames, void *stream);

#ifdef __cplusplus
}
#endif
#endif
