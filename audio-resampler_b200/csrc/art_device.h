/*
 * art_device.h -- the thin C ABI between the C host code (art_context.c, art_biquad.c)
 * and the CUDA translation units (art_device.cu, art_sinc_kernels.cu, ...).
 * Internal: nothing here is exported from libresampler_b200.so.
 *
 * Division of labour
 *   host (C)   : the reference's API surface and scalar state -- filter-bank design in
 *                double, outputOffset/inputIndex bookkeeping through art_plan.h, flags.
 *   device     : everything that touches samples -- the per-channel history (replaces the
 *                16*T ring of resampler.c:139,171-174), the filter bank, the convolution.
 *
 * There is no CPU implementation of any of these: when no usable CUDA device exists
 * artDevCreate() reports the CUDA error and returns NULL.  Later failures (a CUDA error, an
 * allocation that does not fit, an unsupported request) never take the process down: the
 * entry point prints the message once, keeps it for artDevLastError(), and returns non-zero
 * (NULL for the pointer-valued ones); the caller's buffers are then unspecified.
 */
#ifndef ART_DEVICE_H
#define ART_DEVICE_H

#include "art_plan.h"
#include "art_sample.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ArtDev ArtDev;

/* What one resampleProcess* call has to do, as decided on the host by art_plan.h. */
typedef struct {
    ArtLoopState st;            /* loop-entry state (after the flush adjustment, if any)             */
    int          pre;           /* input-region frames swallowed before the loop: T/2 on flush, else 0 */
    int          inValid;       /* frames of caller data in the input region; the rest reads as zero  */
    unsigned int outputs;       /* output_generated                                                   */
    unsigned int consumed;      /* pre + input_used = frames of the input region entering the history */
} ArtCallPlan;

/* Kernel selection hints, fixed per context. */
enum {
    ART_MODE_INTERP  = 1,       /* SUBSAMPLE_INTERPOLATE                                   */
    ART_MODE_LOWPASS = 2,       /* INCLUDE_LOWPASS (disables the pass-through shortcut)     */
    ART_MODE_PRECISE = 4        /* EXTEND_CONVOLUTION_MATH: double accumulation             */
};

/* rows: (filters + 1) rows of `taps` floats.  `lead` of those taps lie in front of the reference's window (a pre-filter folded into
 * the bank, art_context.c): the control loop then runs on taps - lead, the history holds `taps` frames per channel. */
ArtDev *artDevCreate (int channels, int taps, int lead, int filters, int mode, const artsample_t *const *rows);
void    artDevDestroy (ArtDev *dev);
int     artDevReset (ArtDev *dev);                   /* zero the history (resampler.c:387-388) */
int     artDevDeviceIndex (const ArtDev *dev);
int     artDevSelect (int device);                   /* cudaSetDevice; 0 on success            */
int     artDevCount (void);                          /* usable CUDA devices, 0 when none       */

/* Host-memory entry points: stage in, run, stage out, synchronise. */
int  artDevRunHostInterleaved (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *in, artsample_t *out);
int  artDevRunHostPlanar (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *const *in, artsample_t *const *out);
int  artDevRunHostBatchInterleaved (ArtDev *const *devs, const ArtCallPlan *plans, int count,
                                    const artsample_t *const *in, artsample_t *const *out);

/* Device-memory entry points: enqueue on `stream` (a cudaStream_t, NULL = the context's
 * own stream) and return without synchronising. */
int  artDevRunDeviceInterleaved (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *d_in, artsample_t *d_out, void *stream);
int  artDevRunDevicePlanar (ArtDev *dev, const ArtCallPlan *plan, const artsample_t *const *d_in, artsample_t *const *d_out, void *stream);

/* Many contexts of one configuration in a single launch (device memory, interleaved). */
int  artDevRunBatchInterleaved (ArtDev *const *devs, const ArtCallPlan *plans, int count,
                                const artsample_t *const *d_in, artsample_t *const *d_out, void *stream);

/* Consecutive blocks of ONE stream, each with its own ratio (ASRC), in a single launch.
 * Block b reads d_in + inOffset[b]*channels; the frames before it in the same buffer are
 * its history.  Only valid when every block consumed all of its input. */
int  artDevRunBlocksInterleaved (ArtDev *dev, const ArtCallPlan *plans, int count,
                                 const long long *inOffset, const long long *outOffset,
                                 const artsample_t *d_in, artsample_t *d_out, void *stream);

int  artDevSynchronize (ArtDev *dev);
int  artDevGetHistory (ArtDev *dev, artsample_t *hostPlanar);        /* [channels][taps], for tests/extrapolation */
int  artDevSetHistory (ArtDev *dev, const artsample_t *hostPlanar);
/* endpoint extrapolation: small synchronous transfers at a stream's start and end (stream NULL = the context's own) */
int  artDevGetHistoryOn (ArtDev *dev, artsample_t *hostPlanar, void *stream);
int  artDevPatchHistory (ArtDev *dev, int channel, int first, int count, const artsample_t *values, void *stream);
int  artDevFetch (ArtDev *dev, const artsample_t *d_src, size_t floats, artsample_t *host, void *stream);
artsample_t *artDevStage (ArtDev *dev, const artsample_t *host, size_t floats, void *stream);     /* returns the device copy */

/* message of the last failure on the calling thread, or NULL; clear != 0 forgets it */
const char *artDevLastError (int clear);

/* statistics for bench.py's gpu_launches claim and roofline leg */
unsigned long long artDevLaunchCount (void);
void artDevPathCounts (unsigned long long *generic, unsigned long long *periodic);
unsigned long long artDevTensorLaunches (void);               /* launches of the tensor-core kernel (not counted above) */
void artDevSetTensorMode (int mode);                          /* 0 never, 1 large launches (default), 2 whenever eligible */
void artDevSetTensorDigits (int digits);                      /* signal digits of the tensor-core form: 3 (default) or 2 */
void artDevProfileEnable (int on);
unsigned long long artDevProfileCollect (double *totalMs);   /* returns timed launches, clears */

/* ---- biquad cascade (biquad.c:106-163, order <= 4), float32 direct form I ---------- */
typedef struct {
    artsample_t a[5], b[5];     /* as stored in the reference's Biquad (biquad.h:31-35)      */
    artsample_t x[4], y[4];     /* oldest..newest delayed input/output: x[0] = x[n-1] ...    */
    int   order;
} ArtBiquadStage;

/* buffer: host or device memory holding frames*stride floats; channel c of stage-set s uses
 * stages[s*channels + c] and samples buffer[c + f*stride].  States are read from and written
 * back to `stages` (host memory).  onDevice selects the address space of `buffer`. */
int  artBiquadRun (ArtBiquadStage *stages, int numStages, int channels, artsample_t *buffer,
                   long long frames, int stride, int onDevice, void *stream);

/* ---- float <-> integer stages (decimator.c), art_decimate.cu ---------------------------------------------------- */
/* one channel of one context: where its samples are, its conversion parameters and its state (in and out) */
typedef struct {
    const artsample_t *in;      /* first sample of the channel; host or device memory as the call says        */
    unsigned char *out;         /* first output byte of the channel                                            */
    int   inStride, outStride;  /* floats between samples; BYTES between output samples                        */
    int   frames, context;
    int   bits, bytes, pad;     /* outputBits, outputBytes, leading zero bytes of the container                */
    artsample_t scaler;         /* (1 << bits) / 2 * gain                                                      */
    int   dither, ditherType, shaping;
    unsigned int rng;           /* tpdf generator                                                              */
    artsample_t feedback;
    artsample_t a[5], b[5], x[4], y[4];   /* noise shaper: coefficients, delayed input / output newest first          */
    int   order;
    int   clips;                /* out: samples clipped                                                        */
} ArtDecLane;

/* lanes: host array (read and updated); channelsHint: numChannels of the first context.  Host buffers are staged through
 * device memory laid out per lane.  Returns non-zero on failure. */
int artDecimateRun (ArtDecLane *lanes, int numLanes, int numContexts, int channelsHint, int onDevice, void *stream);
int artFloatIntegersRun (const unsigned char *input, double gain, int bits, int bytes, int stride, artsample_t *output, int count,
                         int onDevice, void *stream);

#ifdef __cplusplus
}
#endif
#endif
