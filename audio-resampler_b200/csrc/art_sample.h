/*
 * art_sample.h -- the sample type of a build.  Like the reference (resampler.h:22-26, biquad.h:21-25, decimator.h:23-27) the
 * sources compile twice: plain for the 32-bit float path (libresampler_b200.so) and with -DPATH_WIDTH=64 for the path on which
 * every sample, filter tap and piece of filter state is a double (libresampler_b200_64.so; the reference's art64 / artest64).
 * The wide build keeps to the any-ratio kernel (double multiply-adds on the FP64 pipe); the packed-FP32 and tensor-core forms are
 * float-path optimisations and are compiled out of it.
 */
#ifndef ART_SAMPLE_H
#define ART_SAMPLE_H

#if defined(PATH_WIDTH) && (PATH_WIDTH==64)
typedef double artsample_t;
#define ART_WIDE 1
#else
typedef float artsample_t;
#define ART_WIDE 0
#endif

#endif
