/*
 * art_extrapolate.h -- LPC endpoint extrapolation (EXTRAPOLATE_ENDPOINTS), host side.
 *
 * Replaces extrapolator.c of the reference (extrapolate_forward :22-43, extrapolate_reverse :49-65,
 * calc_lpc_coeffs :93-240).  It is a few hundred samples of strictly serial work per stream end, so it
 * stays on the host; art_context.c moves the synthesised samples into the device-side history (stream
 * start) or feeds them as the flush block (stream end).
 */
#ifndef ART_EXTRAPOLATE_H
#define ART_EXTRAPOLATE_H
#include "art_sample.h"

#ifdef __cplusplus
extern "C" {
#endif

/* x[0 .. known) are given; writes x[known .. known + more) */
void artExtendForward (artsample_t *x, int known, int more);
/* end[-1] (newest) ... end[-known] are given; writes end[-known-1] ... end[-known-more] */
void artExtendBackward (artsample_t *end, int known, int more);

#ifdef __cplusplus
}
#endif
#endif
