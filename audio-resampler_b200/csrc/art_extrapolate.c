/*
 * art_extrapolate.c -- see art_extrapolate.h.
 *
 * The predictor is fitted exactly as the reference fits it: every sum below is taken in the
 * reference's order and width (float products widened into double accumulators), because the fit is
 * a sequence of discrete accept/reject decisions -- a differently rounded error sum can take another
 * branch and end on different coefficients.  Built with -ffp-contract=off like the rest of the host code.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "art_extrapolate.h"

enum { ORDER = 4, ROUND_LIMIT = 100000 };          /* NCOEFFS, MAXLOOPS: extrapolator.h:26-32 */

typedef struct {
    const artsample_t *x;  /* the known samples                       */
    int evals;             /* samples that have ORDER predecessors     */
    float c[ORDER];        /* c[0] weighs the newest predecessor       */
} Fit;

/* the prediction filter applied to the ORDER samples starting at w: sum_j c[ORDER-1-j] * w[j] */
static double predict (const float *c, const artsample_t *w)
{
    double acc = 0.0;
    int j;
    for (j = 0; j < ORDER; ++j)
        acc += c[ORDER - 1 - j] * w[j];
    return acc;
}

static double residual_energy (const Fit *f)       /* extrapolator.c:198-209 */
{
    double total = 0.0;
    int k;
    for (k = 0; k < f->evals; ++k) {
        const double r = predict (f->c, f->x + k) + f->x[k + ORDER];
        total += r * r;
    }
    return total;
}

/* direct-form coefficients -> reflection coefficients (extrapolator.c:244-272), step-down recursion */
static void step_down (const double *a, double *kappa)
{
    double now[ORDER], then[ORDER];
    int m, i;
    memcpy (now, a, sizeof now);
    for (m = ORDER - 1; m >= 0; --m) {
        double den;
        kappa[m] = now[m];
        den = 1.0 - kappa[m] * kappa[m];
        if (fabs (den) < 1e-6) {
            kappa[m] = kappa[m] < 0.0 ? -0.9999995 : 0.9999995;
            den = 1.0 - kappa[m] * kappa[m];
        }
        for (i = 0; i < m; ++i)
            then[i] = (now[i] - kappa[m] * now[m - i - 1]) / den;
        for (i = 0; i < m; ++i)
            now[i] = then[i];
    }
}

/* reflection coefficients -> direct form (extrapolator.c:276-283), step-up recursion */
static void step_up (const double *kappa, double *a)
{
    int i, j;
    for (i = 0; i < ORDER; ++i) {
        a[i] = kappa[i];
        for (j = 0; j < i / 2; ++j) {
            const double held = a[j];
            a[j] += kappa[i] * a[i - 1 - j];
            a[i - 1 - j] += kappa[i] * held;
        }
        if (i & 1)
            a[i >> 1] += a[i >> 1] * kappa[i];
    }
}

static void fit (Fit *f)
{
    double energy = 0.0, diff_energy = 0.0, best, stride = 3.0 / 16.0, *r;
    int rounds = 0, accepted = 0, k;

    memset (f->c, 0, sizeof f->c);
    for (k = 0; k < f->evals; ++k) {                /* extrapolator.c:101-107 */
        const artsample_t s = f->x[k + ORDER], p = f->x[k + ORDER - 1];
        diff_energy += (s - p) * (s - p);
        energy += s * s;
    }
    if (energy == 0.0)                              /* :109-112 */
        return;

    r = malloc (sizeof (double) * (f->evals > 0 ? f->evals : 1));
    best = energy;
    while (best > 0.0 && rounds < ROUND_LIMIT) {    /* :118 */
        int tap;
        for (k = 0; k < f->evals; ++k)              /* :121-129 */
            r[k] = predict (f->c, f->x + k) + f->x[k + ORDER];
        /* try each coefficient one stride down and up; take the first that helps (:131-154) */
        for (tap = 0; rounds++, tap < ORDER; tap++) {
            double lower = 0.0, upper = 0.0;
            for (k = 0; k < f->evals; ++k) {
                const double d = f->x[k + ORDER - tap - 1] * stride;
                lower += (r[k] - d) * (r[k] - d);
                upper += (r[k] + d) * (r[k] + d);
            }
            if (lower < best || upper < best) {
                if (lower < upper) { best = lower; f->c[tap] -= stride; }
                else               { best = upper; f->c[tap] += stride; }
                ++accepted;
                break;
            }
        }
        if (tap == ORDER) {                         /* nothing helped at this stride (:158-163) */
            if (stride > 3.0 / (1 << 22)) stride *= 0.5;
            else break;
        }
    }
    free (r);

    if (accepted) {                                 /* keep the synthesis filter stable (:170-194) */
        double a[ORDER], kappa[ORDER];
        int i, clipped = 0;
        for (i = 0; i < ORDER; ++i) a[i] = f->c[i];
        step_down (a, kappa);
        for (i = 0; i < ORDER; ++i)
            if (fabs (kappa[i]) > 0.9999) { kappa[i] = kappa[i] < 0.0 ? -0.9999 : 0.9999; ++clipped; }
        if (clipped) {
            step_up (kappa, a);
            for (i = 0; i < ORDER; ++i) f->c[i] = (float) a[i];
        }
    }

    {                                               /* fall back to "hold" or to silence (:213-222) */
        const double err = residual_energy (f);
        if (diff_energy < err && diff_energy < energy) {
            memset (f->c, 0, sizeof f->c);
            f->c[0] = -1.0f;
        }
        else if (energy <= err)
            memset (f->c, 0, sizeof f->c);
    }
}

void artExtendForward (artsample_t *x, int known, int more)
{
    Fit f;
    int i;
    memset (x + known, 0, sizeof (artsample_t) * (size_t) more);       /* extrapolator.c:29 */
    f.x = x;
    f.evals = known - ORDER;
    fit (&f);
    for (i = 0; i < more; ++i)                                     /* :32-40 */
        x[known + i] = (artsample_t) -predict (f.c, x + known - ORDER + i);
}

void artExtendBackward (artsample_t *end, int known, int more)
{
    artsample_t *mirror = calloc ((size_t) known + (size_t) more, sizeof (artsample_t));
    int i;
    for (i = 0; i < known; ++i)
        mirror[i] = end[-1 - i];
    artExtendForward (mirror, known, more);
    for (i = known; i < known + more; ++i)
        end[-1 - i] = mirror[i];
    free (mirror);
}
