/* tests/shim/plan_shim.c -- exposes the product's closed-form planner (csrc/art_plan.h) and a
 * host-only emulation of plan_call() (csrc/art_context.c) so that the scalar bookkeeping can be
 * checked against the reference on a machine without a GPU.  Test infrastructure only. */
#include "../../audio-resampler_b200/csrc/art_plan.h"

typedef struct { double P; int I; int T; int F; int flushed; int snap; } ShimState;
typedef struct { unsigned int used, made; } ShimResult;

ShimResult shim_call (ShimState *s, int numIn, int numOut, double ratio)
{
    const int half = s->T / 2, NS = 16 * s->T, D = 15 * s->T;
    ArtLoopState st;
    ArtLoopPlan lp;
    ShimResult r;

    if (s->flushed) numIn = 0;
    if (numIn < 0) {
        if (NS - s->I < half) { s->P -= D; s->I -= D; }
        s->flushed = 1;
        s->I += half;
        numIn = 0;
    }
    st.P = s->P; st.I = s->I; st.T = s->T; st.ratio = ratio;
    lp = art_plan_loop (&st, numIn, numOut);
    s->P = lp.P_after;
    s->I = lp.I_after;
    if (s->snap) {
        double whole = floor (s->P);
        s->P = whole + floor ((s->P - whole) * s->F + 0.5) / s->F;
    }
    r.used = lp.inputs; r.made = lp.outputs;
    return r;
}

/* per-output positions, as the kernels derive them */
double shim_output_pos (const ShimState *s, double ratio, unsigned int n, int *wraps)
{
    ArtLoopState st;
    st.P = s->P; st.I = s->I; st.T = s->T; st.ratio = ratio;
    return art_output_pos (&st, n, wraps);
}

/* ... and as the threads of a tile whose first output is n0 derive them */
double shim_output_pos_from (const ShimState *s, double ratio, unsigned int n0, unsigned int n, int *wraps)
{
    ArtLoopState st;
    int w0;
    st.P = s->P; st.I = s->I; st.T = s->T; st.ratio = ratio;
    (void) art_output_pos (&st, n0, &w0);
    return art_output_pos_from (&st, n, w0, art_ring_base (st.P, st.T, w0), wraps);
}
