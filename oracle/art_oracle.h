/*
 * art_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99, scalar, single-threaded) of the windowed-sinc
 * resampling hot path of dbry/audio-resampler and of its order-2 biquad
 * lowpass.  It exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py have something to check the CUDA path against.
 * Nothing under audio-resampler_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every
 * function here with the unmodified reference compiled into
 * oracle/_ref/libartref.so (recipe: oracle/Makefile), and tests/golden/ holds
 * vectors generated from that reference build (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef ART_ORACLE_H
#define ART_ORACLE_H

/* the sample type, as in the reference (resampler.h:22-26): liboracle.so is the float path, liboracle64.so (-DPATH_WIDTH=64) the
 * path on which samples, taps and filter state are doubles */
#if defined(PATH_WIDTH) && (PATH_WIDTH==64)
typedef double osample_t;
#else
typedef float osample_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* flag values: resampler.h:28-38 */
#define ORC_INTERPOLATE      0x1
#define ORC_BLACKMAN_HARRIS  0x2
#define ORC_LOWPASS          0x4
#define ORC_MULTITHREADED    0x8     /* accepted, ignored (single-threaded oracle) */
#define ORC_NO_REDUCTION     0x10
#define ORC_FIXED_RATIO      0x20
#define ORC_EXTRAPOLATE      0x40
#define ORC_PREFILL_PENDING  0x80
#define ORC_EXTENDED_MATH    0x100
#define ORC_FLUSHED          0x200
#define ORC_SNAP             0x400

typedef struct {
    unsigned int input_used, output_generated;
} OracleResult;

typedef struct OracleResampler {
    int     channels, taps, phases, flags;
    int     ring_len;          /* 16 * taps, resampler.c:139                     */
    int     write_index;       /* "inputIndex"                                   */
    double  read_pos;          /* "outputOffset"                                 */
    double  fixed_ratio, lowpass_ratio;
    osample_t  *bank;              /* (phases + 1) rows of taps floats, contiguous   */
    osample_t  *ring;              /* channels rows of ring_len floats, contiguous   */
} OracleResampler;

OracleResampler *oracle_init (int channels, int taps, int phases, double lowpass_ratio, int flags);
OracleResampler *oracle_fixed_ratio_init (int channels, int taps, int max_phases, double src_rate,
                                          double dst_rate, int lowpass_hz, int flags);
void   oracle_free (OracleResampler *r);
void   oracle_reset (OracleResampler *r);
void   oracle_advance (OracleResampler *r, double delta);
double oracle_position (const OracleResampler *r);
const osample_t *oracle_bank_row (const OracleResampler *r, int row);

/* One strided core serves the planar and the interleaved entry points of the
 * reference: sample (frame f, channel c) lives at base_c[f * frame_stride],
 * with base_c = in[c] (planar, frame_stride 1) or in + c (interleaved,
 * frame_stride = channels).  n_in < 0 requests a flush. */
OracleResult oracle_process_interleaved (OracleResampler *r, const osample_t *in, int n_in,
                                         osample_t *out, int n_out, double ratio);
OracleResult oracle_process_planar (OracleResampler *r, const osample_t *const *in, int n_in,
                                    osample_t *const *out, int n_out, double ratio);
OracleResult oracle_process_flush_interleaved (OracleResampler *r, const osample_t *in, int n_in,
                                               osample_t *out, int n_out, double ratio);
OracleResult oracle_process_flush_planar (OracleResampler *r, const osample_t *const *in, int n_in,
                                          osample_t *const *out, int n_out, double ratio);
unsigned int oracle_required_input (const OracleResampler *r, int n_out, double ratio);
unsigned int oracle_expected_output (const OracleResampler *r, int n_in, double ratio);

/* biquad.h:27-35 */
typedef struct { osample_t a0, a1, a2, a3, a4, b1, b2, b3, b4; } OracleBiquadCoeffs;
typedef struct {
    osample_t a[5], b[5];
    osample_t xh[4], yh[4];
    int   order, cursor;
} OracleBiquad;

void oracle_biquad_lowpass (OracleBiquadCoeffs *c, double frequency);
void oracle_biquad_highpass (OracleBiquadCoeffs *c, double frequency);
void oracle_biquad_init (OracleBiquad *q, const OracleBiquadCoeffs *c, double gain);
void oracle_biquad_run (OracleBiquad *q, osample_t *buf, int count, int stride);

/* decimator.c: float -> integer with TPDF dither and noise shaping (decimateInit :29-100, decimateProcess*LE :112-291),
 * and its lossless inverse floatIntegersLE (:416-450).  Flag values are decimator.h:29-41. */
typedef struct {
    unsigned int rng;           /* "tpdf_generators[ch]" */
    osample_t        feedback;      /* "feedback[ch]"        */
    OracleBiquad shaper;        /* "noise_shapers[ch]"   */
} OracleDecimatorLane;

typedef struct {
    int    channels, bits, bytes, flags, dither;
    double gain;
    OracleDecimatorLane *lane;
} OracleDecimator;

OracleDecimator *oracle_decimate_init (int channels, int bits, int bytes, double gain, int rate, int flags);
void oracle_decimate_free (OracleDecimator *d);
int  oracle_decimate_interleaved (OracleDecimator *d, const osample_t *in, int frames, unsigned char *out);
int  oracle_decimate_planar (OracleDecimator *d, const osample_t *const *in, int frames, unsigned char *const *out);
void oracle_float_integers (const unsigned char *in, double gain, int bits, int bytes, int stride, osample_t *out, int count);

/* artest.c:744-754 -- the reference's synthetic noise generator (state passed explicitly). */
void oracle_noise (unsigned long long *state, osample_t *dst, int count);

/* extrapolator.c:22-65, exported for the tests */
void oracle_extend_forward (osample_t *x, int known, int more);
void oracle_extend_backward (osample_t *end, int known, int more);

#ifdef __cplusplus
}
#endif
#endif
