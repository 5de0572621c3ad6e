"""CPU tests: pin the oracle (oracle/art_oracle.c) to the reference.

* against the committed golden vectors (generated from the unmodified reference by
  tests/golden/make_golden.py) -- these run everywhere;
* against oracle/_ref/libartref.so directly when it is present.
Counts and positions must be bit-identical; samples within 1e-6 of peak (the reference's own
build-to-build noise floor is 1.2e-7).
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import artlibs as A

GOLDEN = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLDEN))
import make_golden  # noqa: E402

TOL = 1e-6


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c["name"] for c in make_golden.CASES])
def test_oracle_matches_golden(case):
    g = np.load(GOLDEN / f"{case['name']}.npz")
    out, meta = make_golden.run_case(case, A.oracle_stream)
    assert np.array_equal(meta, g["meta"]), "input_used / output_generated / position differ from the reference"
    assert out.shape == g["out"].shape
    assert A.peak_error(out, g["out"]) <= TOL


def test_passthrough_golden_is_exact():
    """2x upsampling without lowpass returns every input sample verbatim (resampler.c:1141-1142)."""
    case = next(c for c in make_golden.CASES if c["name"] == "fixed_mono_p1_x2_passthrough")
    out, _ = make_golden.run_case(case, A.oracle_stream)
    rng = np.random.default_rng(case["seed"])
    x = rng.uniform(-0.5, 0.5, (case["calls"][0], 1)).astype(np.float32)
    assert np.array_equal(out[0:2 * len(x):2], x)


def test_bank_known_answers():
    """SURVEY.md 8c known-answer structure of the 48x48 Blackman-Harris bank."""
    g = np.load(GOLDEN / "bank_48x48_bh.npz")["bank"]
    s = A.oracle_stream(1, 48, 48, 0.0)
    bank = s.bank()
    assert bank.shape == (49, 48)
    assert np.max(np.abs(bank - g)) < 1e-12            # summation order of the -fassociative-math build
    assert bank[0, 23] == 1.0 and np.max(np.abs(np.delete(bank[0], 23))) < 1e-15
    assert np.allclose(bank[24, 20:28], [-0.08048406, 0.11964639, -0.20751575, 0.6350419,
                                         0.6350419, -0.20751576, 0.11964639, -0.08048406], atol=1e-7)
    assert np.array_equal(bank[48, 2:], bank[0, 1:-1]) and bank[48, 0] == 0.0 and bank[0, 47] == 0.0
    assert np.max(np.abs(bank[:48].sum(axis=1, dtype=np.float64) - 1.0)) < 1e-12


def test_biquad_design_known_answer():
    g = np.load(GOLDEN / "biquad_lowpass_0p2067.npz")["coeffs"]
    lib = A.oracle()

    class Co(C.Structure):
        _fields_ = [(n, C.c_float) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]
    co = Co()
    lib.oracle_biquad_lowpass(C.byref(co), C.c_double(0.45 * 44100 / 96000))
    got = np.array([co.a0, co.a1, co.a2, co.b1, co.b2], np.float32)
    assert np.array_equal(got, g)
    assert np.allclose(got, [0.21753205, 0.43506411, 0.21753205, -0.31955418, 0.18968238], atol=1e-7)


def test_oracle_biquad_cascade_matches_golden():
    g = np.load(GOLDEN / "biquad_cascade_3ch.npz")["out"]
    lib = A.oracle()

    class Co(C.Structure):
        _fields_ = [(n, C.c_float) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]

    class Bq(C.Structure):
        _fields_ = [("a", C.c_float * 5), ("b", C.c_float * 5), ("xh", C.c_float * 4), ("yh", C.c_float * 4),
                    ("order", C.c_int), ("cursor", C.c_int)]
    co = Co()
    lib.oracle_biquad_lowpass(C.byref(co), C.c_double(0.45 * 44100 / 96000))
    rng = np.random.default_rng(21)
    y = rng.uniform(-0.5, 0.5, (5000, 3)).astype(np.float32)
    stages = [[Bq() for _ in range(3)] for _ in range(2)]
    for st in stages:
        for q in st:
            lib.oracle_biquad_init(C.byref(q), C.byref(co), C.c_double(1.0))
    for lo, hi in [(0, 1234), (1234, 5000)]:
        for c in range(3):
            for st in stages:
                lib.oracle_biquad_run(C.byref(st[c]), y[lo:, c:].ctypes.data_as(A.f32p), hi - lo, 3)
    assert A.peak_error(y, g) <= TOL


@pytest.mark.skipif(A.reference() is None, reason="oracle/_ref/libartref.so not built")
def test_oracle_vs_live_reference_random_sessions():
    rng = np.random.default_rng(5)
    for trial in range(12):
        ch = int(rng.integers(1, 4))
        filters, taps = A.PRESETS[int(rng.integers(1, 4))]
        fixed = rng.random() < 0.3
        if fixed:
            src, dst = [(44100, 48000), (48000, 44100), (22050, 44100), (48000, 32000)][int(rng.integers(0, 4))]
            kw = dict(flags=7, fixed=(src, dst, 0))
            ratio = 0.0
        else:
            ratio = float(np.exp(rng.uniform(np.log(0.3), np.log(3.0))))
            kw = dict(lowpass_ratio=float(rng.choice([0.0, 0.8])), flags=int(rng.choice([1, 3, 2, 0x103])))
        o = A.oracle_stream(ch, taps, filters, **kw)
        r = A.reference_stream(ch, taps, filters, **kw)
        if kw.get("flags", 3) & 1 or fixed:
            o.advance(taps / 2)
            r.advance(taps / 2)
        for call in range(5):
            n = int(rng.integers(0, 3000))
            x = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
            cap = int(rng.integers(0, 4000))
            planar = bool(rng.integers(0, 2))
            yo, uo, go = o.process(x, cap, ratio, planar=planar)
            yr, ur, gr = r.process(x, cap, ratio)
            assert (uo, go) == (ur, gr)
            assert o.position() == r.position()
            assert A.peak_error(yo, yr) <= TOL
        yo, uo, go = o.process(None, 5000, ratio)
        yr, ur, gr = r.process(None, 5000, ratio)
        assert (uo, go) == (ur, gr) and A.peak_error(yo, yr) <= TOL
        assert o.required_input(1000, ratio or 1.0) == r.required_input(1000, ratio or 1.0)
        o.reset(); r.reset()
        assert o.expected_output(1000, ratio or 1.0) == r.expected_output(1000, ratio or 1.0)


def test_artest_noise_generator_known_prefix():
    """artest.c:744-754 with the reference's seed; first values recorded from the reference build."""
    x, _ = A.artest_noise(4)
    assert np.allclose(x, [0.36142448, -0.19244033, -0.4861091, 0.38178906], atol=1e-8)


def test_oracle_extrapolation_matches_the_reference_build():
    """extrapolator.c restated in oracle/art_oracle.c: bit-identical synthesised samples on random, tonal and drifting
    inputs, forwards and backwards; and whole streams with EXTRAPOLATE_ENDPOINTS agree with the reference build."""
    import ctypes as C
    ref = A.reference()
    if ref is None:
        pytest.skip("oracle/_ref/libartref.so is not available")
    orc = A.oracle()
    f32p = C.POINTER(C.c_float)
    orc.oracle_extend_forward.argtypes = [f32p, C.c_int, C.c_int]; orc.oracle_extend_forward.restype = None
    orc.oracle_extend_backward.argtypes = [f32p, C.c_int, C.c_int]; orc.oracle_extend_backward.restype = None
    ref.extrapolate_forward.argtypes = [f32p, C.c_int, C.c_int]; ref.extrapolate_forward.restype = C.c_double
    ref.extrapolate_reverse.argtypes = [f32p, C.c_int, C.c_int]; ref.extrapolate_reverse.restype = C.c_double
    rng = np.random.default_rng(0)
    for trial in range(60):
        known, more = int(rng.integers(8, 500)), int(rng.integers(1, 500))
        t = np.arange(known + more)
        x = [rng.uniform(-0.5, 0.5, known + more), 0.4 * np.sin(2 * np.pi * t * rng.uniform(0.001, 0.3)),
             np.cumsum(rng.normal(0, 0.01, known + more))][trial % 3].astype(np.float32)
        a, b = x.copy(), x.copy()
        orc.oracle_extend_forward(a.ctypes.data_as(f32p), known, more)
        ref.extrapolate_forward(b.ctypes.data_as(f32p), known, more)
        assert np.array_equal(a, b)
        a, b = x.copy(), x.copy()
        orc.oracle_extend_backward(C.cast(a.ctypes.data + 4 * (known + more), f32p), known, more)
        ref.extrapolate_reverse(C.cast(b.ctypes.data + 4 * (known + more), f32p), known, more)
        assert np.array_equal(a, b)
    flags = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS | A.EXTRAPOLATE_ENDPOINTS
    for ch, taps, filters, ratio in [(2, 380, 380, 48000 / 44100), (1, 48, 48, 0.77)]:
        o = A.oracle_stream(ch, taps, filters, 0.0, flags=flags)
        r = A.reference_stream(ch, taps, filters, 0.0, flags=flags)
        o.advance(taps / 2); r.advance(taps / 2)
        for b, n in enumerate([3000, 5, 4096]):
            x = (0.3 * np.sin(0.07 * np.arange(n)[:, None] + np.arange(ch)) + rng.normal(0, 0.02, (n, ch))).astype(np.float32)
            yo, uo, mo = o.process(x, 20000, ratio, flush_after=(b == 2))
            yr, ur, mr = r.process(x, 20000, ratio, flush_after=(b == 2))
            assert (uo, mo) == (ur, mr) and A.peak_error(yo, yr) <= 3e-7


# ---------------------------------------------------------------------------------------------- decimator.c
DECIMATE_FLAGS = [0, 0x1, 0x2, 0x4, 0x100, 0x200, 0x400, 0x800, 0x2 | 0x800, 0x1 | 0x200, 0x4 | 0x100]


def _bind_decimators():
    ol, ref = A.oracle(), A.reference()
    ol.oracle_decimate_init.restype = C.c_void_p
    ol.oracle_decimate_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    ol.oracle_decimate_free.argtypes = [C.c_void_p]
    ol.oracle_decimate_interleaved.argtypes = [C.c_void_p, A.f32p, C.c_int, C.c_char_p]
    ol.oracle_decimate_interleaved.restype = C.c_int
    ol.oracle_float_integers.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, A.f32p, C.c_int]
    ol.oracle_float_integers.restype = None
    if ref is not None:
        ref.decimateInit.restype = C.c_void_p
        ref.decimateInit.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        ref.decimateFree.argtypes = [C.c_void_p]
        ref.decimateProcessInterleavedLE.argtypes = [C.c_void_p, A.f32p, C.c_int, C.c_char_p]
        ref.decimateProcessInterleavedLE.restype = C.c_int
        ref.decimateProcessLE.argtypes = [C.c_void_p, C.POINTER(A.f32p), C.c_int, C.POINTER(C.c_char_p)]
        ref.decimateProcessLE.restype = C.c_int
        ref.floatIntegersLE.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, A.f32p, C.c_int]
        ref.floatIntegersLE.restype = None
    return ol, ref


@pytest.mark.skipif(A.reference() is None, reason="oracle/_ref/libartref.so not built")
@pytest.mark.parametrize("flags", DECIMATE_FLAGS)
def test_oracle_decimator_is_bit_identical_to_the_reference(flags):
    """decimateInit / decimateProcessInterleavedLE (decimator.c:29-100, :205-291) in every dither / shaping mode, 8-, 16- and
    24-bit output (24 in a 32-bit container too), gains that clip, two calls in a row (state carries over)"""
    ol, ref = _bind_decimators()
    rng = np.random.default_rng(flags + 5)
    for ch, bits, bytes_, gain, rate in [(2, 16, 2, 1.0, 44100), (1, 8, 1, 0.9, 48000), (3, 24, 3, 1.3, 96000), (2, 24, 4, 1.0, 32000), (5, 12, 2, 2.5, 12345)]:
        r = ref.decimateInit(ch, bits, bytes_, gain, rate, flags)
        o = ol.oracle_decimate_init(ch, bits, bytes_, gain, rate, flags)
        for n in (777, 1, 2500):
            x = rng.uniform(-1.0, 1.0, (n, ch)).astype(np.float32)
            br, bo = C.create_string_buffer(n * ch * bytes_ + 8), C.create_string_buffer(n * ch * bytes_ + 8)
            cr = ref.decimateProcessInterleavedLE(r, x.ctypes.data_as(A.f32p), n, br)
            co = ol.oracle_decimate_interleaved(o, x.ctypes.data_as(A.f32p), n, bo)
            assert cr == co, (ch, bits, flags, "clipped sample counts differ")
            assert br.raw == bo.raw, (ch, bits, flags, "bytes differ")
        ref.decimateFree(r); ol.oracle_decimate_free(o)


@pytest.mark.skipif(A.reference() is None, reason="oracle/_ref/libartref.so not built")
def test_oracle_float_integers_is_bit_identical_to_the_reference():
    ol, ref = _bind_decimators()
    rng = np.random.default_rng(9)
    for bits, bytes_, stride, gain in [(8, 1, 1, 1.0), (16, 2, 1, 1.0), (16, 2, 3, 0.5), (24, 3, 1, 1.0), (24, 4, 2, 1.7), (20, 3, 1, 1.0), (12, 2, 1, 1.0)]:
        n = 1000
        raw = rng.integers(0, 256, n * stride * bytes_ + 16, dtype=np.uint8).tobytes()
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        ref.floatIntegersLE(raw, gain, bits, bytes_, stride, a.ctypes.data_as(A.f32p), n)
        ol.oracle_float_integers(raw, gain, bits, bytes_, stride, b.ctypes.data_as(A.f32p), n)
        assert np.array_equal(a, b), (bits, bytes_, stride)
