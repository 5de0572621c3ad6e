"""GPU parity tests of the fused pre-filter (resampleB200AttachPrefilter, include/resampler_b200.h): the cascade of biquad
lowpass sections art.c runs over the input in front of a downsampling resampler (art.c:848-851, :1011-1017) folded into the
context's filter bank.  Checked against the COMPOSITION the reference performs -- the oracle's biquad recurrence over the
input (state carried from call to call), then the oracle's resampler -- with the usual bar: counts and position identical,
samples within 1e-6 of the oracle's peak."""
import ctypes as C

import numpy as np
import pytest

import artlibs as A

pytestmark = pytest.mark.gpu
TOL = 1e-6
BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS


class OCo(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class OBq(C.Structure):
    _fields_ = [("a", C.c_float * 5), ("b", C.c_float * 5), ("xh", C.c_float * 4), ("yh", C.c_float * 4),
                ("order", C.c_int), ("cursor", C.c_int)]


@pytest.fixture
def lib():
    lib = A.product()
    yield lib
    lib.resampleB200SetTensorPath(1)


def _sections(lib, freq, count, gain=1.0):
    co = A.BiquadCoefficients()
    lib.biquad_lowpass(C.byref(co), freq)
    arr = (A.Biquad * count)()
    for q in arr:
        lib.biquad_init(C.byref(q), C.byref(co), gain)
    return co, arr


class OraclePrefilter:
    """the reference's side of the composition: `count` sections per channel, applied in place to every block (art.c:1011-1017)"""

    def __init__(self, co, count, channels, gain=1.0):
        self.ol, self.ch = A.oracle(), channels
        oco = OCo()
        for n, _ in OCo._fields_:
            setattr(oco, n, getattr(co, n))
        self.q = [[OBq() for _ in range(channels)] for _ in range(count)]
        for st in self.q:
            for q in st:
                self.ol.oracle_biquad_init(C.byref(q), C.byref(oco), C.c_double(gain))

    def run(self, x):
        y = np.ascontiguousarray(x, dtype=np.float32).copy()
        if y.shape[0]:
            for st in self.q:
                for c in range(self.ch):
                    self.ol.oracle_biquad_run(C.byref(st[c]), y[:, c:].ctypes.data_as(A.f32p), y.shape[0], self.ch)
        return y


@pytest.mark.parametrize("kernel,ch,preset,src,dst,sections", [
    ("tensor", 2, 3, 96000, 44100, 2),      # art's downsampling pre-filter, two sections at 0.45 * dst/src
    ("ffma", 2, 3, 96000, 44100, 2),
    ("generic", 3, 2, 96000, 44117.3, 2),   # a ratio with no small-numerator fraction: any-ratio kernel
    ("tensor", 8, 4, 96000, 44100, 2),      # BASELINE config 3's shape (8 of its 64 channels), 988 + 32 taps, planar scratch
    ("ffma", 1, 1, 48000, 40000, 1),        # one section, short filter (5/6)
    ("tensor", 4, 2, 48000, 44100, 3),      # three sections at 0.41: a longer response (more lead taps)
])
def test_fused_prefilter_equals_biquads_then_resampler(lib, kernel, ch, preset, src, dst, sections):
    lib.resampleB200SetTensorPath({"tensor": 2, "ffma": 0, "generic": 0}[kernel])
    filters, taps = A.PRESETS[preset]
    ratio = dst / src
    lowpass = min(0.95, 0.9 * ratio)
    g = A.product_stream(ch, taps, filters, lowpass_ratio=lowpass, flags=BH_INTERP)
    o = A.oracle_stream(ch, taps, filters, lowpass_ratio=lowpass, flags=BH_INTERP)
    g.advance(taps / 2); o.advance(taps / 2)
    co, arr = _sections(lib, 0.45 * dst / src, sections)
    assert lib.resampleB200AttachPrefilter(g.ctx, arr, sections) == 0
    pre = OraclePrefilter(co, sections, ch)
    rng = np.random.default_rng(500 + ch + preset)
    before = A.path_counts(lib)
    sizes = [9000, 1, 12000, 0, 7001]
    for b, n in enumerate(sizes):
        x = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
        last = b == len(sizes) - 1
        yg, ug, mg = g.process(x, int(n * ratio) + taps + 64, ratio, flush_after=last)
        yo, uo, mo = o.process(pre.run(x), int(n * ratio) + taps + 64, ratio, flush_after=last)
        assert (ug, mg) == (uo, mo), f"block {b}: counts {(ug, mg)} vs {(uo, mo)}"
        assert g.position() == o.position()
        err = A.peak_error(yg, yo)
        assert err <= TOL, f"block {b}: max|d|/peak = {err:.3g}"
    after = A.path_counts(lib)
    used = [a - b for a, b in zip(after, before)]
    assert used[{"generic": 0, "ffma": 1, "tensor": 2}[kernel]] >= 2, f"expected the {kernel} kernel, launches {used}"


def test_prefilter_through_the_batched_device_api(lib):
    """64 lock-step contexts with the pre-filter attached, one launch per block (the shape bench.py times for config 3/4)"""
    lib.resampleB200SetTensorPath(2)
    ch, taps, filters, ratio = 2, 380, 380, 44100 / 48000
    n = 24
    gs = [A.product_stream(ch, taps, filters, lowpass_ratio=0.9 * ratio, flags=BH_INTERP) for _ in range(n)]
    os_ = [A.oracle_stream(ch, taps, filters, lowpass_ratio=0.9 * ratio, flags=BH_INTERP) for _ in range(n)]
    pres = []
    for g, o in zip(gs, os_):
        g.advance(taps / 2); o.advance(taps / 2)
        co, arr = _sections(lib, 0.45 * ratio, 2)
        assert lib.resampleB200AttachPrefilter(g.ctx, arr, 2) == 0
        pres.append(OraclePrefilter(co, 2, ch))
    rng = np.random.default_rng(77)
    for step in range(2):
        xs = [rng.uniform(-0.5, 0.5, (7000, ch)).astype(np.float32) for _ in range(n)]
        got = A.device_batch_process(gs, xs, 8000, ratio)
        for i in range(n):
            yo, uo, mo = os_[i].process(pres[i].run(xs[i]), 8000, ratio)
            y, u, m = got[i]
            assert (u, m) == (uo, mo) and gs[i].position() == os_[i].position()
            assert A.peak_error(y, yo) <= TOL


def test_prefilter_refusals(lib, capfd):
    g = A.product_stream(2, 48, 48, 0.0)
    co, arr = _sections(lib, 0.2, 2)
    x = np.zeros((10, 2), np.float32)
    g.process(x, 100, 1.5)
    assert lib.resampleB200AttachPrefilter(g.ctx, arr, 2) != 0          # the stream has consumed input
    g.reset()
    assert lib.resampleB200AttachPrefilter(g.ctx, arr, 2) == 0          # fine after a reset
    assert lib.resampleB200AttachPrefilter(g.ctx, arr, 2) != 0          # only once
    e = A.product_stream(2, 48, 48, 0.0, flags=BH_INTERP | A.EXTRAPOLATE_ENDPOINTS)
    assert lib.resampleB200AttachPrefilter(e.ctx, arr, 2) != 0          # the reference extrapolates the filtered signal
    long_ = A.product_stream(1, 988, 64, 0.0)
    co2, arr2 = _sections(lib, 0.002, 2)                                # a 0.002 fs lowpass rings for thousands of samples
    assert lib.resampleB200AttachPrefilter(long_.ctx, arr2, 2) != 0
    err = capfd.readouterr().err
    assert "before the first input" in err and "already attached" in err and "EXTRAPOLATE_ENDPOINTS" in err and "too long" in err
