"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libresampler_b200.so (host-pointer entry points of include/resampler.h) and is compared with the
oracle on the same seeded inputs, and with the committed golden vectors of the reference.

Bar: input_used / output_generated / resampleGetPosition bit-identical; samples within
1e-6 * peak(reference) (BASELINE.json north_star; BASELINE.md section 2 explains why the measure is
peak-relative).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

import artlibs as A

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLDEN))
import make_golden  # noqa: E402

TOL = 1e-6
BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS


def _pair(ch, taps, filters, **kw):
    return A.product_stream(ch, taps, filters, **kw), A.oracle_stream(ch, taps, filters, **kw)


def _check_call(g, o, x, cap, ratio, **kw):
    yg, ug, gg = g.process(x, cap, ratio, **kw)
    yo, uo, go = o.process(x, cap, ratio, **kw)
    assert (ug, gg) == (uo, go), f"counts differ: gpu {(ug, gg)} oracle {(uo, go)}"
    assert g.position() == o.position()
    err = A.peak_error(yg, yo)
    assert err <= TOL, f"max|d|/peak = {err:.3g}"
    return yg, err


@pytest.mark.parametrize("case", make_golden.CASES, ids=[c["name"] for c in make_golden.CASES])
def test_golden_vectors(case):
    g = np.load(GOLDEN / f"{case['name']}.npz")
    out, meta = make_golden.run_case(case, A.product_stream)
    assert np.array_equal(meta, g["meta"]), "input_used / output_generated / position differ from the reference"
    assert out.shape == g["out"].shape
    assert A.peak_error(out, g["out"]) <= TOL


def test_filter_bank_equals_oracle_bank():
    for preset in (1, 2, 3):
        filters, taps = A.PRESETS[preset]
        g, o = _pair(2, taps, filters, lowpass_ratio=0.0)
        assert np.array_equal(g.bank(), o.bank())
    g, o = _pair(1, 156, 320, lowpass_ratio=0.83, flags=A.SUBSAMPLE_INTERPOLATE)     # Hann + lowpass
    assert np.array_equal(g.bank(), o.bank())


@pytest.mark.parametrize("ch,preset,src,dst,lowpass_hz", [
    (1, 1, 44100, 48000, 0),            # BASELINE config 1
    (2, 3, 44100, 48000, 0),            # BASELINE config 2 (metric config)
    (64, 4, 96000, 44100, 20000),       # BASELINE config 3
    (2, 3, 48000, 44100, 20000),        # one stream of BASELINE config 4
    (3, 2, 44100, 8000, 3600),          # steep downsampling, odd channel count
    (5, 1, 8000, 48000, 0),             # steep upsampling
])
def test_baseline_configs_artest_style(ch, preset, src, dst, lowpass_hz):
    """artest.c:446-491: 4096-frame calls, advance by T/2 first, flush with the last block."""
    filters, taps = A.PRESETS[preset]
    g, o = _pair(ch, taps, filters, lowpass_ratio=lowpass_hz * 2.0 / src, flags=BH_INTERP)
    g.advance(taps / 2); o.advance(taps / 2)
    ratio = dst / src
    rng = np.random.default_rng(100 + ch)
    blocks = 3 if ch < 64 else 2
    for b in range(blocks):
        x = rng.uniform(-0.5, 0.5, (4096, ch)).astype(np.float32)
        _check_call(g, o, x, int(4096 * ratio) + taps + 10, ratio, flush_after=(b == blocks - 1))


def test_planar_equals_interleaved():
    """artest -v: the non-interleaved API must give the same samples (artest.c:658-740)."""
    g1 = A.product_stream(3, 156, 320, 0.0)
    g2 = A.product_stream(3, 156, 320, 0.0)
    rng = np.random.default_rng(2)
    for n in (1000, 1, 0, 2500):
        x = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
        y1, u1, m1 = g1.process(x, 5000, 1.25, planar=False)
        y2, u2, m2 = g2.process(x, 5000, 1.25, planar=True)
        assert (u1, m1) == (u2, m2) and np.array_equal(y1, y2)
    y1, _, _ = g1.process(None, 5000, 1.25)
    y2, _, _ = g2.process(None, 5000, 1.25, planar=True)
    assert np.array_equal(y1, y2)


def test_asrc_varying_ratio_positions():
    """BASELINE config 5: 8 channels, preset -2, ratio swept by +/-100 ppm per call; the position
    reported after every call is what an ASRC servo reads, so it must be bit-identical."""
    filters, taps = A.PRESETS[2]
    g, o = _pair(8, taps, filters, lowpass_ratio=0.0)
    g.advance(taps / 2); o.advance(taps / 2)
    rng = np.random.default_rng(6)
    K = 24
    for k in range(K):
        ratio = 1.0 + 1e-4 * np.sin(2 * np.pi * k / K)
        n = int(rng.integers(480, 4097))
        x = rng.uniform(-0.5, 0.5, (n, 8)).astype(np.float32)
        _check_call(g, o, x, n + 64, ratio)


def test_ragged_and_limited_calls():
    """empty calls, single frames, output-limited calls that leave input unconsumed, input-limited ones."""
    g, o = _pair(2, 48, 48, lowpass_ratio=0.0)
    rng = np.random.default_rng(8)
    ratio = 0.77
    plan = [(0, 10), (1, 10), (5, 0), (300, 3), (300, 1000), (7, 1000), (2000, 100), (2000, 5000), (0, 50)]
    for n, cap in plan:
        x = rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
        _check_call(g, o, x, cap, ratio)
    for _ in range(3):                     # repeated flushes: the second generates nothing
        _check_call(g, o, None, 100, ratio)
    x = rng.uniform(-0.5, 0.5, (100, 2)).astype(np.float32)
    yg, err = _check_call(g, o, x, 500, ratio)       # input after a flush is ignored until reset
    g.reset(); o.reset()
    _check_call(g, o, x, 500, ratio)


def test_flush_when_the_reference_ring_is_nearly_full():
    """A divergence from the reference that is deliberate.  postfillAllChannels (resampler.c:667-673) compacts the ring when
    fewer than T/2 slots are free by moving slots [15T, 16T) to the front -- but the newest sample then sits at
    inputIndex - 15T < T, so the windows of the flush outputs start BEFORE the buffer: the reference (and the oracle, which
    restates it) read out of bounds and return heap garbage (seen: 1e21..1e32).  The library has the true history on the
    device and returns what the reference evidently intends.  Checked against the oracle after a proper compaction of its
    ring (newest T samples to the front) right before the flush; counts and position are identical either way."""
    g, o = _pair(2, 380, 380, lowpass_ratio=0.0)
    g.advance(190); o.advance(190)
    rng = np.random.default_rng(9)
    ratio = 0.7071067811865476
    # 380 + 22658 - 3 * 5700 = 5938: 142 free slots < T/2 when the flush arrives
    for n in (9000, 13658):
        x = rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
        _check_call(g, o, x, 20000, ratio)
    A.oracle_compact_ring(o)
    assert g.position() == o.position()
    yg, ug, mg = g.process(None, 2000, ratio)
    yo, uo, mo = o.process(None, 2000, ratio)
    # the compaction above shifted the oracle's indices by another amount than the reference's own would have: positions
    # agree to rounding (1e-13), not to the bit, in this one test
    assert (ug, mg) == (uo, mo) == (0, 134) and abs(g.position() - o.position()) < 1e-9
    assert np.max(np.abs(yo)) < 2.0 and A.peak_error(yg, yo) <= TOL


def test_long_call_with_many_ring_compactions():
    """one call of 200k frames at preset -1: the reference compacts its ring ~280 times inside it."""
    g, o = _pair(2, 48, 48, lowpass_ratio=0.9)
    g.advance(24); o.advance(24)
    rng = np.random.default_rng(10)
    x = rng.uniform(-0.5, 0.5, (200_000, 2)).astype(np.float32)
    for ratio in (0.4567, 1.9):
        _check_call(g, o, x, 500_000, ratio)


def test_fixed_ratio_is_chunking_invariant_and_passthrough_exact():
    """fixed-ratio (no interpolation, snap) output is bit-invariant to call chunking;
    resampler.c:1141-1142: integer positions return the stored sample itself."""
    rng = np.random.default_rng(12)
    x = rng.uniform(-0.5, 0.5, (6000, 2)).astype(np.float32)
    outs = []
    for chunks in ([6000], [1, 999, 2000, 3000], [4096, 1904]):
        g = A.product_stream(2, 380, 380, flags=7, fixed=(44100, 48000, 0))
        assert g.num_filters() == 160 and not g.interpolation_used()
        g.advance(190)
        ys, at = [], 0
        for n in chunks:
            y, u, m = g.process(x[at:at + n], 10000, 0.0)
            assert u == n
            ys.append(y); at += n
        outs.append(np.concatenate(ys))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    g = A.product_stream(1, 48, 48, flags=3, fixed=(16000, 48000, 0))
    g.advance(24)
    m = x[:2000, :1].copy()
    y, u, made = g.process(m, 7000, 0.0)
    assert np.array_equal(y[0:3 * u:3][: (made + 2) // 3], m[: (made + 2) // 3])


def test_extended_math_and_hann():
    g, o = _pair(2, 156, 320, lowpass_ratio=0.6, flags=A.SUBSAMPLE_INTERPOLATE | A.EXTEND_CONVOLUTION_MATH)
    rng = np.random.default_rng(14)
    x = rng.uniform(-0.5, 0.5, (5000, 2)).astype(np.float32)
    _, err = _check_call(g, o, x, 9000, 0.61)
    assert err <= 2e-7          # double accumulation on both sides: only the lerp/rounding differs
    g, o = _pair(1, 48, 7, lowpass_ratio=0.0, flags=A.EXTEND_CONVOLUTION_MATH)      # no interpolation, 7 filters
    _check_call(g, o, x[:, :1].copy(), 9000, 1.3)


def test_dry_run_helpers_match_oracle():
    g, o = _pair(2, 380, 380, lowpass_ratio=0.0)
    for ratio in (48000 / 44100, 0.5, 1.0001):
        assert g.required_input(4096, ratio) == o.required_input(4096, ratio)
        assert g.expected_output(4096, ratio) == o.expected_output(4096, ratio)
        assert g.expected_output(-1, ratio) == o.expected_output(-1, ratio)


@pytest.mark.skipif(A.reference() is None, reason="oracle/_ref/libartref.so did not travel")
def test_against_live_reference_build():
    """When the compiled reference travelled with the snapshot, compare with it directly too."""
    rng = np.random.default_rng(16)
    g = A.product_stream(2, 380, 380, 0.0)
    r = A.reference_stream(2, 380, 380, 0.0, flags=BH_INTERP | A.RESAMPLE_MULTITHREADED)
    g.advance(190); r.advance(190)
    for _ in range(3):
        x = rng.uniform(-0.5, 0.5, (4096, 2)).astype(np.float32)
        yg, ug, mg = g.process(x, 8192, 48000 / 44100)
        yr, ur, mr = r.process(x, 8192, 48000 / 44100)
        assert (ug, mg) == (ur, mr) and g.position() == r.position()
        assert A.peak_error(yg, yr) <= TOL


def test_size_independent_properties_full_size():
    """BASELINE config 2 at bench size (2^20 frames in one call): properties that need no oracle run --
    linearity, DC gain of 1 (every bank row sums to 1), and agreement with a chunked run of the same
    stream (interpolated output varies by at most rounding with chunking)."""
    n, ch, ratio = 1 << 20, 2, 48000 / 44100
    rng = np.random.default_rng(18)
    a = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
    cap = int(n * ratio) + 1000

    def run(x, chunks=None):
        s = A.product_stream(ch, 380, 380, 0.0)
        s.advance(190)
        if chunks is None:
            y, u, m = s.process(x, cap, ratio)
            assert u == n
            return y
        ys, at = [], 0
        for c in chunks:
            y, u, m = s.process(x[at:at + c], cap, ratio)
            assert u == c
            ys.append(y); at += c
        return np.concatenate(ys)

    ya, yb, yab = run(a), run(b), run((a + b).astype(np.float32))
    assert ya.shape[0] == yb.shape[0] == yab.shape[0] >= int(n * ratio) - 400
    lin = np.max(np.abs(yab.astype(np.float64) - ya - yb)) / np.max(np.abs(yab))
    assert lin <= 2e-6, lin                      # three float32 roundings
    ydc = run(np.full((n, ch), 0.25, np.float32))
    assert np.max(np.abs(ydc[1000:-1000] - 0.25)) <= 1e-6
    ych = run(a, chunks=[4096] * 255 + [n - 4096 * 255])
    assert ych.shape == ya.shape and A.peak_error(ych, ya) <= 5e-7
    # oracle spot check on a window in the middle of the big call
    o = A.oracle_stream(ch, 380, 380, 0.0)
    o.advance(190)
    yo, _, _ = o.process(a[:20000], cap, ratio)
    assert A.peak_error(ya[:yo.shape[0] - 400], yo[:-400]) <= TOL


def test_kernel_selection():
    """Rational ratios with a small numerator must run on the periodic kernel, everything else on the
    generic one -- and both must agree with each other on a ratio both can take."""
    import ctypes as C
    import os
    lib = A.product()
    lib.resampleB200PathCounts.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.resampleB200SetTensorPath(1)        # default policy: calls of this size stay on the FFMA form

    def counts():
        g, p = C.c_ulonglong(), C.c_ulonglong()
        lib.resampleB200PathCounts(C.byref(g), C.byref(p))
        return g.value, p.value

    rng = np.random.default_rng(20)
    x = rng.uniform(-0.5, 0.5, (20000, 2)).astype(np.float32)
    s = A.product_stream(2, 380, 380, 0.0); s.advance(190)
    g0, p0 = counts()
    y_per, _, _ = s.process(x, 30000, 48000 / 44100)
    g1, p1 = counts()
    assert (g1 - g0, p1 - p0) == (0, 1)
    s2 = A.product_stream(2, 380, 380, 0.0); s2.advance(190)
    y_gen, _, _ = s2.process(x, 30000, 0.91234567)                 # no small-numerator fraction
    g2, p2 = counts()
    assert (g2 - g1, p2 - p1) == (1, 0)
    s3 = A.product_stream(2, 380, 380, 0.0, flags=BH_INTERP | A.EXTEND_CONVOLUTION_MATH); s3.advance(190)
    y_prec, _, _ = s3.process(x, 30000, 48000 / 44100)             # double accumulation: generic kernel
    g3, p3 = counts()
    assert (g3 - g2, p3 - p2) == (1, 0)
    assert A.peak_error(y_per, y_prec) <= 3e-7                     # the two kernels, same stream


@pytest.mark.parametrize("ch,preset,src,dst,advance", [
    (2, 3, 44100, 48000, True),         # art's default: extrapolation on, half-filter delay cancelled
    (1, 1, 48000, 37000, True),
    (3, 2, 32000, 40000, False),        # no advance: the first output comes after one input frame, nothing to extrapolate from
    (2, 4, 96000, 44100, True),
])
def test_endpoint_extrapolation(ch, preset, src, dst, advance):
    """EXTRAPOLATE_ENDPOINTS (resampler.c:516-522, :663-698; extrapolator.c): LPC prediction backwards in front of the
    first sample and forwards behind the last one, instead of silence.  Oracle restatement pinned bit-exactly against the
    compiled reference (tests/test_oracle_golden.py)."""
    filters, taps = A.PRESETS[preset]
    flags = BH_INTERP | A.EXTRAPOLATE_ENDPOINTS
    ratio = dst / src
    t = np.arange(40000)[:, None]
    rng = np.random.default_rng(22)
    for planar in (False, True):
        g, o = _pair(ch, taps, filters, lowpass_ratio=0.0, flags=flags)
        if advance:
            g.advance(taps / 2); o.advance(taps / 2)
        at = 0
        for b, n in enumerate([3000, 4096, 5, 7000]):
            x = (0.3 * np.sin(2 * np.pi * 0.013 * (t[at:at + n] + 17.0) + np.arange(ch)) + rng.normal(0, 0.02, (n, ch))).astype(np.float32)
            at += n
            yg, ug, gg = g.process(x, 20000, ratio, flush_after=(b == 3), planar=planar)
            yo, uo, go = o.process(x, 20000, ratio, flush_after=(b == 3))
            assert (ug, gg) == (uo, go)
            assert A.peak_error(yg, yo) <= TOL
        # the extrapolated ends are not silence: the first and last outputs continue the sine instead of fading in/out
        g.reset(); o.reset()
        x = rng.uniform(-0.5, 0.5, (2000, ch)).astype(np.float32)
        yg, ug, gg = g.process(x, 20000, ratio, flush_after=True, planar=planar)
        yo, uo, go = o.process(x, 20000, ratio, flush_after=True)
        assert (ug, gg) == (uo, go) and A.peak_error(yg, yo) <= TOL


@pytest.mark.parametrize("ch,preset,ratio", [
    (1, 2, 1.0), (2, 2, 1.00002), (3, 1, 0.9999), (8, 2, 1.0001), (5, 3, 0.99997), (2, 4, 1.00005), (4, 2, 1.0 - 3.8e-4), (2, 2, 1.0 + 4.5e-4),
])
def test_near_unity_ratios(ch, preset, ratio):
    """asynchronous sample-rate conversion: ratios within a few hundred ppm of 1 take the register-blocked form of the any-ratio
    kernel (runs of consecutive outputs that share a filter-row pair); the last two cases sit either side of its applicability
    limit |1/r - 1| * filters <= 1/8.  Several calls, one of them flushing."""
    filters, taps = A.PRESETS[preset]
    g, o = _pair(ch, taps, filters, lowpass_ratio=0.0)
    g.advance(taps / 2); o.advance(taps / 2)
    rng = np.random.default_rng(int(ratio * 1e6) % 1000 + ch)
    for b, n in enumerate([6000, 333, 9000]):
        x = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
        _check_call(g, o, x, n + taps + 64, ratio, flush_after=(b == 2))
