"""GPU parity tests of the tensor-core (tcgen05) form of the rational-ratio path (art_sinc_umma.cu).

By default the library only takes that kernel when a launch holds enough work to fill the GPU; these tests
force it (resampleB200SetTensorPath(2)) so that small, oracle-checkable calls run on it too, and check that it
really ran (resampleB200TensorLaunches).  Same bar as test_gpu_parity.py: counts and position bit-identical,
samples within 1e-6 * peak(reference).
"""
import numpy as np
import pytest

import artlibs as A

pytestmark = pytest.mark.gpu

TOL = 1e-6
BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS


@pytest.fixture(autouse=True)
def forced_tensor_path():
    lib = A.product()
    lib.resampleB200SetTensorPath(2)
    yield lib
    lib.resampleB200SetTensorPath(1)


def _pair(ch, taps, filters, **kw):
    return A.product_stream(ch, taps, filters, **kw), A.oracle_stream(ch, taps, filters, **kw)


def _check_call(g, o, x, cap, ratio, tol=TOL, **kw):
    yg, ug, gg = g.process(x, cap, ratio, **kw)
    yo, uo, go = o.process(x, cap, ratio, **kw)
    assert (ug, gg) == (uo, go), f"counts differ: gpu {(ug, gg)} oracle {(uo, go)}"
    assert g.position() == o.position()
    err = A.peak_error(yg, yo)
    assert err <= tol, f"max|d|/peak = {err:.3g}"
    return yg, yo


@pytest.mark.parametrize("ch,preset,src,dst,lowpass_hz", [
    (2, 3, 44100, 48000, 0),            # BASELINE config 2 (metric config): 160/147
    (1, 1, 44100, 48000, 0),            # BASELINE config 1: short filter, few k-steps
    (2, 3, 48000, 44100, 20000),        # one stream of BASELINE config 4: 147/160, phases padded to 160
    (3, 2, 32000, 40000, 0),            # 5/4: 32 periods grouped into one row of 160 phases
    (5, 1, 8000, 48000, 0),             # 6/1: 26 periods per row, odd channel count
    (2, 4, 44100, 48000, 0),            # preset -4: 988 taps, 8 row shifts
    (8, 4, 96000, 44100, 20000),        # BASELINE config 3's shape (8 of its 64 channels): 147/320, operand A as a ring of plane pairs
    (4, 3, 96000, 48000, 0),            # 1/2: 160 periods per row, M = 320
    (2, 3, 44100, 96000, 0),            # 320/147: two phase groups of 160
    (1, 2, 8000, 44100, 0),             # 441/80: three phase groups of 147
    (12, 2, 44100, 48000, 0),           # many channels, a multiple of 4: tiles of four channels read 16-byte slices of the interleaved frames
    (10, 2, 96000, 44100, 0),           # many channels, not a multiple of 4: through planar scratch (art_device.cu)
])
def test_configs_on_the_tensor_path(forced_tensor_path, ch, preset, src, dst, lowpass_hz):
    lib = forced_tensor_path
    filters, taps = A.PRESETS[preset]
    g, o = _pair(ch, taps, filters, lowpass_ratio=lowpass_hz * 2.0 / src, flags=BH_INTERP)
    g.advance(taps / 2); o.advance(taps / 2)
    ratio = dst / src
    rng = np.random.default_rng(200 + ch + preset)
    before = lib.resampleB200TensorLaunches()
    blocks = 3
    for b in range(blocks):
        x = rng.uniform(-0.5, 0.5, (8192, ch)).astype(np.float32)
        _check_call(g, o, x, int(8192 * ratio) + taps + 10, ratio, flush_after=(b == blocks - 1))
    assert lib.resampleB200TensorLaunches() - before >= blocks - 1, "the tensor-core kernel did not run"


def test_ragged_calls_planar_and_limits(forced_tensor_path):
    """odd call sizes, output-limited calls, the planar API, and calls too small for the kernel in between"""
    lib = forced_tensor_path
    g, o = _pair(2, 380, 380, lowpass_ratio=0.0)
    g2 = A.product_stream(2, 380, 380, 0.0)
    rng = np.random.default_rng(31)
    ratio = 48000 / 44100
    before = lib.resampleB200TensorLaunches()
    for n, cap in [(5000, 9000), (1, 10), (12345, 20000), (0, 10), (7000, 3000), (4097, 9000), (30011, 40000)]:
        x = rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
        yg, _ = _check_call(g, o, x, cap, ratio)
        yp, up, mp = g2.process(x, cap, ratio, planar=True)
        assert np.array_equal(yg, yp), "planar and interleaved calls differ"
    assert lib.resampleB200TensorLaunches() - before >= 8
    _check_call(g, o, None, 500, ratio)                                 # flush


def test_bit_reproducible(forced_tensor_path):
    """two issuer warps feed the tensor pipe; the order in which an accumulator sees its MMAs must be fixed"""
    rng = np.random.default_rng(33)
    x = rng.uniform(-0.5, 0.5, (50000, 2)).astype(np.float32)
    outs = []
    for _ in range(4):
        g = A.product_stream(2, 380, 380, 0.0)
        g.advance(190)
        y, u, m = g.process(x, 60000, 48000 / 44100)
        outs.append(y)
    assert all(np.array_equal(outs[0], y) for y in outs[1:])


@pytest.mark.parametrize("scale", [1e-30, 3e-5, 1.0, 32768.0, 1e20])
def test_block_scaling_follows_the_signal(forced_tensor_path, scale):
    """the fixed-point split is relative to each tile's maximum: any overall level must give the same relative error"""
    g, o = _pair(2, 380, 380, lowpass_ratio=0.0)
    g.advance(190); o.advance(190)
    rng = np.random.default_rng(35)
    x = (rng.uniform(-0.5, 0.5, (20000, 2)) * scale).astype(np.float32)
    _check_call(g, o, x, 30000, 48000 / 44100)


def test_quiet_passage_after_a_loud_one(forced_tensor_path):
    """a tile holding full-scale and -120 dB material: the error stays below 1e-6 of the call's peak, and a tile
    that is quiet throughout keeps its own relative accuracy"""
    g, o = _pair(1, 380, 380, lowpass_ratio=0.0)
    g.advance(190); o.advance(190)
    rng = np.random.default_rng(37)
    x = rng.uniform(-0.5, 0.5, (60000, 1)).astype(np.float32)
    x[30000:] *= 1e-6
    yg, yo = _check_call(g, o, x, 80000, 48000 / 44100)
    tail_g, tail_o = yg[45000:60000], yo[45000:60000]               # tiles that only hold the quiet part
    assert A.peak_error(tail_g, tail_o) <= TOL


def test_silence_and_impulse(forced_tensor_path):
    g, o = _pair(2, 380, 380, lowpass_ratio=0.0)
    g.advance(190); o.advance(190)
    x = np.zeros((20000, 2), np.float32)
    yg, yo = _check_call(g, o, x, 30000, 48000 / 44100)
    assert not yg.any()
    x[10000, 0] = 1.0
    x[10001, 1] = -0.75
    _check_call(g, o, x, 30000, 48000 / 44100)


def test_many_streams_batched(forced_tensor_path):
    """resampleBatchProcessInterleaved: independent streams with DIFFERENT states (tables are per state) and lengths;
    the host-pointer batch pipelines stream by stream (upload, launch, download), hence one launch per stream"""
    lib = forced_tensor_path
    rng = np.random.default_rng(39)
    ratio = 48000 / 44100
    n_streams = 5
    gs = [A.product_stream(2, 380, 380, 0.0) for _ in range(n_streams)]
    os_ = [A.oracle_stream(2, 380, 380, 0.0) for _ in range(n_streams)]
    for i, (g, o) in enumerate(zip(gs, os_)):
        g.advance(190); o.advance(190)
        if i % 2:                                                   # desynchronise some of the streams
            x = rng.uniform(-0.5, 0.5, (1000 + 37 * i, 2)).astype(np.float32)
            g.process(x, 5000, ratio); o.process(x, 5000, ratio)
    xs = [rng.uniform(-0.5, 0.5, (20000 + 1111 * i, 2)).astype(np.float32) for i in range(n_streams)]
    before = lib.resampleB200TensorLaunches()
    ys = A.batch_process(gs, xs, 40000, ratio)
    assert lib.resampleB200TensorLaunches() - before >= 1
    for g, o, x, (y, used, made) in zip(gs, os_, xs, ys):
        yo, uo, mo = o.process(x, 40000, ratio)
        assert (used, made) == (uo, mo) and g.position() == o.position()
        assert A.peak_error(y, yo) <= TOL


def test_fixed_ratio_contexts_opt_in(forced_tensor_path):
    """resampleFixedRatioInit contexts (no interpolation) stay on the FFMA form by default -- it keeps their output
    bit-identical across call chunkings -- and take the tensor-core kernel only in mode 3"""
    lib = forced_tensor_path
    rng = np.random.default_rng(41)
    x = rng.uniform(-0.5, 0.5, (200000, 2)).astype(np.float32)

    def run():
        g = A.product_stream(2, 380, 380, flags=7, fixed=(44100, 48000, 0))
        o = A.oracle_stream(2, 380, 380, flags=7, fixed=(44100, 48000, 0))
        g.advance(190); o.advance(190)
        before = lib.resampleB200TensorLaunches()
        _check_call(g, o, x, 250000, 0.0)
        return lib.resampleB200TensorLaunches() - before

    lib.resampleB200SetTensorPath(1)
    assert run() == 0
    lib.resampleB200SetTensorPath(3)
    assert run() == 1


def test_random_streams_tensor_equals_ffma(forced_tensor_path):
    """Seeded random sweep at sizes the oracle would take minutes for: the tensor-core form against the FFMA form of the
    same library (mode 0) -- counts and position identical, samples within 3e-7 of peak (both sit within 2e-7 of the
    oracle on the oracle-sized cases).  Ratios, presets, channel counts, chunkings, flush and extrapolation vary."""
    lib = forced_tensor_path
    rng = np.random.default_rng(77)
    rates = [(44100, 48000), (48000, 44100), (96000, 44100), (32000, 48000), (48000, 32000), (44100, 88200), (22050, 48000),
             (48000, 96000), (88200, 48000)]
    ran = 0
    for trial in range(14):
        src, dst = rates[int(rng.integers(len(rates)))]
        preset = int(rng.integers(1, 5))
        ch = int(rng.choice([1, 2, 2, 3, 4, 8]))
        filters, taps = A.PRESETS[preset]
        flags = BH_INTERP | (A.EXTRAPOLATE_ENDPOINTS if trial % 3 == 0 else 0)
        lowpass = 0.0 if dst >= src else 0.9 * dst / src
        ratio = dst / src
        total = int(rng.integers(40000, 120000))
        cuts = np.sort(rng.integers(1, total, size=int(rng.integers(0, 4))))
        chunks = np.diff(np.concatenate(([0], cuts, [total]))).tolist()
        x = (rng.uniform(-0.5, 0.5, (total, ch)) * rng.choice([1.0, 1e-3, 300.0])).astype(np.float32)
        outs = []
        for mode in (0, 2):
            lib.resampleB200SetTensorPath(mode)
            before = lib.resampleB200TensorLaunches()
            s = A.product_stream(ch, taps, filters, lowpass_ratio=lowpass, flags=flags)
            s.advance(taps / 2)
            ys, at, log = [], 0, []
            for i, n in enumerate(chunks):
                cap = int(n * ratio) + taps + 64
                y, u, m = s.process(x[at:at + n], cap, ratio, flush_after=(i == len(chunks) - 1), planar=bool(trial & 1))
                ys.append(y); log.append((u, m, s.position())); at += u
            outs.append((np.concatenate(ys), log, lib.resampleB200TensorLaunches() - before))
        (y0, log0, n0), (y2, log2, n2) = outs
        assert log0 == log2, "counts / positions differ between the two forms"
        assert n0 == 0
        ran += n2
        assert A.peak_error(y2, y0) <= 3e-7, (trial, src, dst, preset, ch, A.peak_error(y2, y0))
    assert ran >= 10, "the sweep hardly reached the tensor-core kernel"


@pytest.mark.parametrize("digits", [3, 2])
@pytest.mark.parametrize("step_db", [0, 30, 60, 90, 120])
def test_error_relative_to_the_local_peak(forced_tensor_path, step_db, digits):
    """The tensor-core form's error measured against the LOCAL signal level instead of the call's peak: a stretch `step_db`
    below full scale that shares a tile (0.2 s of a mono stream) with full-scale material.
      3 signal digits (default): every sample keeps >= 22 bits of its own magnitude -> float accuracy relative to every
                                 window, like the reference's float path and this library's FFMA kernels (checked here too);
      2 digits (opt-in):         exact to 2^-24 of the TILE's peak -> relative to the quiet stretch the error grows with the
                                 level difference (the curve goes into gpurun_out/tensor_local_error.json for DESIGN.md)."""
    import json, os
    lib = forced_tensor_path
    ratio, taps = 48000 / 44100, 380
    rng = np.random.default_rng(43)
    n = 56448                                               # three tiles of 18816 frames
    x = rng.uniform(-0.5, 0.5, (n, 1)).astype(np.float32)
    lo, hi = 18816 + 6000, 18816 + 14000                     # a quiet stretch strictly inside the second tile
    x[lo:hi] *= np.float32(10.0 ** (-step_db / 20.0))
    o = A.oracle_stream(1, taps, 380, 0.0); o.advance(taps / 2)
    yo, uo, mo = o.process(x, 70000, ratio)
    qlo, qhi = int((lo + taps) * ratio), int((hi - taps) * ratio)      # outputs whose windows lie inside the quiet stretch
    local_peak = float(np.max(np.abs(yo[qlo:qhi])))
    row = {"step_db": step_db, "digits": digits, "local_peak": local_peak}
    lib.resampleB200SetTensorDigits(digits)
    try:
        for name, mode in (("tensor", 2), ("ffma", 0)):
            lib.resampleB200SetTensorPath(mode)
            g = A.product_stream(1, taps, 380, 0.0); g.advance(taps / 2)
            yg, ug, mg = g.process(x, 70000, ratio)
            assert (ug, mg) == (uo, mo)
            d = np.abs(yg.astype(np.float64) - yo)
            row[name] = {"call_relative": float(d.max() / np.max(np.abs(yo))), "local_relative": float(d[qlo:qhi].max() / local_peak),
                         "local_abs_over_tile_peak": float(d[qlo:qhi].max() / 0.5)}
    finally:
        lib.resampleB200SetTensorDigits(3)
        lib.resampleB200SetTensorPath(2)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/tensor_local_error.json", "a") as f:
        f.write(json.dumps(row) + "\n")
    assert row["ffma"]["local_relative"] <= TOL                          # float accuracy relative to every window
    assert row["tensor"]["call_relative"] <= TOL
    if digits == 3:
        assert row["tensor"]["local_relative"] <= TOL                    # ... and so has the tensor-core form by default
    else:
        assert row["tensor"]["local_abs_over_tile_peak"] <= 2.0 ** -22   # the documented guarantee of the 2-digit mode


def test_non_finite_samples_poison_their_neighbourhood_only():
    """A NaN or Inf input sample makes the reference's outputs non-finite exactly where a window holds it (every tap multiplies it,
    resampler.c:1033-1044).  The kernels pad: the FFMA form's windows to a multiple of 32 taps, the tensor-core form's to the band
    of a tile row -- zero taps, but 0 x NaN is NaN -- so they poison a superset: at most one tile row (L outputs) more on either
    side.  Everything else must be finite and within tolerance: a non-finite sample must not upset a tile's block scaling."""
    lib = A.product()
    filters, taps = A.PRESETS[3]
    ratio = 48000 / 44100
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, (40000, 2)).astype(np.float32)
    x[12345, 0] = np.nan
    x[30000, 1] = np.inf
    o = A.oracle_stream(2, taps, filters, 0.0); o.advance(taps / 2)
    yo, uo, mo = o.process(x, 50000, ratio)
    bad_o = ~np.isfinite(yo)
    assert bad_o.any()
    try:
        for mode in (0, 2):
            lib.resampleB200SetTensorPath(mode)
            g = A.product_stream(2, taps, filters, 0.0); g.advance(taps / 2)
            y, u, m = g.process(x, 50000, ratio)
            assert (u, m) == (uo, mo)
            bad = ~np.isfinite(y)
            assert not (bad_o & ~bad).any(), "an output the reference poisons came out finite"
            extra = np.argwhere(bad & ~bad_o)[:, 0]
            lo_, hi_ = np.argwhere(bad_o)[:, 0].min(), np.argwhere(bad_o)[:, 0].max()
            assert extra.size == 0 or (extra.min() >= lo_ - 160 and extra.max() <= hi_ + 160)
            ok = ~bad
            assert np.max(np.abs(y[ok] - yo[ok])) <= TOL * np.max(np.abs(yo[ok]))
    finally:
        lib.resampleB200SetTensorPath(1)
