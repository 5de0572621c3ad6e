"""GPU parity tests of the PATH_WIDTH=64 library (libresampler_b200_64.so: every sample, tap and piece of filter state a double;
reference resampler.h:22-26, Makefile:12-19) through its C ABI, against oracle/liboracle64.so -- which tests/test_wide_cpu.py
pins to the reference's own PATH_WIDTH=64 build.

Bar: input_used / output_generated / resampleGetPosition bit-identical; samples within 1e-12 of the peak (double products
summed in another order than the reference's: observed ~1e-15); integer stages (decimator) bit-identical."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import artlibs as A
import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
TOL64 = 1e-12
BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS
ROOT = Path(__file__).resolve().parents[1]


def _pair(ch, taps, filters, **kw):
    return A.product_stream64(ch, taps, filters, **kw), A.oracle_stream64(ch, taps, filters, **kw)


def _check(g, o, x, cap, ratio, **kw):
    yg, ug, gg = g.process(x, cap, ratio, **kw)
    yo, uo, go = o.process(x, cap, ratio, **kw)
    assert (ug, gg) == (uo, go) and g.position() == o.position()
    assert yg.dtype == np.float64
    err = A.peak_error(yg, yo)
    assert err <= TOL64, f"max|d|/peak = {err:.3g}"
    return err


def test_bank_is_bit_identical_to_the_oracle():
    for preset in (1, 2, 3):
        filters, taps = A.PRESETS[preset]
        g, o = _pair(2, taps, filters, lowpass_ratio=0.0)
        assert g.bank().dtype == np.float64 and np.array_equal(g.bank(), o.bank())


@pytest.mark.parametrize("ch,preset,src,dst,lowpass_hz", [
    (1, 1, 44100, 48000, 0), (2, 3, 44100, 48000, 0), (5, 2, 96000, 44100, 20000), (2, 3, 48000, 44100, 20000), (8, 2, 48000, 48004.8, 0),
    (3, 4, 44100, 96000, 0), (2, 2, 44100, 44100 * 1.08843537, 0),
])
def test_baseline_shapes(ch, preset, src, dst, lowpass_hz):
    """the BASELINE configs' shapes artest-style (4096-frame calls, flush at the end), interleaved and planar"""
    filters, taps = A.PRESETS[preset]
    ratio = dst / src
    for planar in (False, True):
        g, o = _pair(ch, taps, filters, lowpass_ratio=lowpass_hz * 2.0 / src, flags=BH_INTERP)
        g.advance(taps / 2); o.advance(taps / 2)
        rng = np.random.default_rng(ch * 100 + preset)
        for b in range(4):
            x = rng.uniform(-0.5, 0.5, (4096, ch))
            _check(g, o, x, int(4096 * ratio) + 64, ratio, flush_after=(b == 3), planar=planar)


def test_fixed_ratio_is_chunking_invariant_and_passthrough_exact():
    """resampleFixedRatioInit contexts: bit-identical output for any call chunking (SURVEY 8b semantic 5), exact pass-through of
    the stored samples at integer positions (resampler.c:1141-1142)"""
    filters, taps = A.PRESETS[1]
    rng = np.random.default_rng(4)
    x = rng.uniform(-0.5, 0.5, (6000, 2))
    outs = []
    for chunks in ([6000], [1000, 2500, 17, 2483], [1] * 40 + [5960]):
        g = A.product_stream64(2, taps, filters, flags=BH_INTERP, fixed=(44100, 88200, 0))
        g.advance(taps / 2)
        y, at = [], 0
        for n in chunks:
            yy, u, m = g.process(x[at:at + n], 2 * n + 200, 1.0)
            assert u == n
            y.append(yy); at += n
        outs.append(np.concatenate(y))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[0][0::2][:5000], x[:5000])           # every other output IS an input sample
    o = A.oracle_stream64(2, taps, filters, flags=BH_INTERP, fixed=(44100, 88200, 0)); o.advance(taps / 2)
    yo, _, _ = o.process(x, 12200, 1.0)
    assert A.peak_error(outs[0], yo) <= TOL64


def test_ragged_limited_flush_reset_and_asrc():
    filters, taps = A.PRESETS[2]
    g, o = _pair(3, taps, filters, lowpass_ratio=0.0)
    g.advance(taps / 2); o.advance(taps / 2)
    rng = np.random.default_rng(11)
    for n, cap, ratio in [(900, 2000, 1.0001), (1, 10, 0.9999), (0, 10, 1.0), (5000, 700, 1.00003), (4300, 9000, 0.99995), (100, 300, 1.0)]:
        _check(g, o, rng.uniform(-0.5, 0.5, (n, 3)), cap, ratio)
    _check(g, o, None, 500, 1.0)                 # flush
    _check(g, o, rng.uniform(-0.5, 0.5, (50, 3)), 100, 1.0)      # input after a flush is ignored
    g.reset(); o.reset()
    g.advance(taps / 2); o.advance(taps / 2)
    _check(g, o, rng.uniform(-0.5, 0.5, (3000, 3)), 4000, 1.37)


def test_endpoint_extrapolation():
    t = np.arange(5000)
    x = (0.4 * np.sin(2 * np.pi * 0.013 * t) + 0.2 * np.sin(2 * np.pi * 0.071 * t + 1.0))[:, None] * np.array([[1.0, -0.5]])
    filters, taps = A.PRESETS[3]
    g, o = _pair(2, taps, filters, lowpass_ratio=0.0, flags=BH_INTERP | A.EXTRAPOLATE_ENDPOINTS)
    g.advance(taps / 2); o.advance(taps / 2)
    _check(g, o, x[:3000], 6000, 48000 / 44100)
    _check(g, o, x[3000:], 6000, 48000 / 44100, flush_after=True)


def test_device_pointer_batch_and_blocks():
    """the extension entry points with double device buffers: a batch of desynchronised contexts in one launch, and an ASRC block
    sequence with per-block positions"""
    import torch
    lib = A.product64()
    filters, taps = A.PRESETS[2]
    ch, n = 2, 12
    gs = [A.product_stream64(ch, taps, filters, 0.0) for _ in range(n)]
    os_ = [A.oracle_stream64(ch, taps, filters, 0.0) for _ in range(n)]
    rng = np.random.default_rng(8)
    for i, (g, o) in enumerate(zip(gs, os_)):
        g.advance(taps / 2); o.advance(taps / 2)
        x = rng.uniform(-0.5, 0.5, (100 + 37 * i, ch))
        _check(g, o, x, 4000, 1.1)                                    # desynchronise the contexts
    xs = [rng.uniform(-0.5, 0.5, (3000 + 11 * i, ch)) for i in range(n)]
    dx = [torch.from_numpy(x).cuda() for x in xs]
    dy = [torch.full((5000, ch), float("nan"), device="cuda", dtype=torch.float64) for _ in range(n)]
    st = torch.cuda.Stream(); torch.cuda.synchronize()
    ctxs = (type(gs[0].ctx) * n)(*[g.ctx for g in gs])
    ins = (C.c_void_p * n)(*[t.data_ptr() for t in dx]); outs = (C.c_void_p * n)(*[t.data_ptr() for t in dy])
    nin = (C.c_int * n)(*[x.shape[0] for x in xs]); nout = (C.c_int * n)(*([5000] * n))
    rat = (C.c_double * n)(*([1.1] * n)); res = (A.Result * n)()
    lib.resampleBatchProcessInterleavedDevice(ctxs, n, ins, nin, outs, nout, rat, res, C.c_void_p(st.cuda_stream))
    st.synchronize()
    for i in range(n):
        yo, uo, go = os_[i].process(xs[i], 5000, 1.1)
        assert (res[i].input_used, res[i].output_generated) == (uo, go) and gs[i].position() == os_[i].position()
        assert A.peak_error(dy[i][:go].cpu().numpy(), yo) <= TOL64
    # ASRC blocks
    g, o = _pair(ch, taps, filters, lowpass_ratio=0.0)
    g.advance(taps / 2); o.advance(taps / 2)
    nb, bf = 16, 480
    x = rng.uniform(-0.5, 0.5, (nb * bf, ch))
    ratios = [1.0 + 1e-4 * np.sin(2 * np.pi * k / nb) for k in range(nb)]
    dxt = torch.from_numpy(x).cuda(); dyt = torch.full((nb * bf + 400, ch), float("nan"), device="cuda", dtype=torch.float64)
    bfa = (C.c_int * nb)(*([bf] * nb)); rt = (C.c_double * nb)(*ratios); res = (A.Result * nb)(); pos = (C.c_double * nb)()
    done = lib.resampleProcessBlocksInterleavedDevice(g.ctx, dxt.data_ptr(), bfa, rt, nb, dyt.data_ptr(), nb * bf + 400, res, pos, C.c_void_p(st.cuda_stream))
    st.synchronize()
    assert done == nb
    at = 0
    for k in range(nb):
        yo, uo, go = o.process(x[k * bf:(k + 1) * bf], 1000, ratios[k])
        assert (res[k].input_used, res[k].output_generated) == (uo, go) and pos[k] == o.position()
        assert A.peak_error(dyt[at:at + go].cpu().numpy(), yo) <= TOL64
        at += go


class OCo64(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class OBq64(C.Structure):
    _fields_ = [("a", C.c_double * 5), ("b", C.c_double * 5), ("xh", C.c_double * 4), ("yh", C.c_double * 4),
                ("order", C.c_int), ("cursor", C.c_int)]


def test_biquad_cascade_and_struct_state():
    """biquad_apply_buffer / the one-pass cascade in double: samples within 1e-12 of the oracle's recurrence (the chunked scan
    propagates state in double and reruns every chunk with the reference's operation order), rings and index as the reference
    leaves them"""
    pkg = entry.load_package(); lib = pkg.load64(); ol = A.oracle64()
    ol.oracle_biquad_init.argtypes = [C.POINTER(OBq64), C.POINTER(OCo64), C.c_double]
    ol.oracle_biquad_run.argtypes = [C.POINTER(OBq64), A.f64p, C.c_int, C.c_int]
    co = pkg.BiquadCoefficients64()
    lib.biquad_lowpass(C.byref(co), 0.45 * 44100 / 96000)
    oco = OCo64(*[getattr(co, n) for n, _ in OCo64._fields_])
    ch, n = 5, 40000
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.5, 0.5, (n, ch))
    got, ref = x.copy(), x.copy()
    stages = [(pkg.Biquad64 * ch)() for _ in range(2)]
    for st in stages:
        for q in st:
            lib.biquad_init(C.byref(q), C.byref(co), 1.0)
    arr = (C.POINTER(pkg.Biquad64) * 2)(*[C.cast(st, C.POINTER(pkg.Biquad64)) for st in stages])
    for lo, hi in [(0, 12345), (12345, n)]:
        lib.biquad_apply_cascade_interleaved(arr, 2, ch, got[lo:].ctypes.data_as(A.f64p), hi - lo)
    for c in range(ch):
        for _ in range(2):
            oq = OBq64(); ol.oracle_biquad_init(C.byref(oq), C.byref(oco), 1.0)
            ol.oracle_biquad_run(C.byref(oq), ref[:, c:].ctypes.data_as(A.f64p), n, ch)
    assert A.peak_error(got, ref) <= 1e-11
    assert all(q.index == n for st in stages for q in st)


@pytest.mark.parametrize("flags", [0, 0x2, 0x1 | 0x200, 0x800, 0x4 | 0x100])
def test_decimator_is_bit_identical(flags):
    """double in, integer bytes out: every byte, the clipped-sample count and a second call (state carried) equal the oracle's"""
    pkg = entry.load_package(); lib = pkg.load64(); ol = A.oracle64()
    ol.oracle_decimate_init.restype = C.c_void_p
    ol.oracle_decimate_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    ol.oracle_decimate_free.argtypes = [C.c_void_p]
    ol.oracle_decimate_interleaved.argtypes = [C.c_void_p, A.f64p, C.c_int, C.c_char_p]
    ol.oracle_decimate_interleaved.restype = C.c_int
    ol.oracle_float_integers.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, A.f64p, C.c_int]
    rng = np.random.default_rng(flags + 7)
    for ch, bits, bytes_, gain, rate in [(2, 16, 2, 1.0, 44100), (3, 24, 3, 1.3, 96000), (1, 8, 1, 0.9, 48000), (2, 24, 4, 1.0, 32000)]:
        g = lib.decimateInit(ch, bits, bytes_, gain, rate, flags)
        o = ol.oracle_decimate_init(ch, bits, bytes_, gain, rate, flags)
        for n in (777, 1, 2500):
            x = rng.uniform(-1.0, 1.0, (n, ch))
            bg, bo = C.create_string_buffer(n * ch * bytes_ + 8), C.create_string_buffer(n * ch * bytes_ + 8)
            cg = lib.decimateProcessInterleavedLE(g, x.ctypes.data_as(A.f64p), n, C.cast(bg, C.c_void_p))
            co = ol.oracle_decimate_interleaved(o, x.ctypes.data_as(A.f64p), n, bo)
            assert cg == co and bg.raw == bo.raw, (ch, bits, flags, n)
            # and back: floatIntegersLE
            back_g, back_o = np.zeros(n * ch), np.zeros(n * ch)
            lib.floatIntegersLE(C.cast(bg, C.c_void_p), 1.0 / gain, bits, bytes_, 1, back_g.ctypes.data_as(A.f64p), n * ch)
            ol.oracle_float_integers(bo, 1.0 / gain, bits, bytes_, 1, back_o.ctypes.data_as(A.f64p), n * ch)
            assert np.array_equal(back_g, back_o)
        lib.decimateFree(g); ol.oracle_decimate_free(o)


def test_artest64_relinked_against_the_library():
    """the reference's own test program compiled with -DPATH_WIDTH=64 (artest64 of its Makefile) and linked against
    libresampler_b200_64.so instead of the reference sources (oracle/Makefile): same frame counts as the reference binary, and its
    round-trip residual check passes"""
    ref, mine = ROOT / "oracle" / "_ref" / "artest64_ref", ROOT / "oracle" / "_ref" / "artest64_b200"
    if not (ref.exists() and mine.exists()):
        pytest.skip("oracle/_ref/artest64_* not built (needs /root/reference at build time)")
    args = ["-2", "-c2", "-n4", "-s44100", "-d48000", "-i"]
    a = subprocess.run([str(ref)] + args, capture_output=True, text=True, timeout=300)
    b = subprocess.run([str(mine)] + args, capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-500:], b.stderr[-500:])
    import re
    def rows(s):
        out = []
        for ln in (s.stdout + s.stderr).splitlines():
            m = re.search(r"\((-w\d)\): count =\s*(\d+), checksum = \w+, range = (\S+) to (\S+), RMS = (\S+) dB", ln)
            if m:
                out.append((m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))))
        return out
    ra, rb = rows(a), rows(b)
    assert len(ra) == 4 and [r[:2] for r in ra] == [r[:2] for r in rb]           # same streams, same frame counts
    for x, y in zip(ra, rb):
        assert abs(x[2] - y[2]) <= 2e-7 and abs(x[3] - y[3]) <= 2e-7 and abs(x[4] - y[4]) <= 0.05      # printed to 7 digits / 0.01 dB
