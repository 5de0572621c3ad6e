"""ctypes views of the three libraries the tests juggle.

* ``Oracle``   -- oracle/liboracle.so, this project's CPU restatement (the checker).
* ``Reference``-- oracle/_ref/libartref.so, the unmodified reference compiled by
                  oracle/Makefile (present here and, prebuilt, on the GPU box).
* ``Product``  -- audio-resampler_b200/lib/libresampler_b200.so through its C ABI
                  (include/resampler.h, include/biquad.h, include/resampler_b200.h).

All three expose the same small Python surface (``Stream`` objects with
``process`` / ``flush`` / ``position`` ...) so parity tests read like the reference's
own artest loop (artest.c:446-491).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
PKG_DIR = ROOT / "audio-resampler_b200"

# flag values, resampler.h:28-38
SUBSAMPLE_INTERPOLATE = 0x1
BLACKMAN_HARRIS = 0x2
INCLUDE_LOWPASS = 0x4
RESAMPLE_MULTITHREADED = 0x8
NO_FILTER_REDUCTION = 0x10
EXTRAPOLATE_ENDPOINTS = 0x40
EXTEND_CONVOLUTION_MATH = 0x100

PRESETS = {1: (48, 48), 2: (320, 156), 3: (380, 380), 4: (988, 988)}  # (filters, taps), artest.c:154-169


class Result(C.Structure):
    _fields_ = [("input_used", C.c_uint), ("output_generated", C.c_uint)]


f32p = C.POINTER(C.c_float)
f32pp = C.POINTER(f32p)
f64p = C.POINTER(C.c_double)
f64pp = C.POINTER(f64p)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(f64p if a.dtype == np.float64 else f32p)


def _sample(lib):
    """numpy dtype and ctypes pointer type of a library's samples: float32, or float64 for the PATH_WIDTH=64 builds"""
    return (np.float64, f64p) if getattr(lib, "_wide", False) else (np.float32, f32p)


def _build_oracle():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True, capture_output=True)


def load_oracle(wide: bool = False) -> C.CDLL:
    so = ORACLE_DIR / ("liboracle64.so" if wide else "liboracle.so")
    if not so.exists() or so.stat().st_mtime < (ORACLE_DIR / "art_oracle.c").stat().st_mtime:
        _build_oracle()
    lib = C.CDLL(str(so), mode=os.RTLD_LOCAL)
    lib._wide = wide
    f32p, f32pp = (f64p, f64pp) if wide else (globals()["f32p"], globals()["f32pp"])
    vp = C.c_void_p
    lib.oracle_init.restype = vp
    lib.oracle_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    lib.oracle_fixed_ratio_init.restype = vp
    lib.oracle_fixed_ratio_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.oracle_free.argtypes = [vp]
    lib.oracle_reset.argtypes = [vp]
    lib.oracle_advance.argtypes = [vp, C.c_double]
    lib.oracle_position.restype = C.c_double
    lib.oracle_position.argtypes = [vp]
    lib.oracle_bank_row.restype = f32p
    lib.oracle_bank_row.argtypes = [vp, C.c_int]
    for name in ("oracle_process_interleaved", "oracle_process_flush_interleaved"):
        fn = getattr(lib, name)
        fn.restype = Result
        fn.argtypes = [vp, f32p, C.c_int, f32p, C.c_int, C.c_double]
    for name in ("oracle_process_planar", "oracle_process_flush_planar"):
        fn = getattr(lib, name)
        fn.restype = Result
        fn.argtypes = [vp, f32pp, C.c_int, f32pp, C.c_int, C.c_double]
    lib.oracle_required_input.restype = C.c_uint
    lib.oracle_required_input.argtypes = [vp, C.c_int, C.c_double]
    lib.oracle_expected_output.restype = C.c_uint
    lib.oracle_expected_output.argtypes = [vp, C.c_int, C.c_double]
    lib.oracle_noise.argtypes = [C.POINTER(C.c_ulonglong), f32p, C.c_int]
    return lib


class RefResample(C.Structure):
    """Leading public fields of the reference context (resampler.h:44-48)."""
    _fields_ = [("numChannels", C.c_int), ("numSamples", C.c_int), ("numFilters", C.c_int),
                ("numTaps", C.c_int), ("inputIndex", C.c_int), ("flags", C.c_int),
                ("tempFilter", C.c_void_p), ("outputOffset", C.c_double), ("fixedRatio", C.c_double),
                ("lowpassRatio", C.c_double), ("subsample", C.c_void_p),
                ("buffers", f32pp), ("filters", f32pp)]


class BiquadCoefficients(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class Biquad(C.Structure):
    _fields_ = [("a", C.c_float * 5), ("b", C.c_float * 5), ("x", C.c_float * 4), ("y", C.c_float * 4),
                ("order", C.c_int), ("index", C.c_int)]


class RefResample64(C.Structure):
    _fields_ = [(n, (f64pp if n in ("buffers", "filters") else t)) for n, t in RefResample._fields_]


class BiquadCoefficients64(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class Biquad64(C.Structure):
    _fields_ = [("a", C.c_double * 5), ("b", C.c_double * 5), ("x", C.c_double * 4), ("y", C.c_double * 4),
                ("order", C.c_int), ("index", C.c_int)]


def bind_reference_api(lib: C.CDLL, wide: bool = False) -> C.CDLL:
    """Attach the prototypes of resampler.h:64-78 and biquad.h:41-47 (shared by the
    reference build and by the product, which exports the same symbols).  wide: the PATH_WIDTH=64 builds."""
    lib._wide = wide
    f32p, f32pp = (f64p, f64pp) if wide else (globals()["f32p"], globals()["f32pp"])
    Biquad, BiquadCoefficients = (Biquad64, BiquadCoefficients64) if wide else (globals()["Biquad"], globals()["BiquadCoefficients"])
    sample_c = C.c_double if wide else C.c_float
    ctx = C.POINTER(RefResample64 if wide else RefResample)
    lib._ctx_type = ctx
    lib.resampleInit.restype = ctx
    lib.resampleInit.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    lib.resampleFixedRatioInit.restype = ctx
    lib.resampleFixedRatioInit.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    for name in ("resampleProcess", "resampleProcessAndFlush"):
        fn = getattr(lib, name)
        fn.restype = Result
        fn.argtypes = [ctx, f32pp, C.c_int, f32pp, C.c_int, C.c_double]
    for name in ("resampleProcessInterleaved", "resampleProcessAndFlushInterleaved"):
        fn = getattr(lib, name)
        fn.restype = Result
        fn.argtypes = [ctx, f32p, C.c_int, f32p, C.c_int, C.c_double]
    lib.resampleGetRequiredSamples.restype = C.c_uint
    lib.resampleGetRequiredSamples.argtypes = [ctx, C.c_int, C.c_double]
    lib.resampleGetExpectedOutput.restype = C.c_uint
    lib.resampleGetExpectedOutput.argtypes = [ctx, C.c_int, C.c_double]
    lib.resampleAdvancePosition.argtypes = [ctx, C.c_double]
    lib.resampleAdvancePosition.restype = None
    lib.resampleGetLowpassRatio.restype = C.c_double
    lib.resampleGetLowpassRatio.argtypes = [ctx]
    lib.resampleGetPosition.restype = C.c_double
    lib.resampleGetPosition.argtypes = [ctx]
    lib.resampleGetNumFilters.restype = C.c_int
    lib.resampleGetNumFilters.argtypes = [ctx]
    lib.resampleInterpolationUsed.restype = C.c_int
    lib.resampleInterpolationUsed.argtypes = [ctx]
    lib.resampleReset.argtypes = [ctx]
    lib.resampleReset.restype = None
    lib.resampleFree.argtypes = [ctx]
    lib.resampleFree.restype = None
    lib.biquad_init.argtypes = [C.POINTER(Biquad), C.POINTER(BiquadCoefficients), C.c_double]
    lib.biquad_init.restype = None
    lib.biquad_lowpass.argtypes = [C.POINTER(BiquadCoefficients), C.c_double]
    lib.biquad_lowpass.restype = None
    lib.biquad_highpass.argtypes = [C.POINTER(BiquadCoefficients), C.c_double]
    lib.biquad_highpass.restype = None
    lib.biquad_apply_buffer.argtypes = [C.POINTER(Biquad), f32p, C.c_int, C.c_int]
    lib.biquad_apply_buffer.restype = None
    lib.biquad_apply_sample.argtypes = [C.POINTER(Biquad), sample_c]
    lib.biquad_apply_sample.restype = sample_c
    return lib


def reference_path(wide: bool = False) -> Path:
    return ORACLE_DIR / "_ref" / ("libartref64.so" if wide else "libartref.so")


def load_reference(wide: bool = False) -> C.CDLL | None:
    so = reference_path(wide)
    if not so.exists():
        try:
            _build_oracle()
        except Exception:
            return None
    if not so.exists():
        return None
    # RTLD_LOCAL: the product exports the same symbol names
    return bind_reference_api(C.CDLL(str(so), mode=os.RTLD_LOCAL), wide)


def product_path(wide: bool = False) -> Path:
    return PKG_DIR / "lib" / ("libresampler_b200_64.so" if wide else "libresampler_b200.so")


def load_product(wide: bool = False) -> C.CDLL:
    so = product_path(wide)
    if not so.exists():
        raise RuntimeError(f"{so} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = bind_reference_api(C.CDLL(str(so), mode=os.RTLD_LOCAL), wide)
    # the extension entry points the tests use (include/resampler_b200.h)
    ctx = lib._ctx_type
    f32p = f64p if wide else globals()["f32p"]
    Biquad = Biquad64 if wide else globals()["Biquad"]
    lib.resampleB200SetTensorPath.restype = None
    lib.resampleB200SetTensorPath.argtypes = [C.c_int]
    lib.resampleB200AttachPrefilter.restype = C.c_int
    lib.resampleB200AttachPrefilter.argtypes = [ctx, C.POINTER(Biquad), C.c_int]
    lib.resampleB200SetTensorDigits.restype = None
    lib.resampleB200SetTensorDigits.argtypes = [C.c_int]
    lib.resampleB200TensorLaunches.restype = C.c_ulonglong
    lib.resampleB200TensorLaunches.argtypes = []
    lib.resampleBatchProcessInterleaved.restype = None
    lib.resampleBatchProcessInterleaved.argtypes = [C.POINTER(ctx), C.c_int, C.POINTER(f32p), C.POINTER(C.c_int),
                                                    C.POINTER(f32p), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(Result)]
    # device-pointer entry points: device addresses travel as integers
    vp = C.c_void_p
    lib.resampleProcessInterleavedDevice.restype = Result
    lib.resampleProcessInterleavedDevice.argtypes = [ctx, vp, C.c_int, vp, C.c_int, C.c_double, vp]
    lib.resampleProcessDevice.restype = Result
    lib.resampleProcessDevice.argtypes = [ctx, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.c_double, vp]
    lib.resampleBatchProcessInterleavedDevice.restype = None
    lib.resampleBatchProcessInterleavedDevice.argtypes = [C.POINTER(ctx), C.c_int, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(vp),
                                                          C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(Result), vp]
    lib.resampleProcessBlocksInterleavedDevice.restype = C.c_int
    lib.resampleProcessBlocksInterleavedDevice.argtypes = [ctx, vp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, vp, C.c_int,
                                                           C.POINTER(Result), C.POINTER(C.c_double), vp]
    lib.resampleB200PathCounts.restype = None
    lib.resampleB200PathCounts.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.resampleB200LastError.restype = C.c_char_p
    lib.resampleB200LastError.argtypes = [C.c_int]
    lib.resampleB200Synchronize.restype = None
    lib.resampleB200Synchronize.argtypes = [ctx]
    return lib


def path_counts(lib):
    """(generic, periodic FFMA, tensor) convolution launches so far"""
    g, p = C.c_ulonglong(), C.c_ulonglong()
    lib.resampleB200PathCounts(C.byref(g), C.byref(p))
    return g.value, p.value, lib.resampleB200TensorLaunches()


# ---------------------------------------------------------------------------
# device-pointer entry points (include/resampler_b200.h): torch only provides the device memory and the stream

def _cuda():
    import torch
    return torch


def device_batch_process(streams, xs, n_out, ratios, use_stream=True):
    """resampleBatchProcessInterleavedDevice: ONE launch over all contexts.  xs[i] is (frames, ch) float32 or None for
    a flush (numInputFrames = -1).  Returns [(y, input_used, output_generated)]."""
    torch = _cuda()
    lib, n, ch = streams[0].lib, len(streams), streams[0].channels
    st = torch.cuda.Stream() if use_stream else None
    dx, nin = [], []
    for x in xs:
        if x is None:
            dx.append(None); nin.append(-1)
        else:
            x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ch)
            dx.append(torch.from_numpy(x).cuda() if x.size else torch.empty((1, ch), device="cuda"))
            nin.append(x.shape[0])
    caps = [n_out] * n if np.isscalar(n_out) else list(n_out)
    dy = [torch.full((max(c, 1), ch), float("nan"), device="cuda", dtype=torch.float32) for c in caps]
    torch.cuda.synchronize()
    ctxs = (type(streams[0].ctx) * n)(*[s.ctx for s in streams])
    ins = (C.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in dx])
    outs = (C.c_void_p * n)(*[t.data_ptr() for t in dy])
    nin_a = (C.c_int * n)(*nin)
    nout_a = (C.c_int * n)(*caps)
    rat = (C.c_double * n)(*([ratios] * n if np.isscalar(ratios) else list(ratios)))
    res = (Result * n)()
    lib.resampleBatchProcessInterleavedDevice(ctxs, n, ins, nin_a, outs, nout_a, rat, res, C.c_void_p(st.cuda_stream) if st else None)
    if st:
        st.synchronize()
    else:
        for s in streams:
            lib.resampleB200Synchronize(s.ctx)
    torch.cuda.synchronize()
    return [(dy[i][:res[i].output_generated].cpu().numpy(), res[i].input_used, res[i].output_generated) for i in range(n)]


def device_blocks_process(stream, x, block_frames, ratios, capacity):
    """resampleProcessBlocksInterleavedDevice: consecutive blocks of one stream in one launch.
    Returns (y, blocks_done, [(input_used, output_generated)], [positions])."""
    torch = _cuda()
    lib, ch = stream.lib, stream.channels
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ch)
    nb = len(block_frames)
    dxt = torch.from_numpy(x).cuda()
    dyt = torch.full((max(capacity, 1), ch), float("nan"), device="cuda", dtype=torch.float32)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    bf = (C.c_int * nb)(*[int(b) for b in block_frames])
    rt = (C.c_double * nb)(*[float(r) for r in ratios])
    res = (Result * nb)()
    pos = (C.c_double * nb)()
    done = lib.resampleProcessBlocksInterleavedDevice(stream.ctx, dxt.data_ptr(), bf, rt, nb, dyt.data_ptr(), capacity, res, pos,
                                                      C.c_void_p(st.cuda_stream))
    st.synchronize()
    made = sum(res[i].output_generated for i in range(done))
    return (dyt[:made].cpu().numpy(), done, [(res[i].input_used, res[i].output_generated) for i in range(done)],
            [pos[i] for i in range(done)])


def device_process(stream, x, n_out, ratio, planar=False, scattered=False):
    """resampleProcessInterleavedDevice / resampleProcessDevice (planar; `scattered` puts every plane in its own
    allocation at irregular distances, so the library needs its per-channel pointer table)."""
    torch = _cuda()
    lib, ch = stream.lib, stream.channels
    st = torch.cuda.Stream()
    if x is None:
        n_in = -1
    else:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ch)
        n_in = x.shape[0]
    if not planar:
        dxt = torch.from_numpy(x).cuda() if x is not None and x.size else None
        dyt = torch.full((max(n_out, 1), ch), float("nan"), device="cuda", dtype=torch.float32)
        torch.cuda.synchronize()
        res = lib.resampleProcessInterleavedDevice(stream.ctx, dxt.data_ptr() if dxt is not None else None, n_in, dyt.data_ptr(), n_out,
                                                   ratio, C.c_void_p(st.cuda_stream))
        st.synchronize()
        return dyt[:res.output_generated].cpu().numpy(), res.input_used, res.output_generated
    if scattered:
        pad = [torch.empty(17 + 13 * c, device="cuda") for c in range(ch)]           # keeps the planes at irregular distances
        din = [torch.from_numpy(np.ascontiguousarray(x[:, c])).cuda() for c in range(ch)] if x is not None else None
        dout = [torch.full((max(n_out, 1) + 5 * c,), float("nan"), device="cuda", dtype=torch.float32) for c in range(ch)]
        del pad
    else:
        tin = torch.from_numpy(np.ascontiguousarray(x.T)).cuda() if x is not None else None
        tout = torch.full((ch, max(n_out, 1)), float("nan"), device="cuda", dtype=torch.float32)
        din = [tin[c] for c in range(ch)] if tin is not None else None
        dout = [tout[c] for c in range(ch)]
    torch.cuda.synchronize()
    in_arr = (C.c_void_p * ch)(*[t.data_ptr() for t in din]) if din is not None else None
    out_arr = (C.c_void_p * ch)(*[t.data_ptr() for t in dout])
    res = lib.resampleProcessDevice(stream.ctx, in_arr, n_in, out_arr, n_out, ratio, C.c_void_p(st.cuda_stream))
    st.synchronize()
    y = np.stack([dout[c][:res.output_generated].cpu().numpy() for c in range(ch)], axis=1) if res.output_generated else np.zeros((0, ch), np.float32)
    return y, res.input_used, res.output_generated


def batch_process(streams, xs, n_out: int, ratio: float):
    """resampleBatchProcessInterleaved over product streams: returns [(y, input_used, output_generated)] per stream."""
    lib, n = streams[0].lib, len(streams)
    ch = streams[0].channels
    xs = [np.ascontiguousarray(x, dtype=np.float32).reshape(-1, ch) for x in xs]
    outs = [np.zeros((max(n_out, 1), ch), np.float32) for _ in range(n)]
    ctxs = (type(streams[0].ctx) * n)(*[s.ctx for s in streams])
    ins = (f32p * n)(*[_ptr(x) for x in xs])
    out_arr = (f32p * n)(*[_ptr(o) for o in outs])
    nin = (C.c_int * n)(*[x.shape[0] for x in xs])
    nout = (C.c_int * n)(*([n_out] * n))
    ratios = (C.c_double * n)(*([ratio] * n))
    res = (Result * n)()
    lib.resampleBatchProcessInterleaved(ctxs, n, ins, nin, out_arr, nout, ratios, res)
    return [(outs[i][:res[i].output_generated].copy(), res[i].input_used, res[i].output_generated) for i in range(n)]


# ---------------------------------------------------------------------------
# uniform Stream facade


class _ApiStream:
    """A context of a library that speaks the reference C API (reference build or product)."""

    def __init__(self, lib, channels, taps, filters, lowpass_ratio=0.0, flags=SUBSAMPLE_INTERPOLATE | BLACKMAN_HARRIS,
                 fixed=None):
        self.lib, self.channels = lib, channels
        if fixed is None:
            self.ctx = lib.resampleInit(channels, taps, filters, lowpass_ratio, flags)
        else:
            src, dst, lowpass_hz = fixed
            self.ctx = lib.resampleFixedRatioInit(channels, taps, filters, float(src), float(dst), int(lowpass_hz), flags)
        if not self.ctx:
            raise ValueError("init returned NULL")

    # -- state -----------------------------------------------------------
    def advance(self, delta): self.lib.resampleAdvancePosition(self.ctx, delta)
    def position(self): return self.lib.resampleGetPosition(self.ctx)
    def reset(self): self.lib.resampleReset(self.ctx)
    def num_filters(self): return self.lib.resampleGetNumFilters(self.ctx)
    def lowpass_ratio(self): return self.lib.resampleGetLowpassRatio(self.ctx)
    def interpolation_used(self): return self.lib.resampleInterpolationUsed(self.ctx)
    def required_input(self, n_out, ratio): return self.lib.resampleGetRequiredSamples(self.ctx, n_out, ratio)
    def expected_output(self, n_in, ratio): return self.lib.resampleGetExpectedOutput(self.ctx, n_in, ratio)
    def taps(self): return self.ctx.contents.numTaps

    def bank(self):
        c = self.ctx.contents
        return np.stack([np.ctypeslib.as_array(c.filters[i], shape=(c.numTaps,)).copy()
                         for i in range(c.numFilters + 1)])

    def close(self):
        if self.ctx:
            self.lib.resampleFree(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- processing --------------------------------------------------------
    def process(self, x: np.ndarray | None, n_out: int, ratio: float, *, flush_after=False, planar=False):
        """x: (frames, channels) float32, or None for an explicit flush call (numInputFrames = -1)."""
        ch = self.channels
        dt, f32p = _sample(self.lib)
        if x is None:
            n_in = -1
        else:
            x = np.ascontiguousarray(x, dtype=dt).reshape(-1, ch)
            n_in = x.shape[0]
        if not planar:
            out = np.zeros((max(n_out, 1), ch), dt)
            fn = self.lib.resampleProcessAndFlushInterleaved if flush_after else self.lib.resampleProcessInterleaved
            res = fn(self.ctx, _ptr(x) if x is not None else None, n_in, _ptr(out), n_out, ratio)
            return out[:res.output_generated].copy(), res.input_used, res.output_generated
        planes_in = np.ascontiguousarray(x.T) if x is not None else None
        planes_out = np.zeros((ch, max(n_out, 1)), dt)
        in_arr = (f32p * ch)(*[_ptr(planes_in[c]) for c in range(ch)]) if x is not None else None
        out_arr = (f32p * ch)(*[_ptr(planes_out[c]) for c in range(ch)])
        fn = self.lib.resampleProcessAndFlush if flush_after else self.lib.resampleProcess
        res = fn(self.ctx, in_arr, n_in, out_arr, n_out, ratio)
        return planes_out[:, :res.output_generated].T.copy(), res.input_used, res.output_generated


class _OracleStream:
    def __init__(self, lib, channels, taps, filters, lowpass_ratio=0.0, flags=SUBSAMPLE_INTERPOLATE | BLACKMAN_HARRIS,
                 fixed=None):
        self.lib, self.channels, self._taps = lib, channels, taps
        if fixed is None:
            self.ctx = lib.oracle_init(channels, taps, filters, lowpass_ratio, flags)
        else:
            src, dst, lowpass_hz = fixed
            self.ctx = lib.oracle_fixed_ratio_init(channels, taps, filters, float(src), float(dst), int(lowpass_hz), flags)
        if not self.ctx:
            raise ValueError("init returned NULL")

    class _Hdr(C.Structure):
        _fields_ = [("channels", C.c_int), ("taps", C.c_int), ("phases", C.c_int), ("flags", C.c_int),
                    ("ring_len", C.c_int), ("write_index", C.c_int), ("read_pos", C.c_double),
                    ("fixed_ratio", C.c_double), ("lowpass_ratio", C.c_double)]

    def _hdr(self): return C.cast(self.ctx, C.POINTER(self._Hdr)).contents
    def advance(self, delta): self.lib.oracle_advance(self.ctx, delta)
    def position(self): return self.lib.oracle_position(self.ctx)
    def reset(self): self.lib.oracle_reset(self.ctx)
    def num_filters(self): return self._hdr().phases
    def lowpass_ratio(self): return self._hdr().lowpass_ratio
    def interpolation_used(self): return self._hdr().flags & SUBSAMPLE_INTERPOLATE
    def required_input(self, n_out, ratio): return self.lib.oracle_required_input(self.ctx, n_out, ratio)
    def expected_output(self, n_in, ratio): return self.lib.oracle_expected_output(self.ctx, n_in, ratio)
    def taps(self): return self._taps

    def bank(self):
        h = self._hdr()
        return np.stack([np.ctypeslib.as_array(self.lib.oracle_bank_row(self.ctx, i), shape=(h.taps,)).copy()
                         for i in range(h.phases + 1)])

    def close(self):
        if self.ctx:
            self.lib.oracle_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def process(self, x, n_out, ratio, *, flush_after=False, planar=False):
        ch = self.channels
        dt, f32p = _sample(self.lib)
        if x is None:
            n_in = -1
        else:
            x = np.ascontiguousarray(x, dtype=dt).reshape(-1, ch)
            n_in = x.shape[0]
        if not planar:
            out = np.zeros((max(n_out, 1), ch), dt)
            fn = self.lib.oracle_process_flush_interleaved if flush_after else self.lib.oracle_process_interleaved
            res = fn(self.ctx, _ptr(x) if x is not None else None, n_in, _ptr(out), n_out, ratio)
            return out[:res.output_generated].copy(), res.input_used, res.output_generated
        planes_in = np.ascontiguousarray(x.T) if x is not None else None
        planes_out = np.zeros((ch, max(n_out, 1)), dt)
        in_arr = (f32p * ch)(*[_ptr(planes_in[c]) for c in range(ch)]) if x is not None else None
        out_arr = (f32p * ch)(*[_ptr(planes_out[c]) for c in range(ch)])
        fn = self.lib.oracle_process_flush_planar if flush_after else self.lib.oracle_process_planar
        res = fn(self.ctx, in_arr, n_in, out_arr, n_out, ratio)
        return planes_out[:, :res.output_generated].T.copy(), res.input_used, res.output_generated


_cache: dict = {}


def oracle():
    if "oracle" not in _cache:
        _cache["oracle"] = load_oracle()
    return _cache["oracle"]


def reference():
    if "reference" not in _cache:
        _cache["reference"] = load_reference()
    return _cache["reference"]


def product():
    if "product" not in _cache:
        _cache["product"] = load_product()
    return _cache["product"]


class OracleState(C.Structure):
    """public head of OracleResampler (oracle/art_oracle.h)"""
    _fields_ = [("channels", C.c_int), ("taps", C.c_int), ("phases", C.c_int), ("flags", C.c_int),
                ("ring_len", C.c_int), ("write_index", C.c_int), ("read_pos", C.c_double),
                ("fixed_ratio", C.c_double), ("lowpass_ratio", C.c_double),
                ("bank", f32p), ("ring", f32p)]


def oracle_compact_ring(o):
    """Move the newest T samples of every channel to the front of the oracle's ring and shift its indices accordingly
    (what resampler.c:497-503 does when the ring is full, done here at an arbitrary fill level): the stream is unchanged."""
    st = C.cast(o.ctx, C.POINTER(OracleState)).contents
    T, I = st.taps, st.write_index
    drop = I - T
    for c in range(st.channels):
        base = c * st.ring_len
        vals = [st.ring[base + drop + i] for i in range(T)]
        for i in range(T):
            st.ring[base + i] = vals[i]
    st.write_index = T
    st.read_pos -= drop


def oracle_stream(*a, **k): return _OracleStream(oracle(), *a, **k)
def reference_stream(*a, **k): return _ApiStream(reference(), *a, **k)
def product_stream(*a, **k): return _ApiStream(product(), *a, **k)


# the PATH_WIDTH=64 builds (double samples) of the three libraries
def oracle64():
    if "oracle64" not in _cache:
        _cache["oracle64"] = load_oracle(True)
    return _cache["oracle64"]


def reference64():
    if "reference64" not in _cache:
        _cache["reference64"] = load_reference(True)
    return _cache["reference64"]


def product64():
    if "product64" not in _cache:
        _cache["product64"] = load_product(True)
    return _cache["product64"]


def oracle_stream64(*a, **k): return _OracleStream(oracle64(), *a, **k)
def reference_stream64(*a, **k): return _ApiStream(reference64(), *a, **k)
def product_stream64(*a, **k): return _ApiStream(product64(), *a, **k)


def artest_noise(count: int, state: int = 0x3141592653589793):
    """artest.c:744-754 via the oracle's restatement; returns (samples, new_state)."""
    st = C.c_ulonglong(state)
    out = np.empty(count, np.float32)
    oracle().oracle_noise(C.byref(st), _ptr(out), count)
    return out, st.value


def peak_error(a: np.ndarray, b: np.ndarray) -> float:
    """max|a-b| / peak(b) -- the parity measure of BASELINE.md section 2."""
    if a.shape != b.shape:
        return float("inf")
    if a.size == 0:
        return 0.0
    peak = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / (peak if peak > 0 else 1.0)
