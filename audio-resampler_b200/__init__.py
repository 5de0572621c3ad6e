"""audio-resampler_b200 -- Python view of libresampler_b200.so.

The product is the C-ABI shared library built from ``csrc/`` (C host code + sm_100a CUDA
kernels; see ../include/*.h).  This module only locates/builds it and hands out a ctypes
handle with the prototypes attached; tests and bench.py go through that handle, i.e.
through the same boundary a C caller of the reference would use.

The directory name contains a hyphen (it mirrors the upstream project name), so import it
with ``importlib`` -- ``__graft_entry__.load_package()`` does that.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "lib" / "libresampler_b200.so"

# flag values, reference resampler.h:28-38
SUBSAMPLE_INTERPOLATE = 0x1
BLACKMAN_HARRIS = 0x2
INCLUDE_LOWPASS = 0x4
RESAMPLE_MULTITHREADED = 0x8
NO_FILTER_REDUCTION = 0x10
EXTRAPOLATE_ENDPOINTS = 0x40
EXTEND_CONVOLUTION_MATH = 0x100

# (filters, taps) of the reference's quality presets -1..-4, artest.c:154-169
PRESETS = {1: (48, 48), 2: (320, 156), 3: (380, 380), 4: (988, 988)}


class ResampleResult(C.Structure):
    _fields_ = [("input_used", C.c_uint), ("output_generated", C.c_uint)]


class Resample(C.Structure):
    """include/resampler.h -- leading public fields (reference resampler.h:44-48)."""
    _fields_ = [("numChannels", C.c_int), ("numSamples", C.c_int), ("numFilters", C.c_int),
                ("numTaps", C.c_int), ("inputIndex", C.c_int), ("flags", C.c_int),
                ("tempFilter", C.c_void_p), ("outputOffset", C.c_double), ("fixedRatio", C.c_double),
                ("lowpassRatio", C.c_double), ("subsample", C.c_void_p),
                ("buffers", C.c_void_p), ("filters", C.POINTER(C.POINTER(C.c_float))),
                ("device", C.c_void_p), ("prefilterLead", C.c_int), ("prefilterTaps", C.c_void_p), ("plainDevice", C.c_void_p)]


class BiquadCoefficients(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class Decimate(C.Structure):
    """Leading public fields of the decimator context (decimator.h:43-51)."""
    _fields_ = [("numChannels", C.c_int), ("outputBits", C.c_int), ("outputBytes", C.c_int), ("dither_type", C.c_int),
                ("flags", C.c_int), ("outputGain", C.c_double), ("feedback", C.POINTER(C.c_float)),
                ("tpdf_generators", C.POINTER(C.c_uint32)), ("noise_shapers", C.c_void_p)]


class Biquad(C.Structure):
    _fields_ = [("a", C.c_float * 5), ("b", C.c_float * 5), ("x", C.c_float * 4), ("y", C.c_float * 4),
                ("order", C.c_int), ("index", C.c_int)]


# the PATH_WIDTH=64 library (libresampler_b200_64.so): the same structures with double samples
LIB64_PATH = HERE / "lib" / "libresampler_b200_64.so"


class Resample64(C.Structure):
    _fields_ = [(n, (C.POINTER(C.POINTER(C.c_double)) if n == "filters" else t)) for n, t in Resample._fields_]


class BiquadCoefficients64(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a0", "a1", "a2", "a3", "a4", "b1", "b2", "b3", "b4")]


class Decimate64(C.Structure):
    _fields_ = [(n, (C.POINTER(C.c_double) if n == "feedback" else t)) for n, t in Decimate._fields_]


class Biquad64(C.Structure):
    _fields_ = [("a", C.c_double * 5), ("b", C.c_double * 5), ("x", C.c_double * 4), ("y", C.c_double * 4),
                ("order", C.c_int), ("index", C.c_int)]


def build(verbose: bool = False) -> Path:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(HERE), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libresampler_b200.so failed")
    return LIB_PATH


_lib = None
_lib64 = None


def load() -> C.CDLL:
    """dlopen the library and attach the prototypes of include/*.h.  Raises when the
    library is missing: there is no Python or CPU fallback for any of it."""
    global _lib
    if _lib is None:
        _lib = _open(LIB_PATH, C.c_float, Resample, Biquad, BiquadCoefficients, Decimate)
    return _lib


def load64() -> C.CDLL:
    """the PATH_WIDTH=64 build of the same sources: every sample pointer is a double pointer"""
    global _lib64
    if _lib64 is None:
        _lib64 = _open(LIB64_PATH, C.c_double, Resample64, Biquad64, BiquadCoefficients64, Decimate64)
    return _lib64


def _open(path, sample, Resample, Biquad, BiquadCoefficients, Decimate) -> C.CDLL:
    if not path.exists():
        raise RuntimeError(f"{path} not built; run __graft_entry__.build()")
    lib = C.CDLL(str(path), mode=os.RTLD_LOCAL)
    ctx, f32p = C.POINTER(Resample), C.POINTER(sample)
    f32pp, vp, dbl, i32 = C.POINTER(f32p), C.c_void_p, C.c_double, C.c_int
    proto = {
        # include/resampler.h
        "resampleInit": (ctx, [i32, i32, i32, dbl, i32]),
        "resampleFixedRatioInit": (ctx, [i32, i32, i32, dbl, dbl, i32, i32]),
        "resampleProcess": (ResampleResult, [ctx, f32pp, i32, f32pp, i32, dbl]),
        "resampleProcessInterleaved": (ResampleResult, [ctx, f32p, i32, f32p, i32, dbl]),
        "resampleProcessAndFlush": (ResampleResult, [ctx, f32pp, i32, f32pp, i32, dbl]),
        "resampleProcessAndFlushInterleaved": (ResampleResult, [ctx, f32p, i32, f32p, i32, dbl]),
        "resampleGetRequiredSamples": (C.c_uint, [ctx, i32, dbl]),
        "resampleGetExpectedOutput": (C.c_uint, [ctx, i32, dbl]),
        "resampleAdvancePosition": (None, [ctx, dbl]),
        "resampleGetLowpassRatio": (dbl, [ctx]),
        "resampleGetPosition": (dbl, [ctx]),
        "resampleGetNumFilters": (i32, [ctx]),
        "resampleInterpolationUsed": (i32, [ctx]),
        "resampleReset": (None, [ctx]),
        "resampleFree": (None, [ctx]),
        # include/biquad.h
        "biquad_init": (None, [C.POINTER(Biquad), C.POINTER(BiquadCoefficients), dbl]),
        "biquad_lowpass": (None, [C.POINTER(BiquadCoefficients), dbl]),
        "biquad_highpass": (None, [C.POINTER(BiquadCoefficients), dbl]),
        "biquad_apply_buffer": (None, [C.POINTER(Biquad), f32p, i32, i32]),
        "biquad_apply_sample": (sample, [C.POINTER(Biquad), sample]),
        # include/resampler_b200.h (device pointers travel as integers)
        "resampleB200SetDevice": (i32, [i32]),
        "resampleB200GetDeviceCount": (i32, []),
        "resampleB200Synchronize": (None, [ctx]),
        "resampleB200KernelLaunches": (C.c_ulonglong, []),
        "resampleB200LastError": (C.c_char_p, [i32]),
        "resampleB200PathCounts": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
        "resampleB200SetTensorPath": (None, [i32]),
        "resampleB200SetTensorDigits": (None, [i32]),
        "resampleB200AttachPrefilter": (i32, [ctx, C.POINTER(Biquad), i32]),
        "resampleB200TensorLaunches": (C.c_ulonglong, []),
        "resampleB200ProfileEnable": (None, [i32]),
        "resampleB200ProfileCollect": (C.c_ulonglong, [C.POINTER(dbl)]),
        "resampleProcessInterleavedDevice": (ResampleResult, [ctx, vp, i32, vp, i32, dbl, vp]),
        "resampleProcessDevice": (ResampleResult, [ctx, C.POINTER(vp), i32, C.POINTER(vp), i32, dbl, vp]),
        "resampleBatchProcessInterleavedDevice": (None, [C.POINTER(ctx), i32, C.POINTER(vp), C.POINTER(i32),
                                                         C.POINTER(vp), C.POINTER(i32), C.POINTER(dbl),
                                                         C.POINTER(ResampleResult), vp]),
        "resampleBatchProcessInterleaved": (None, [C.POINTER(ctx), i32, C.POINTER(f32p), C.POINTER(i32),
                                                   C.POINTER(f32p), C.POINTER(i32), C.POINTER(dbl),
                                                   C.POINTER(ResampleResult)]),
        "resampleProcessBlocksInterleavedDevice": (i32, [ctx, vp, C.POINTER(i32), C.POINTER(dbl), i32, vp, i32,
                                                          C.POINTER(ResampleResult), C.POINTER(dbl), vp]),
        # include/decimator.h
        "floatIntegersLE": (None, [vp, dbl, i32, i32, i32, f32p, i32]),
        "floatIntegersLEDevice": (None, [vp, dbl, i32, i32, i32, vp, i32, vp]),
        "decimateInit": (C.POINTER(Decimate), [i32, i32, i32, dbl, i32, i32]),
        "decimateProcessLE": (i32, [C.POINTER(Decimate), f32pp, i32, C.POINTER(vp)]),
        "decimateProcessInterleavedLE": (i32, [C.POINTER(Decimate), f32p, i32, vp]),
        "decimateProcessInterleavedLEDevice": (i32, [C.POINTER(Decimate), vp, i32, vp, vp]),
        "decimateBatchProcessInterleavedLE": (i32, [C.POINTER(C.POINTER(Decimate)), i32, C.POINTER(f32p), C.POINTER(i32), C.POINTER(vp), C.POINTER(i32)]),
        "decimateBatchProcessInterleavedLEDevice": (i32, [C.POINTER(C.POINTER(Decimate)), i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(vp), C.POINTER(i32), vp]),
        "decimateFree": (None, [C.POINTER(Decimate)]),
        "biquad_apply_cascade_interleaved": (None, [C.POINTER(C.POINTER(Biquad)), i32, i32, f32p, i32]),
        "biquad_apply_cascade_interleaved_device": (None, [C.POINTER(C.POINTER(Biquad)), i32, i32, vp, i32, vp]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not export what include/*.h declares
        fn.restype, fn.argtypes = res, args
    return lib


EXPORTED_SYMBOLS = [
    "resampleInit", "resampleFixedRatioInit", "resampleProcess", "resampleProcessInterleaved",
    "resampleProcessAndFlush", "resampleProcessAndFlushInterleaved", "resampleGetRequiredSamples",
    "resampleGetExpectedOutput", "resampleAdvancePosition", "resampleGetLowpassRatio", "resampleGetPosition",
    "resampleGetNumFilters", "resampleInterpolationUsed", "resampleReset", "resampleFree",
    "biquad_init", "biquad_lowpass", "biquad_highpass", "biquad_apply_buffer", "biquad_apply_sample",
    "resampleB200SetDevice", "resampleB200GetDeviceCount", "resampleB200Synchronize", "resampleB200KernelLaunches", "resampleB200LastError",
    "resampleB200PathCounts", "resampleB200SetTensorPath", "resampleB200SetTensorDigits", "resampleB200AttachPrefilter", "resampleB200TensorLaunches", "resampleB200ProfileEnable", "resampleB200ProfileCollect",
    "resampleProcessInterleavedDevice", "resampleProcessDevice", "resampleBatchProcessInterleavedDevice",
    "resampleBatchProcessInterleaved", "resampleProcessBlocksInterleavedDevice", "biquad_apply_cascade_interleaved",
    "biquad_apply_cascade_interleaved_device",
    "floatIntegersLE", "floatIntegersLEDevice", "decimateInit", "decimateProcessLE", "decimateProcessInterleavedLE",
    "decimateProcessInterleavedLEDevice", "decimateBatchProcessInterleavedLE", "decimateBatchProcessInterleavedLEDevice", "decimateFree",
]
