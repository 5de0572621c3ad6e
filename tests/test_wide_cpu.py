"""CPU tests of the PATH_WIDTH=64 variant (reference resampler.h:22-26, Makefile:12-19: every sample is a double).
The restatement oracle/liboracle64.so is pinned against the unmodified reference built with -DPATH_WIDTH=64
(oracle/_ref/libartref64.so); libresampler_b200_64.so must export the same API as the float library."""
import ctypes as C

import numpy as np
import pytest

import artlibs as A
import __graft_entry__ as entry

TOL64 = 1e-13            # of the peak: two orders of summation of <= 988 double products


def _ref64():
    r = A.reference64()
    if r is None:
        pytest.skip("reference build not available")
    return r


@pytest.mark.parametrize("seed", range(6))
def test_oracle64_vs_live_reference64_random_sessions(seed):
    """random contexts (channels, preset, flags, fixed / free ratio), several ragged calls and a flush each: counts and position
    bit-identical, samples within 1e-13 of the peak"""
    _ref64()
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        ch = int(rng.integers(1, 5)); preset = int(rng.integers(1, 5))
        filters, taps = A.PRESETS[preset]
        flags = int(rng.choice([0x3, 0x1, 0x2, 0x0, 0x103]))
        fixed = None
        if rng.random() < 0.4:
            fixed = tuple(rng.choice([(44100, 48000, 0), (48000, 44100, 20000), (96000, 44100, 0), (32000, 48000, 0)]))
            flags |= 0x4 if fixed[2] or fixed[0] > fixed[1] else 0
        lowpass = float(rng.choice([0.0, 0.9, 0.6]))
        ratio = float(rng.choice([48000 / 44100, 44100 / 48000, 0.4593, 2.0, 1.00007, 1.37]))
        o = A.oracle_stream64(ch, taps, filters, lowpass, flags, fixed=fixed)
        r = A.reference_stream64(ch, taps, filters, lowpass, flags, fixed=fixed)
        assert o.num_filters() == r.num_filters() and o.lowpass_ratio() == r.lowpass_ratio()
        if o.interpolation_used() or not fixed:
            o.advance(taps / 2); r.advance(taps / 2)
        for call in range(4):
            n = int(rng.integers(0, 3000))
            x = rng.uniform(-0.5, 0.5, (n, ch))
            cap = int(rng.integers(1, 6000))
            yo, uo, go = o.process(x, cap, ratio, planar=bool(call & 1))
            yr, ur, gr = r.process(x, cap, ratio, planar=bool(call & 1))
            # counts exact.  The position getter is outputOffset + T/2 - inputIndex (resampler.c:965-968): the reference build's
            # -fassociative-math may evaluate it as outputOffset + (T/2 - inputIndex), one rounding fewer, so compare to an ulp
            assert (uo, go) == (ur, gr) and abs(o.position() - r.position()) <= 2e-13
            assert yo.dtype == np.float64 and A.peak_error(yo, yr) <= TOL64
        # the flush, unless the reference's ring is so full that postfillAllChannels compacts it: the flush outputs then read in
        # front of the buffer (heap garbage in the reference and in its restatement; tests/test_gpu_parity.py has the details)
        if 16 * taps - r.ctx.contents.inputIndex >= taps // 2:
            yo, uo, go = o.process(None, 4000, ratio)
            yr, ur, gr = r.process(None, 4000, ratio)
            assert (uo, go) == (ur, gr) and abs(o.position() - r.position()) <= 2e-13 and A.peak_error(yo, yr) <= TOL64


def test_oracle64_extrapolated_endpoints_vs_reference64():
    """EXTRAPOLATE_ENDPOINTS on the wide path: the LPC fit runs on double samples with float coefficients
    (extrapolator.c:20-43); tonal material, as in the float-path test"""
    _ref64()
    t = np.arange(6000)
    x = (0.4 * np.sin(2 * np.pi * 0.013 * t) + 0.2 * np.sin(2 * np.pi * 0.071 * t + 1.0))[:, None] * np.ones((1, 2))
    for preset in (1, 3):
        filters, taps = A.PRESETS[preset]
        o = A.oracle_stream64(2, taps, filters, 0.0, 0x43); r = A.reference_stream64(2, taps, filters, 0.0, 0x43)
        o.advance(taps / 2); r.advance(taps / 2)
        yo, uo, go = o.process(x, 9000, 48000 / 44100, flush_after=True)
        yr, ur, gr = r.process(x, 9000, 48000 / 44100, flush_after=True)
        assert (uo, go) == (ur, gr) and A.peak_error(yo, yr) <= 1e-12


def test_bank64_equals_reference_bank_to_an_ulp():
    _ref64()
    for preset in (1, 3):
        filters, taps = A.PRESETS[preset]
        o = A.oracle_stream64(1, taps, filters, 0.0); r = A.reference_stream64(1, taps, filters, 0.0)
        assert np.max(np.abs(o.bank() - r.bank())) <= 1e-15        # the reference's -fassociative-math build sums the row in another order


def test_wide_library_exports_the_same_api():
    pkg = entry.load_package()
    lib64 = pkg.load64()
    for name in pkg.EXPORTED_SYMBOLS:
        assert hasattr(lib64, name), f"{name} missing from libresampler_b200_64.so"
    assert C.sizeof(pkg.Biquad64) == 152 and pkg.Resample64.filters.offset == 72      # reference biquad.h:31-35 with doubles


def test_wide_library_designs_biquads_like_the_reference64():
    _ref64()
    pkg = entry.load_package(); lib64 = pkg.load64(); ref = A.reference64()
    for kind in ("biquad_lowpass", "biquad_highpass"):
        for f in (0.45 * 44100 / 96000, 0.01, 0.3):
            a, b = pkg.BiquadCoefficients64(), A.BiquadCoefficients64()
            getattr(lib64, kind)(C.byref(a), f); getattr(ref, kind)(C.byref(b), f)
            # the reference build's -fassociative-math regroups (1 - K/Q + K*K): equal to an ulp or two, not to the bit (in the float
            # build the rounding to float hides it, tests/test_abi_cpu.py compares that one exactly)
            for n, _ in a._fields_:
                assert abs(getattr(a, n) - getattr(b, n)) <= 4e-16 * max(1.0, abs(getattr(b, n))), n
