"""CPU tests of the product's host logic: the closed-form control loop (csrc/art_plan.h) must give
bit-identical input_used / output_generated / outputOffset / inputIndex to the reference's
frame-by-frame loop (resampler.c:494-535), including ring compactions inside long calls, flushes,
snap, output-limited and input-limited calls.  Runs through tests/shim/plan_shim.c (no GPU)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import artlibs as A

SHIM = Path(__file__).resolve().parent / "shim" / "libplanshim.so"


class St(C.Structure):
    _fields_ = [("P", C.c_double), ("I", C.c_int), ("T", C.c_int), ("F", C.c_int), ("flushed", C.c_int), ("snap", C.c_int)]


class Rs(C.Structure):
    _fields_ = [("used", C.c_uint), ("made", C.c_uint)]


def shim():
    lib = C.CDLL(str(SHIM))
    lib.shim_call.restype = Rs
    lib.shim_call.argtypes = [C.POINTER(St), C.c_int, C.c_int, C.c_double]
    lib.shim_output_pos.restype = C.c_double
    lib.shim_output_pos.argtypes = [C.POINTER(St), C.c_double, C.c_uint, C.POINTER(C.c_int)]
    lib.shim_output_pos_from.restype = C.c_double
    lib.shim_output_pos_from.argtypes = [C.POINTER(St), C.c_double, C.c_uint, C.c_uint, C.POINTER(C.c_int)]
    return lib


class ModelStream:
    """The reference's control loop in a few lines of Python (oracle-side model for this test)."""

    def __init__(self, T, F, snap=False):
        self.T, self.F, self.P, self.I, self.flushed, self.snap = T, F, float(T // 2), T, False, snap

    def call(self, n_in, n_out, ratio):
        T, half, NS, D = self.T, self.T // 2, 16 * self.T, 15 * self.T
        if self.flushed:
            n_in = 0
        if n_in < 0:
            if NS - self.I < half:
                self.P -= D; self.I -= D
            self.flushed = True
            self.I += half
        used = made = 0
        step = 0.0
        while n_out > 0:
            if self.P + step >= self.I - half:
                if n_in <= 0:
                    break
                if self.I == NS:
                    self.P -= D; self.I -= D
                self.I += 1; used += 1; n_in -= 1
            else:
                made += 1
                step = made / ratio
                n_out -= 1
        self.P += step
        if self.snap:
            w = np.floor(self.P)
            self.P = float(w + np.floor((self.P - w) * self.F + 0.5) / self.F)
        return used, made


def _sessions(rng, count):
    for _ in range(count):
        T = int(rng.choice([4, 8, 48, 156, 380, 988, 1024]))
        F = int(rng.integers(1, 1025))
        ratio = float(rng.choice([48000 / 44100, 44100 / 48000, 44100 / 96000, 2.0, 0.5, 1.0, 1.0001, 0.9999,
                                  160 / 147, 8.0, 0.125, float(np.exp(rng.uniform(np.log(0.05), np.log(20))))]))
        yield T, F, ratio


def test_closed_form_matches_python_model():
    lib, rng = shim(), np.random.default_rng(3)
    for T, F, ratio in _sessions(rng, 60):
        snap = bool(rng.integers(0, 2))
        m = ModelStream(T, F, snap)
        s = St(m.P, m.I, T, F, 0, int(snap))
        if rng.random() < 0.6:
            adv = float(T // 2)
            m.P += adv; s.P += adv
        for _ in range(int(rng.integers(1, 12))):
            n_in = int(rng.integers(0, 3000)) if rng.random() < 0.9 else -1
            n_out = int(rng.integers(0, 4000))
            want = m.call(n_in, n_out, ratio)
            got = lib.shim_call(C.byref(s), n_in, n_out, ratio)
            assert (got.used, got.made) == want
            assert s.P == m.P and s.I == m.I


@pytest.mark.skipif(A.reference() is None, reason="oracle/_ref/libartref.so not built")
def test_closed_form_matches_reference_long_calls():
    """Calls long enough for many ring compactions (outputOffset runs far negative inside the call)."""
    lib, rng = shim(), np.random.default_rng(9)
    for T, F, ratio in _sessions(rng, 40):
        r = A.reference_stream(1, T, F, 0.0, flags=3)
        r.advance(T / 2)
        c = r.ctx.contents
        s = St(c.outputOffset, c.inputIndex, T, F, 0, 0)
        for _ in range(4):
            n_in = int(rng.integers(0, 60000))
            n_out = int(rng.integers(0, 60000)) if rng.random() < 0.7 else 10 ** 6
            _, used, made = r.process(np.zeros((n_in, 1), np.float32), n_out, ratio)
            got = lib.shim_call(C.byref(s), n_in, n_out, ratio)
            c = r.ctx.contents
            assert (got.used, got.made) == (used, made)
            assert s.P == c.outputOffset and s.I == c.inputIndex
        _, used, made = r.process(None, 10 ** 6, ratio)
        got = lib.shim_call(C.byref(s), -1, 10 ** 6, ratio)
        assert (got.used, got.made) == (used, made) and s.P == r.ctx.contents.outputOffset


def test_tile_relative_positions_equal_direct_positions():
    """art_output_pos_from (what a CUDA thread evaluates) == art_output_pos for every output."""
    lib, rng = shim(), np.random.default_rng(4)
    for T, F, ratio in _sessions(rng, 30):
        s = St(float(T // 2) + T / 2, T, T, F, 0, 0)
        w1, w2 = C.c_int(), C.c_int()
        n0 = int(rng.integers(0, 100000))
        for n in rng.integers(n0, n0 + 4096, 64):
            a = lib.shim_output_pos(C.byref(s), ratio, int(n), C.byref(w1))
            b = lib.shim_output_pos_from(C.byref(s), ratio, n0, int(n), C.byref(w2))
            assert a == b and w1.value == w2.value


def test_zero_and_degenerate_calls():
    lib = shim()
    s = St(190.0 + 190.0, 380, 380, 380, 0, 0)
    assert tuple(getattr(lib.shim_call(C.byref(s), 0, 0, 1.0), f) for f in ("used", "made")) == (0, 0)
    assert tuple(getattr(lib.shim_call(C.byref(s), 100, 0, 1.0), f) for f in ("used", "made")) == (0, 0)
    r = lib.shim_call(C.byref(s), 0, 100, 1.0)
    assert (r.used, r.made) == (0, 0)
    # ratio 0 on a non-fixed context: the first output (offset2 == 0.0) then nothing, all input eaten
    s2, m = St(24.0, 48, 48, 48, 0, 0), ModelStream(48, 48)
    with np.errstate(divide="ignore"):
        want = m.call(500, 500, 1e-300)
    got = lib.shim_call(C.byref(s2), 500, 500, 1e-300)
    assert (got.used, got.made) == want
