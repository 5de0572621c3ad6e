"""GPU parity tests of the float <-> integer stages (include/decimator.h; reference decimator.c) through the C ABI.
Integer work: output BYTES, clipped-sample counts and the per-channel state must be bit-identical to the oracle
(oracle/art_oracle.c, itself pinned bit-for-bit to the compiled reference in tests/test_oracle_golden.py)."""
import ctypes as C

import numpy as np
import pytest

import artlibs as A
import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
FLAGS = [0, 0x1, 0x2, 0x4, 0x100, 0x200, 0x400, 0x800, 0x2 | 0x800, 0x1 | 0x200, 0x4 | 0x100]


def _oracle():
    ol = A.oracle()
    ol.oracle_decimate_init.restype = C.c_void_p
    ol.oracle_decimate_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    ol.oracle_decimate_free.argtypes = [C.c_void_p]
    ol.oracle_decimate_interleaved.argtypes = [C.c_void_p, A.f32p, C.c_int, C.c_char_p]
    ol.oracle_decimate_interleaved.restype = C.c_int
    ol.oracle_float_integers.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_int, A.f32p, C.c_int]
    ol.oracle_float_integers.restype = None
    return ol


@pytest.mark.parametrize("flags", FLAGS)
def test_decimate_is_bit_identical(flags):
    """every dither / shaping mode, 8-, 12-, 16-, 24-bit output (24 in a 32-bit container too), clipping gains, several
    calls in a row (generator, feedback and shaper state carry over), interleaved and planar entry points"""
    pkg = entry.load_package(); lib = pkg.load(); ol = _oracle()
    rng = np.random.default_rng(flags + 50)
    for ch, bits, bytes_, gain, rate in [(2, 16, 2, 1.0, 44100), (1, 8, 1, 0.9, 48000), (3, 24, 3, 1.3, 96000), (2, 24, 4, 1.0, 32000), (5, 12, 2, 2.5, 12345)]:
        g = lib.decimateInit(ch, bits, bytes_, gain, rate, flags)
        gp = lib.decimateInit(ch, bits, bytes_, gain, rate, flags)
        o = ol.oracle_decimate_init(ch, bits, bytes_, gain, rate, flags)
        for n in (777, 1, 0, 2500, 9):
            x = rng.uniform(-1.0, 1.0, (n, ch)).astype(np.float32)
            bg, bo = C.create_string_buffer(n * ch * bytes_ + 8), C.create_string_buffer(n * ch * bytes_ + 8)
            cg = lib.decimateProcessInterleavedLE(g, x.ctypes.data_as(A.f32p), n, C.cast(bg, C.c_void_p))
            co = ol.oracle_decimate_interleaved(o, x.ctypes.data_as(A.f32p), n, bo)
            assert cg == co, (ch, bits, flags, n, "clipped sample counts differ")
            assert bg.raw == bo.raw, (ch, bits, flags, n, "bytes differ")
            # the planar entry point: same samples channel by channel
            planes = np.ascontiguousarray(x.T)
            outs = [C.create_string_buffer(n * bytes_ + 8) for _ in range(ch)]
            ins = (A.f32p * ch)(*[planes[c].ctypes.data_as(A.f32p) for c in range(ch)])
            oarr = (C.c_void_p * ch)(*[C.cast(b, C.c_void_p) for b in outs])
            cp = lib.decimateProcessLE(gp, ins, n, oarr)
            assert cp == co
            want = np.frombuffer(bo.raw[:n * ch * bytes_], np.uint8).reshape(n, ch, bytes_)
            for c in range(ch):
                assert np.array_equal(np.frombuffer(outs[c].raw[:n * bytes_], np.uint8).reshape(n, bytes_), want[:, c, :])
        lib.decimateFree(g); lib.decimateFree(gp); ol.oracle_decimate_free(o)


def test_decimate_many_contexts_in_one_launch():
    """decimateBatchProcessInterleavedLE: 40 contexts of different shapes and modes in one launch, twice"""
    pkg = entry.load_package(); lib = pkg.load(); ol = _oracle()
    rng = np.random.default_rng(60)
    shapes = [(1 + i % 4, [8, 16, 24, 20][i % 4], [1, 2, 3, 4][i % 4], 0.8 + 0.05 * i, [44100, 48000, 96000][i % 3], FLAGS[i % len(FLAGS)]) for i in range(40)]
    gs = [lib.decimateInit(*s) for s in shapes]
    os_ = [ol.oracle_decimate_init(*s) for s in shapes]
    for rnd in range(2):
        frames = [int(rng.integers(0, 3000)) for _ in shapes]
        xs = [rng.uniform(-1.1, 1.1, (f, s[0])).astype(np.float32) for f, s in zip(frames, shapes)]
        bufs = [C.create_string_buffer(f * s[0] * s[2] + 8) for f, s in zip(frames, shapes)]
        n = len(shapes)
        carr = (C.POINTER(pkg.Decimate) * n)(*gs)
        ins = (A.f32p * n)(*[x.ctypes.data_as(A.f32p) for x in xs])
        outs = (C.c_void_p * n)(*[C.cast(b, C.c_void_p) for b in bufs])
        fr = (C.c_int * n)(*frames)
        clips = (C.c_int * n)()
        total = lib.decimateBatchProcessInterleavedLE(carr, n, ins, fr, outs, clips)
        want_total = 0
        for i, s in enumerate(shapes):
            bo = C.create_string_buffer(frames[i] * s[0] * s[2] + 8)
            co = ol.oracle_decimate_interleaved(os_[i], xs[i].ctypes.data_as(A.f32p), frames[i], bo)
            assert clips[i] == co and bufs[i].raw == bo.raw, (rnd, i, s)
            want_total += co
        assert total == want_total
    for g in gs: lib.decimateFree(g)
    for o in os_: ol.oracle_decimate_free(o)


def test_decimate_device_pointers_and_float_integers_round_trip():
    import torch
    pkg = entry.load_package(); lib = pkg.load(); ol = _oracle()
    rng = np.random.default_rng(61)
    st = torch.cuda.Stream()
    for bits, bytes_, stride, gain in [(8, 1, 1, 1.0), (16, 2, 1, 1.0), (16, 2, 3, 0.5), (24, 3, 1, 1.0), (24, 4, 2, 1.7), (20, 3, 1, 1.0)]:
        n = 5000
        raw = rng.integers(0, 256, n * stride * bytes_, dtype=np.uint8)
        a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
        lib.floatIntegersLE(raw.ctypes.data_as(C.c_void_p), gain, bits, bytes_, stride, a.ctypes.data_as(A.f32p), n)
        ol.oracle_float_integers(raw.tobytes(), gain, bits, bytes_, stride, b.ctypes.data_as(A.f32p), n)
        assert np.array_equal(a, b), (bits, bytes_, stride)
        d_raw = torch.from_numpy(raw).cuda()
        d_out = torch.zeros(n, device="cuda")
        torch.cuda.synchronize()
        lib.floatIntegersLEDevice(d_raw.data_ptr(), gain, bits, bytes_, stride, d_out.data_ptr(), n, C.c_void_p(st.cuda_stream))
        st.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), b)
    # float -> 16 bit on the device, plain rounding: integers -> float -> integers is the identity
    g = lib.decimateInit(2, 16, 2, 1.0, 44100, 0)
    pcm = rng.integers(-32768, 32768, (4000, 2), dtype=np.int16)
    x = np.zeros((4000, 2), np.float32)
    lib.floatIntegersLE(pcm.ctypes.data_as(C.c_void_p), 1.0, 16, 2, 1, x.ctypes.data_as(A.f32p), 8000)
    d_x = torch.from_numpy(x).cuda()
    d_y = torch.zeros(4000 * 2 * 2, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    clipped = lib.decimateProcessInterleavedLEDevice(g, d_x.data_ptr(), 4000, d_y.data_ptr(), C.c_void_p(st.cuda_stream))
    assert clipped == 0
    assert np.array_equal(d_y.cpu().numpy().view(np.int16).reshape(4000, 2), pcm)
    lib.decimateFree(g)
