"""GPU parity tests of the DEVICE-POINTER entry points of include/resampler_b200.h -- the launches every throughput
number of bench.py is quoted on:

  resampleBatchProcessInterleavedDevice    many contexts in ONE launch (the multi-job branch of all three kernels: job array,
                                           tile -> job lookup, filter-table de-duplication)
  resampleProcessBlocksInterleavedDevice   consecutive blocks of one stream with per-block ratios in one launch (ASRC)
  resampleProcessInterleavedDevice         device twin of resampleProcessInterleaved (resampler.c:550)
  resampleProcessDevice                    device twin of resampleProcess (resampler.c:433), uniform and scattered planes

Same bar as everywhere: input_used / output_generated / resampleGetPosition bit-identical to the oracle, samples within
1e-6 of the oracle's peak.  torch only provides device memory and streams.
"""
import numpy as np
import pytest

import artlibs as A

pytestmark = pytest.mark.gpu

TOL = 1e-6
BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS


@pytest.fixture
def lib():
    lib = A.product()
    yield lib
    lib.resampleB200SetTensorPath(1)


def _streams(n, ch, taps, filters, advance=True, **kw):
    gs = [A.product_stream(ch, taps, filters, **kw) for _ in range(n)]
    os_ = [A.oracle_stream(ch, taps, filters, **kw) for _ in range(n)]
    if advance:
        for g, o in zip(gs, os_):
            g.advance(taps / 2); o.advance(taps / 2)
    return gs, os_


def _compare_batch(gs, os_, xs, caps, ratios, got, tol=TOL):
    caps = [caps] * len(gs) if np.isscalar(caps) else caps
    ratios = [ratios] * len(gs) if np.isscalar(ratios) else ratios
    worst = 0.0
    for i, (g, o) in enumerate(zip(gs, os_)):
        y, used, made = got[i]
        yo, uo, mo = o.process(xs[i], caps[i], ratios[i])
        assert (used, made) == (uo, mo), f"stream {i}: counts {(used, made)} vs oracle {(uo, mo)}"
        assert g.position() == o.position(), f"stream {i}: position"
        err = A.peak_error(y, yo)
        assert err <= tol, f"stream {i}: max|d|/peak = {err:.3g}"
        worst = max(worst, err)
    return worst


# kernel under test -> (tensor mode, ratio): 160/147 is periodic (FFMA form in mode 0, tcgen05 form in mode 2),
# an irrational ratio takes the any-ratio kernel
KERNELS = {"ffma": (0, 48000 / 44100, 1), "tensor": (2, 48000 / 44100, 2), "generic": (0, 0.9123456789, 0)}


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_batch_of_64_contexts_in_lock_step(lib, kernel):
    """the bench.py launch: 64 stereo preset -3 contexts, identical state, one launch (they share ONE filter table)"""
    mode, ratio, path = KERNELS[kernel]
    lib.resampleB200SetTensorPath(mode)
    gs, os_ = _streams(64, 2, 380, 380)
    rng = np.random.default_rng(300)
    for step in range(2):                                   # state carries over from launch to launch
        xs = [rng.uniform(-0.5, 0.5, (6000, 2)).astype(np.float32) for _ in gs]
        before = A.path_counts(lib)
        got = A.device_batch_process(gs, xs, 8000, ratio)
        after = A.path_counts(lib)
        assert after[path] - before[path] == 1 and sum(after) - sum(before) == 1, "not one launch of the expected kernel"
        _compare_batch(gs, os_, xs, 8000, ratio, got)


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_batch_of_desynchronised_contexts(lib, kernel):
    """70 contexts in 70 different states (more filter tables than the de-duplication search looks at), ragged lengths,
    empty calls, an output-limited call and a flush -- all in one launch"""
    mode, ratio, path = KERNELS[kernel]
    lib.resampleB200SetTensorPath(mode)
    n = 70
    gs, os_ = _streams(n, 2, 380, 380)
    rng = np.random.default_rng(301)
    for i, (g, o) in enumerate(zip(gs, os_)):
        if i % 7 != 3:                                      # every 7th stays fresh; the rest are all at different positions
            x = rng.uniform(-0.5, 0.5, (500 + 41 * i, 2)).astype(np.float32)
            g.process(x, 6000, ratio); o.process(x, 6000, ratio)
    xs, caps = [], []
    for i in range(n):
        frames = 5000 + 97 * i
        cap = 9000
        if i == 5: frames = 0
        if i == 11: cap = 2000                              # output-limited: input left unconsumed
        if i == 12: cap = 0
        xs.append(None if i == 20 else rng.uniform(-0.5, 0.5, (frames, 2)).astype(np.float32))
        caps.append(cap)
    before = A.path_counts(lib)
    got = A.device_batch_process(gs, xs, caps, ratio)
    after = A.path_counts(lib)
    assert after[path] - before[path] == 1 and sum(after) - sum(before) == 1
    _compare_batch(gs, os_, xs, caps, ratio, got)
    # and once more from the states that launch left behind (the flushed stream now ignores input)
    xs = [rng.uniform(-0.5, 0.5, (4000, 2)).astype(np.float32) for _ in range(n)]
    got = A.device_batch_process(gs, xs, 6000, ratio)
    _compare_batch(gs, os_, xs, 6000, ratio, got)


def test_batch_with_per_context_ratios(lib):
    """different ratios in one batch cannot share the periodic form: the any-ratio kernel takes the whole launch"""
    lib.resampleB200SetTensorPath(1)
    n = 16
    gs, os_ = _streams(n, 3, 156, 320)
    rng = np.random.default_rng(302)
    ratios = [0.5 + 0.11 * i for i in range(n)]
    xs = [rng.uniform(-0.5, 0.5, (3000 + 10 * i, 3)).astype(np.float32) for i in range(n)]
    caps = [int(x.shape[0] * r) + 200 for x, r in zip(xs, ratios)]
    before = A.path_counts(lib)
    got = A.device_batch_process(gs, xs, caps, ratios, use_stream=False)       # NULL stream = the lead context's own
    after = A.path_counts(lib)
    assert after[0] - before[0] == 1
    _compare_batch(gs, os_, xs, caps, ratios, got)


@pytest.mark.parametrize("mode", [0, 2])
def test_batch_many_channels_and_other_ratios(lib, mode):
    """8-channel contexts (tensor form: planar scratch in HBM), 147/160 downsampling with lowpass = BASELINE config 4's shape"""
    lib.resampleB200SetTensorPath(mode)
    rng = np.random.default_rng(303)
    for ch, taps, filters, src, dst, lp in [(8, 156, 320, 96000, 44100, 20000), (2, 380, 380, 48000, 44100, 20000)]:
        gs, os_ = _streams(9, ch, taps, filters, lowpass_ratio=lp * 2.0 / src)
        ratio = dst / src
        for step in range(2):
            xs = [rng.uniform(-0.5, 0.5, (9000 + 500 * (i % 3) * step, ch)).astype(np.float32) for i in range(9)]
            got = A.device_batch_process(gs, xs, 9000, ratio)
            _compare_batch(gs, os_, xs, 9000, ratio, got)


@pytest.mark.parametrize("blocks,frames", [(256, 4096), (1024, 480)])
def test_asrc_block_sequence(lib, blocks, frames):
    """BASELINE config 5 through the block API: 8 channels, preset -2, ratio swept +/-100 ppm per block; per-block counts
    and positions must equal what per-call resampleProcessInterleaved + resampleGetPosition (resampler.c:965-968) give"""
    filters, taps = A.PRESETS[2]
    ch = 8
    g = A.product_stream(ch, taps, filters, 0.0)
    o = A.oracle_stream(ch, taps, filters, 0.0)
    g.advance(taps / 2); o.advance(taps / 2)
    rng = np.random.default_rng(304 + blocks)
    x = rng.uniform(-0.5, 0.5, (blocks * frames, ch)).astype(np.float32)
    ratios = [1.0 + 1e-4 * np.sin(2 * np.pi * k / 64.0) for k in range(blocks)]
    cap = blocks * (frames + 2) + taps
    before = A.path_counts(lib)
    y, done, counts, positions = A.device_blocks_process(g, x, [frames] * blocks, ratios, cap)
    after = A.path_counts(lib)
    assert done == blocks
    assert after[0] - before[0] == 1 and sum(after) - sum(before) == 1, "the sequence must be ONE launch of the any-ratio kernel"
    ys = []
    for b in range(blocks):
        yo, uo, mo = o.process(x[b * frames:(b + 1) * frames], frames + taps, ratios[b])
        assert counts[b] == (uo, mo), f"block {b}: counts"
        assert positions[b] == o.position(), f"block {b}: position"
        ys.append(yo)
    yo = np.concatenate(ys)
    assert A.peak_error(y, yo) <= TOL
    assert g.position() == o.position()
    # the history the sequence leaves behind: an ordinary call continues the stream
    x2 = rng.uniform(-0.5, 0.5, (3000, ch)).astype(np.float32)
    yg2, ug, mg = g.process(x2, 4000, 1.00003)
    yo2, uo, mo = o.process(x2, 4000, 1.00003)
    assert (ug, mg) == (uo, mo) and A.peak_error(yg2, yo2) <= TOL


def test_block_sequence_stops_when_output_is_full_and_handles_endpoints(lib):
    filters, taps = A.PRESETS[1]
    rng = np.random.default_rng(306)
    # (a) capacity for ~2.5 blocks: two are completed, the third is not started, state equals two per-call blocks
    g = A.product_stream(2, taps, filters, 0.0); o = A.oracle_stream(2, taps, filters, 0.0)
    x = rng.uniform(-0.5, 0.5, (4 * 1000, 2)).astype(np.float32)
    y, done, counts, positions = A.device_blocks_process(g, x, [1000] * 4, [1.5] * 4, 3750)
    assert done == 2
    yo = np.concatenate([o.process(x[b * 1000:(b + 1) * 1000], 3000, 1.5)[0] for b in range(2)])
    assert g.position() == o.position() and A.peak_error(y, yo) <= TOL
    # (b) endpoint extrapolation: the blocks up to the first output go one at a time, then one launch
    flags = BH_INTERP | A.EXTRAPOLATE_ENDPOINTS
    g = A.product_stream(2, taps, filters, 0.0, flags=flags); o = A.oracle_stream(2, taps, filters, 0.0, flags=flags)
    g.advance(taps / 2); o.advance(taps / 2)
    t = np.arange(6 * 700)[:, None]
    x = (0.4 * np.sin(0.05 * t + np.arange(2)) + rng.normal(0, 0.01, (6 * 700, 2))).astype(np.float32)
    y, done, counts, positions = A.device_blocks_process(g, x, [700] * 6, [1.1 + 0.001 * b for b in range(6)], 6000)
    assert done == 6
    ys = []
    for b in range(6):
        yo, uo, mo = o.process(x[b * 700:(b + 1) * 700], 2000, 1.1 + 0.001 * b)
        assert counts[b] == (uo, mo) and positions[b] == o.position()
        ys.append(yo)
    assert A.peak_error(y, np.concatenate(ys)) <= TOL


@pytest.mark.parametrize("mode,ratio", [(0, 48000 / 44100), (2, 48000 / 44100), (0, 0.7071067811865476)])
def test_single_context_device_calls(lib, mode, ratio):
    """resampleProcessInterleavedDevice / resampleProcessDevice (uniform planes, scattered planes): a stream fed call by
    call through each, flush included, against the oracle; the three layouts agree bit for bit"""
    lib.resampleB200SetTensorPath(mode)
    ch, taps, filters = 3, 380, 380
    rng = np.random.default_rng(307)
    variants = [dict(planar=False), dict(planar=True), dict(planar=True, scattered=True)]
    gs = [A.product_stream(ch, taps, filters, 0.0) for _ in variants]
    o = A.oracle_stream(ch, taps, filters, 0.0)
    for s in gs + [o]:
        s.advance(taps / 2)
    for n, cap in [(9000, 12000), (1, 10), (0, 10), (12000, 4000), (8200, 12000), (None, 2000), (100, 200)]:
        x = None if n is None else rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
        yo, uo, mo = o.process(x, cap, ratio)
        outs = []
        for g, kw in zip(gs, variants):
            y, u, m = A.device_process(g, x, cap, ratio, **kw)
            assert (u, m) == (uo, mo) and g.position() == o.position(), kw
            assert A.peak_error(y, yo) <= TOL, kw
            outs.append(y)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("ch", [8, 10])
def test_host_call_of_many_channels_in_pipelined_pieces(lib, ch):
    """a host interleaved call large enough to be cut into pipelined pieces (>= 2^21 output samples), many channels, tensor
    form: pieces after the first start at an output offset; 8 channels run as tiles of four channels straight from the interleaved
    block, 10 channels go through planar scratch (which must transpose only what each piece owns) -- against the FFMA form of
    the same call and, on its first stretch, against the oracle"""
    taps, filters = 156, 320
    ratio = 48000 / 44100
    n = 300_000                                              # -> 326 k frames x 8 (10) channels = 2.6 M (3.3 M) output samples: 2-3 pieces
    rng = np.random.default_rng(308)
    x = rng.uniform(-0.5, 0.5, (n, ch)).astype(np.float32)
    cap = int(n * ratio) + 1000
    outs = []
    for mode in (0, 2):
        lib.resampleB200SetTensorPath(mode)
        g = A.product_stream(ch, taps, filters, 0.0)
        g.advance(taps / 2)
        before = lib.resampleB200TensorLaunches()
        y, u, m = g.process(x, cap, ratio)
        assert u == n
        assert (lib.resampleB200TensorLaunches() - before >= 2) == (mode == 2), "expected several tensor launches (pieces)"
        outs.append((y, g.position()))
    assert outs[0][1] == outs[1][1]
    assert A.peak_error(outs[1][0], outs[0][0]) <= 3e-7
    o = A.oracle_stream(ch, taps, filters, 0.0)
    o.advance(taps / 2)
    yo, _, _ = o.process(x[:30000], cap, ratio)
    assert A.peak_error(outs[1][0][:yo.shape[0] - 200], yo[:-200]) <= TOL


def test_failures_are_reported_not_fatal(lib):
    """a batch that mixes configurations is refused: message, zero counts, positions untouched -- and the process lives"""
    a = A.product_stream(2, 48, 48, 0.0)
    b = A.product_stream(2, 156, 320, 0.0)
    pa, pb = a.position(), b.position()
    x = np.zeros((100, 2), np.float32)
    lib.resampleB200LastError(1)
    got = A.device_batch_process([a, b], [x, x], 300, 1.5)
    assert all(u == 0 and m == 0 for _, u, m in got)
    assert (a.position(), b.position()) == (pa, pb)
    msg = lib.resampleB200LastError(1)
    assert msg and b"one configuration" in msg
    assert lib.resampleB200LastError(0) is None
    y, u, m = a.process(x, 300, 1.5)                        # the contexts are still usable
    assert u == 100
