"""Device-resident throughput of the five BASELINE.json configs (and the fixed-ratio variants art.c uses).
Not the contract bench (bench.py is); a survey to see every kernel path at size.  One line per config."""
import ctypes as C, json, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); lib = pkg.load()
dev = torch.device("cuda")
st = torch.cuda.Stream(); sp = C.c_void_p(st.cuda_stream)
PRESET = {1: (48, 48), 2: (320, 156), 3: (380, 380), 4: (988, 988)}


def paths():
    g, p = C.c_ulonglong(), C.c_ulonglong()
    lib.resampleB200PathCounts(C.byref(g), C.byref(p))
    return g.value, p.value


def run(name, ch, preset, src, dst, streams, frames, lowpass_hz=0, fixed=False, flags=3, steps=10):
    filters, taps = PRESET[preset]
    ratio = dst / src
    if fixed:
        ctxs = [lib.resampleFixedRatioInit(ch, taps, filters, float(src), float(dst), lowpass_hz, flags | 4) for _ in range(streams)]
    else:
        ctxs = [lib.resampleInit(ch, taps, filters, lowpass_hz * 2.0 / src, flags) for _ in range(streams)]
    for c in ctxs:
        lib.resampleAdvancePosition(c, taps / 2)
    x = torch.rand((streams, frames, ch), device=dev) - 0.5
    cap = int(frames * ratio) + taps + 16
    y = torch.empty((streams, cap, ch), device=dev)
    n = streams
    ca = (C.POINTER(pkg.Resample) * n)(*ctxs)
    ia = (C.c_void_p * n)(*[x[i].data_ptr() for i in range(n)]); oa = (C.c_void_p * n)(*[y[i].data_ptr() for i in range(n)])
    ni, no = (C.c_int * n)(*([frames] * n)), (C.c_int * n)(*([cap] * n))
    ra = (C.c_double * n)(*([ratio] * n)); res = (pkg.ResampleResult * n)()

    def step():
        lib.resampleBatchProcessInterleavedDevice(ca, n, ia, ni, oa, no, ra, res, sp)
        return int(np.frombuffer(res, dtype=np.uint32)[1::2].sum(dtype=np.int64))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    g0, p0 = paths()
    t0n = lib.resampleB200TensorLaunches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st); made = 0
    for _ in range(steps):
        made += step()
    e1.record(st); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    g1, p1 = paths()
    sps = made * ch / (ms * 1e-3)
    byts = 4.0 * (1.0 + 1.0 / ratio)
    print(json.dumps({"config": name, "Gsamples_per_s": round(sps / 1e9, 2), "ms_per_step": round(ms / steps, 3),
                      "wall_ms_per_step": round(wall / steps, 3), "hbm_frac": round(sps * byts / 6553e9, 4),
                      "kernel": "tensor" if lib.resampleB200TensorLaunches() > t0n else "periodic" if p1 > p0 else "generic", "filters": lib.resampleGetNumFilters(ctxs[0]),
                      "interp": bool(lib.resampleInterpolationUsed(ctxs[0]))}), flush=True)
    for c in ctxs:
        lib.resampleFree(c)
    del x, y
    torch.cuda.empty_cache()


def run_asrc(name, ch, preset, blocks, block_frames, steps=10):
    """config 5: one stream, ratio swept +/-100 ppm per block, all blocks of a step in one launch"""
    import math
    filters, taps = PRESET[preset]
    ctx = lib.resampleInit(ch, taps, filters, 0.0, 3)
    lib.resampleAdvancePosition(ctx, taps / 2)
    total = blocks * block_frames
    x = torch.rand((total, ch), device=dev) - 0.5
    cap = total + blocks * 8 + 1024
    y = torch.empty((cap, ch), device=dev)
    bf = (C.c_int * blocks)(*([block_frames] * blocks))
    ra = (C.c_double * blocks)(*[1.0 + 1e-4 * math.sin(2 * math.pi * k / blocks) for k in range(blocks)])
    res = (pkg.ResampleResult * blocks)(); pos = (C.c_double * blocks)()

    def step():
        done = lib.resampleProcessBlocksInterleavedDevice(ctx, C.c_void_p(x.data_ptr()), bf, ra, blocks,
                                                          C.c_void_p(y.data_ptr()), cap, res, pos, sp)
        assert done == blocks
        return int(np.frombuffer(res, dtype=np.uint32)[1::2].sum(dtype=np.int64))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st); made = 0
    for _ in range(steps):
        made += step()
    host = (time.perf_counter() - t0) * 1e3          # time the host spent planning and enqueueing (no waiting)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lib.resampleB200ProfileEnable(1)
    for _ in range(steps):
        step()
    kms = C.c_double(); nk = lib.resampleB200ProfileCollect(C.byref(kms)); lib.resampleB200ProfileEnable(0)
    sps = made * ch / (ms * 1e-3)
    print(json.dumps({"config": name, "Gsamples_per_s": round(sps / 1e9, 2), "ms_per_step": round(ms / steps, 3),
                      "host_ms_per_step": round(host / steps, 3), "kernel_ms_per_step": round(kms.value / steps, 3),
                      "kernel_Gsamples_per_s": round(made * ch / (kms.value * 1e-3) / 1e9, 2) if kms.value else None,
                      "hbm_frac": round(sps * 8.0 / 6553e9, 4), "kernel": "generic (per-block ratio)"}), flush=True)
    lib.resampleFree(ctx)


if __name__ == "__main__" and len(sys.argv) >= 1 and not sys.argv[0].endswith("_probe.py"):
    run("cfg1 mono -1 44.1->48k (64 streams x 2^20)", 1, 1, 44100, 48000, 64, 1 << 20)
    run("cfg2 stereo -3 44.1->48k (64 streams x 2^18)", 2, 3, 44100, 48000, 64, 1 << 18)
    run("cfg2 single stream x 2^22 frames", 2, 3, 44100, 48000, 1, 1 << 22)
    run("cfg2 fixed-ratio init (160 filters, no interp)", 2, 3, 44100, 48000, 64, 1 << 18, fixed=True)
    run("cfg3 64ch -4 96->44.1k lowpass 20k (1 ctx x 2^19)", 64, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000)
    run("cfg3 fixed-ratio init (147 filters, auto lowpass)", 64, 4, 96000, 44100, 1, 1 << 19, fixed=True)
    run("cfg4 1024 stereo streams -3 48->44.1k lowpass (2^15 each)", 2, 3, 48000, 44100, 1024, 1 << 15, lowpass_hz=20000)
    run("cfg4 128 stereo streams -3 48->44.1k lowpass (2^18 each: one GPU's share of 1024, long blocks)", 2, 3, 48000, 44100, 128, 1 << 18, lowpass_hz=20000)
    run("stereo -4 44.1->48k (64 streams x 2^18)", 2, 4, 44100, 48000, 64, 1 << 18)
    run("stereo -2 44.1->48k (64 streams x 2^18)", 2, 2, 44100, 48000, 64, 1 << 18)
    run("cfg2 stereo -3 irrational ratio 1.0884 (generic kernel)", 2, 3, 44100, 44100 * 1.08843537, 64, 1 << 18)
    run("stereo -2 1:1.0001 (near unity, generic)", 2, 2, 48000, 48004.8, 64, 1 << 18)
    run_asrc("cfg5 8ch -2 ASRC +/-100ppm, 256 blocks x 4096 frames", 8, 2, 256, 4096)
    run_asrc("cfg5 8ch -2 ASRC +/-100ppm, 1024 blocks x 480 frames", 8, 2, 1024, 480)
