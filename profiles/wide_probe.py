"""Device-resident throughput of the PATH_WIDTH=64 library (double samples, any-ratio kernel with double multiply-adds); profiles/ aid."""
import ctypes as C, json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import __graft_entry__ as entry

pkg = entry.load_package(); lib = pkg.load64()
st = torch.cuda.Stream(); sp = C.c_void_p(st.cuda_stream)
PRESET = {1: (48, 48), 2: (320, 156), 3: (380, 380), 4: (988, 988)}


def run(name, ch, preset, src, dst, streams, frames, steps=5):
    filters, taps = PRESET[preset]
    ratio = dst / src
    ctxs = [lib.resampleInit(ch, taps, filters, 0.0, 3) for _ in range(streams)]
    for c in ctxs:
        lib.resampleAdvancePosition(c, taps / 2)
    x = torch.rand((streams, frames, ch), device="cuda", dtype=torch.float64) - 0.5
    cap = int(frames * ratio) + taps + 16
    y = torch.empty((streams, cap, ch), device="cuda", dtype=torch.float64)
    n = streams
    ca = (C.POINTER(pkg.Resample64) * n)(*ctxs)
    ia = (C.c_void_p * n)(*[x[i].data_ptr() for i in range(n)]); oa = (C.c_void_p * n)(*[y[i].data_ptr() for i in range(n)])
    ni, no = (C.c_int * n)(*([frames] * n)), (C.c_int * n)(*([cap] * n))
    ra = (C.c_double * n)(*([ratio] * n)); res = (pkg.ResampleResult * n)()

    def step():
        lib.resampleBatchProcessInterleavedDevice(ca, n, ia, ni, oa, no, ra, res, sp)
        return sum(r.output_generated for r in res)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); made = 0
    for _ in range(steps):
        made += step()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sps = made * ch / (ms * 1e-3)
    print(json.dumps({"config": name, "Gsamples_per_s": round(sps / 1e9, 2), "ms_per_step": round(ms / steps, 3),
                      "hbm_frac": round(sps * 8.0 * (1.0 + 1.0 / ratio) / 6553e9, 4), "dfma_per_sample": 2 * taps}), flush=True)
    for c in ctxs:
        lib.resampleFree(c)


run("PATH_WIDTH=64 stereo preset -3 44.1->48k (64 streams x 2^16)", 2, 3, 44100, 48000, 64, 1 << 16)
run("PATH_WIDTH=64 stereo preset -1 44.1->48k (64 streams x 2^18)", 2, 1, 44100, 48000, 64, 1 << 18)
run("PATH_WIDTH=64 8 ch preset -2 ratio 1.0001 (16 streams x 2^16)", 8, 2, 48000, 48004.8, 16, 1 << 16)
