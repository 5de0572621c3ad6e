"""Host-path probe: batched host API over n contexts carved from ONE large pinned allocation."""
import ctypes as C, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.load()
T, F, CH, R = 380, 380, 2, 48000 / 44100
frames = 262144
cap = int(frames * R) + 400
NMAX = 16
hx = torch.empty((NMAX, frames, CH)).uniform_(-0.5, 0.5).pin_memory(); hy = torch.empty((NMAX, cap, CH)).pin_memory()
f32p = C.POINTER(C.c_float)
ctx_t = C.POINTER(pkg.Resample)
for n in (1, 2, 4, 8, 16):
    ctxs = [lib.resampleInit(CH, T, F, 0.0, 3) for _ in range(n)]
    for c in ctxs: lib.resampleAdvancePosition(c, T / 2)
    ca = (ctx_t * n)(*ctxs)
    ia = (f32p * n)(*[C.cast(hx[i].data_ptr(), f32p) for i in range(n)])
    oa = (f32p * n)(*[C.cast(hy[i].data_ptr(), f32p) for i in range(n)])
    ni, no = (C.c_int * n)(*([frames] * n)), (C.c_int * n)(*([cap] * n))
    ra = (C.c_double * n)(*([R] * n)); res = (pkg.ResampleResult * n)()
    def batch(): lib.resampleBatchProcessInterleaved(ca, n, ia, ni, oa, no, ra, res)
    for _ in range(3): batch()
    t0 = time.perf_counter()
    for _ in range(10): batch()
    dt = (time.perf_counter() - t0) / 10
    made = sum(r.output_generated for r in res) * CH
    print(f"n={n:2d}: {dt * 1e6:8.1f} us per batch call, {dt / n * 1e6:7.1f} us per context, {made / dt / 1e9:6.2f} Gsamples/s")
    def single(i): return lib.resampleProcessInterleaved(ctxs[i], ia[i], frames, oa[i], cap, R)
    for i in range(n): single(i)
    t0 = time.perf_counter()
    for _ in range(10):
        for i in range(n): single(i)
    dt = (time.perf_counter() - t0) / 10
    print(f"      sequential single calls: {dt / n * 1e6:7.1f} us per context")
    for c in ctxs: lib.resampleFree(c)
