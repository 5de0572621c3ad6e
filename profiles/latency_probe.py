"""Per-call latency probe: one stereo preset -3 context, 262144-frame calls; device-pointer vs host-pointer API."""
import ctypes as C, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.load()
T, F, CH, R = 380, 380, 2, 48000 / 44100
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cap = int(frames * R) + 400
dev = torch.device("cuda")
x = torch.rand((frames, CH), device=dev) - 0.5
y = torch.empty((cap, CH), device=dev)
hx = torch.empty((frames, CH)).uniform_(-0.5, 0.5).pin_memory(); hy = torch.empty((cap, CH)).pin_memory()
ctx = lib.resampleInit(CH, T, F, 0.0, 3); lib.resampleAdvancePosition(ctx, T / 2)
st = torch.cuda.Stream(); sp = C.c_void_p(st.cuda_stream)
f32p = C.POINTER(C.c_float)
def dev_call():
    lib.resampleProcessInterleavedDevice(ctx, C.c_void_p(x.data_ptr()), frames, C.c_void_p(y.data_ptr()), cap, R, sp)
def host_call():
    lib.resampleProcessInterleaved(ctx, C.cast(hx.data_ptr(), f32p), frames, C.cast(hy.data_ptr(), f32p), cap, R)
for name, fn, sync in (("device-pointer call + sync", dev_call, True), ("device-pointer call, enqueue only", dev_call, False), ("host-pointer call", host_call, False)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
        if sync: st.synchronize()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:36s}: {(t1 - t0) / 50 * 1e6:8.1f} us/call host-side, {(t2 - t0) / 50 * 1e6:8.1f} us/call incl. drain")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(50): dev_call()
e1.record(st); torch.cuda.synchronize()
print(f"GPU time per device call: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")

# A/B: the same upload -> resample -> download sequence with torch issuing the copies on the work stream
def torch_copies():
    with torch.cuda.stream(st):
        x.copy_(hx, non_blocking=True)
        dev_call()
        hy.copy_(y, non_blocking=True)
    st.synchronize()
for _ in range(5): torch_copies()
t0 = time.perf_counter()
for _ in range(50): torch_copies()
print(f"torch copies + device-pointer call + sync : {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us/call")
# and the library's host path fed with pageable memory for comparison
import numpy as np
px = np.random.default_rng(0).uniform(-0.5, 0.5, (frames, CH)).astype(np.float32); py = np.empty((cap, CH), np.float32)
def pageable():
    lib.resampleProcessInterleaved(ctx, px.ctypes.data_as(f32p), frames, py.ctypes.data_as(f32p), cap, R)
for _ in range(5): pageable()
t0 = time.perf_counter()
for _ in range(50): pageable()
print(f"host-pointer call, pageable numpy buffers  : {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us/call")
