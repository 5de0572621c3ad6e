"""The biquad cascade alone at BASELINE config 3's size (64 channels x 2^19 frames interleaved, two sections per channel, device
resident) for an ncu capture and a GB/s figure against its 12 B/sample model; profiles/ aid."""
import ctypes as C, json, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg.load()
ch, frames, sections = 64, 1 << 19, 2
coef = pkg.BiquadCoefficients()
lib.biquad_lowpass(C.byref(coef), 0.45 * 44100 / 96000)
sets = [(pkg.Biquad * ch)() for _ in range(sections)]
for st_ in sets:
    for q in st_:
        lib.biquad_init(C.byref(q), C.byref(coef), 1.0)
stages = (C.POINTER(pkg.Biquad) * sections)(*[C.cast(st_, C.POINTER(pkg.Biquad)) for st_ in sets])
x = torch.rand((frames, ch), device="cuda") - 0.5
st = torch.cuda.Stream(); sp = C.c_void_p(st.cuda_stream)
def step():
    lib.biquad_apply_cascade_interleaved_device(stages, sections, ch, C.c_void_p(x.data_ptr()), frames, sp)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record(st)
for _ in range(steps): step()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
samples = frames * ch
print(json.dumps({"probe": "biquad cascade, 64 ch x 2^19 frames, 2 sections", "ms_per_call": round(ms, 4),
                  "Gsamples_per_s": round(samples / ms / 1e6, 2), "model_bytes_per_sample": 12,
                  "GBps_model": round(samples * 12 / ms / 1e6, 1), "hbm_frac": round(samples * 12 / ms / 1e6 / 6547.2, 4)}))
