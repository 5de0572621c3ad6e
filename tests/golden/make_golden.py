"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libartref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only -- the GPU box has no
/root/reference; the committed .npz files are what travels.

    python tests/golden/make_golden.py

Each case records the call sequence (seeded inputs are regenerated from the seed, not stored),
the reference's per-call (input_used, output_generated, position) and its concatenated output.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import artlibs as A  # noqa: E402

BH_INTERP = A.SUBSAMPLE_INTERPOLATE | A.BLACKMAN_HARRIS

# name, channels, (filters, taps), init kind, ratio, flags, call plan
CASES = [
    # BASELINE config 1: mono, preset -1, 44.1k -> 48k, artest-style 4096-frame calls
    dict(name="cfg1_mono_p1_441_48", ch=1, preset=1, ratio=48000 / 44100, flags=BH_INTERP, lowpass=0.0,
         calls=[4096] * 4, flush=True, seed=11),
    # BASELINE config 2 (metric config): stereo, preset -3, 44.1k -> 48k
    dict(name="cfg2_stereo_p3_441_48", ch=2, preset=3, ratio=48000 / 44100, flags=BH_INTERP, lowpass=0.0,
         calls=[4096, 1000, 1, 0, 7000], flush=True, seed=12),
    # BASELINE config 3 shape (fewer channels): preset -4, 96k -> 44.1k with 20 kHz lowpass
    dict(name="cfg3_4ch_p4_96_441_lp", ch=4, preset=4, ratio=44100 / 96000, flags=BH_INTERP,
         lowpass=20000 * 2.0 / 96000, calls=[6000, 3000], flush=True, seed=13),
    # BASELINE config 4: one stream of it, 48k -> 44.1k preset -3 with lowpass
    dict(name="cfg4_stereo_p3_48_441_lp", ch=2, preset=3, ratio=44100 / 48000, flags=BH_INTERP,
         lowpass=20000 * 2.0 / 48000, calls=[5000, 5000], flush=False, seed=14),
    # fixed-ratio init as art.c uses it: 44.1k -> 48k reduces to 160 filters, no interpolation, snap
    dict(name="fixed_stereo_p3_441_48", ch=2, preset=3, fixed=(44100, 48000, 0), flags=BH_INTERP | A.INCLUDE_LOWPASS,
         calls=[4096, 333, 5000], flush=True, seed=15),
    # fixed-ratio downsample with automatic lowpass (147 filters)
    dict(name="fixed_stereo_p3_48_441_autolp", ch=2, preset=3, fixed=(48000, 44100, 0),
         flags=BH_INTERP | A.INCLUDE_LOWPASS, calls=[4800, 4800], flush=True, seed=16),
    # 2x upsampling without lowpass: every other output is an input sample verbatim (resampler.c:1141)
    dict(name="fixed_mono_p1_x2_passthrough", ch=1, preset=1, fixed=(24000, 48000, 0), flags=BH_INTERP,
         calls=[3000], flush=True, seed=17),
    # Hann window + double-precision convolution
    dict(name="hann_precise_stereo_p2", ch=2, preset=2, ratio=0.731, flags=A.SUBSAMPLE_INTERPOLATE | A.EXTEND_CONVOLUTION_MATH,
         lowpass=0.7, calls=[5000, 2500], flush=True, seed=18),
]


def run_case(case, make_stream):
    filters, taps = A.PRESETS[case["preset"]]
    if "fixed" in case:
        s = make_stream(case["ch"], taps, filters, flags=case["flags"], fixed=case["fixed"])
        ratio = 0.0
        eff = case["fixed"][1] / case["fixed"][0]
    else:
        s = make_stream(case["ch"], taps, filters, case["lowpass"], flags=case["flags"])
        ratio = eff = case["ratio"]
    s.advance(taps / 2)
    rng = np.random.default_rng(case["seed"])
    outs, meta = [], []
    for i, n in enumerate(case["calls"]):
        x = rng.uniform(-0.5, 0.5, (n, case["ch"])).astype(np.float32)
        last = i == len(case["calls"]) - 1
        cap = int(n * eff) + taps + 16
        y, used, made = s.process(x, cap, ratio, flush_after=last and case["flush"])
        outs.append(y)
        meta.append((used, made, s.position()))
    s.close()
    return np.concatenate(outs, axis=0), np.array(meta, dtype=np.float64)


def main():
    ref = A.reference()
    assert ref is not None, "oracle/_ref/libartref.so missing: run `make -C oracle ref` in the build container"
    for case in CASES:
        out, meta = run_case(case, A.reference_stream)
        np.savez_compressed(HERE / f"{case['name']}.npz", out=out, meta=meta)
        print(f"{case['name']}: {out.shape[0]} frames x {out.shape[1]} ch, peak {np.abs(out).max():.4f}")
    # known-answer bank structure from SURVEY.md 8c, straight from the reference build
    s = A.reference_stream(1, 48, 48, 0.0)
    bank = s.bank()
    np.savez_compressed(HERE / "bank_48x48_bh.npz", bank=bank)
    coeffs = A.BiquadCoefficients()
    ref.biquad_lowpass(coeffs, 0.45 * 44100 / 96000)
    np.savez_compressed(HERE / "biquad_lowpass_0p2067.npz",
                        coeffs=np.array([coeffs.a0, coeffs.a1, coeffs.a2, coeffs.b1, coeffs.b2], np.float32))
    # biquad cascade response of the reference (art.c:1011-1017 usage): 3 channels, 2 stages
    rng = np.random.default_rng(21)
    x = rng.uniform(-0.5, 0.5, (5000, 3)).astype(np.float32)
    y = x.copy()
    import ctypes as C
    stages = [[A.Biquad() for _ in range(3)] for _ in range(2)]
    for st in stages:
        for q in st:
            ref.biquad_init(C.byref(q), C.byref(coeffs), 1.0)
    for lo, hi in [(0, 1234), (1234, 5000)]:
        for c in range(3):
            for st in stages:
                ptr = y[lo:, c:].ctypes.data_as(A.f32p)
                ref.biquad_apply_buffer(C.byref(st[c]), ptr, hi - lo, 3)
    np.savez_compressed(HERE / "biquad_cascade_3ch.npz", out=y)
    print("golden vectors written")


if __name__ == "__main__":
    main()
