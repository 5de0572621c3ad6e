"""One BASELINE config alone with the tensor kernel's role counters (ART_B200_UPROF=1); profiles/ aid.
usage: python profiles/cfg_probe.py cfg1|cfg2|cfg3|cfg4"""
import sys
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
sys.argv = [sys.argv[0]]
sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import configs_bench as cb
if which == "cfg1":
    cb.run("cfg1 mono -1 44.1->48k (64 streams x 2^20)", 1, 1, 44100, 48000, 64, 1 << 20, steps=5)
elif which == "cfg2":
    cb.run("cfg2 stereo -3 44.1->48k (64 streams x 2^18)", 2, 3, 44100, 48000, 64, 1 << 18, steps=5)
elif which == "cfg4":
    cb.run("cfg4 1024 stereo streams -3 48->44.1k lowpass (2^15 each)", 2, 3, 48000, 44100, 1024, 1 << 15, lowpass_hz=20000, steps=5)
else:
    cb.run("cfg3 64ch -4 96->44.1k lowpass 20k (1 ctx x 2^19)", 64, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000, steps=5)
