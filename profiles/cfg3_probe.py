"""BASELINE config 3 alone (64 ch, preset -4, 96k->44.1k + lowpass) with kernel timing; profiles/ aid."""
import sys
sys.argv = [sys.argv[0]]
import ctypes as C
sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import configs_bench as cb
cb.lib.resampleB200ProfileEnable(1)
cb.run("cfg3 64ch -4 96->44.1k lowpass 20k (1 ctx x 2^19)", 64, 4, 96000, 44100, 1, 1 << 19, lowpass_hz=20000, steps=5)
ms = C.c_double(0.0)
n = cb.lib.resampleB200ProfileCollect(C.byref(ms))
print("product kernel launches", n, "avg ms", ms.value / max(1, n))
