"""The FFMA rational-ratio kernel alone on the metric config (tensor path off) for an ncu capture; profiles/ aid."""
import sys
sys.argv = [sys.argv[0]]
sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import configs_bench as cb
cb.lib.resampleB200SetTensorPath(0)
cb.run("cfg2 stereo -3 44.1->48k (64 streams x 2^18), FFMA form", 2, 3, 44100, 48000, 64, 1 << 18, steps=3)
cb.run("cfg2 fixed-ratio init (160 filters, no interp), FFMA form", 2, 3, 44100, 48000, 64, 1 << 18, fixed=True, steps=3)
