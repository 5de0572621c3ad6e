"""config 5 alone (256 blocks x 4096 frames, one launch per step) for an ncu capture of the any-ratio kernel; profiles/ aid."""
import sys
sys.argv = [sys.argv[0]]
sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import configs_bench as cb
cb.run_asrc("cfg5 8ch -2 ASRC +/-100ppm, 256 blocks x 4096 frames", 8, 2, 256, 4096, steps=2)
