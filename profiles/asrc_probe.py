"""config 5 (ASRC, ratio swept +/-100 ppm per block) and a near-unity fixed ratio, device resident; profiles/ aid.
ART_B200_NOUNITY=1 switches the register-blocked consecutive-output form of the any-ratio kernel off."""
import sys
sys.argv = [sys.argv[0]]
sys.path.insert(0, "profiles"); sys.path.insert(0, ".")
import configs_bench as cb
cb.run_asrc("cfg5 8ch -2 ASRC +/-100ppm, 256 blocks x 4096 frames", 8, 2, 256, 4096)
cb.run_asrc("cfg5 8ch -2 ASRC +/-100ppm, 1024 blocks x 480 frames", 8, 2, 1024, 480)
cb.run("stereo -2 1:1.0001 (near unity, generic)", 2, 2, 48000, 48004.8, 64, 1 << 18)
cb.run("cfg2 stereo -3 irrational ratio 1.0884 (generic kernel)", 2, 3, 44100, 44100 * 1.08843537, 64, 1 << 18)
