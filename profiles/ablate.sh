#!/bin/bash
# build an ablation copy of the library and time the bench launch under a few skip masks
set -e
cd /root/repo
make -C audio-resampler_b200 clean >/dev/null
make -C audio-resampler_b200 -j8 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -DART_B200_ABLATE" >/dev/null 2>&1
for m in 0 4 32 1 8; do
  ART_B200_UDBG=$m python bench.py --steps 1 --launches-per-step 30 --no-configs --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('udbg $m kernel_ms', round(d['roofline']['kernel_ms_per_launch'],4))" || echo "udbg $m failed"
done
