#!/bin/bash
# role counters of the tensor-core kernel: build a measurement copy of the float library (-DART_B200_ABLATE compiles the per-role
# cycle counters and the work-skipping switches in) on the GPU box and run a few launches of the metric workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
rm -rf audio-resampler_b200/build
make -C audio-resampler_b200 -j16 audio-resampler_b200/lib/libresampler_b200.so NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -DART_B200_ABLATE" > /dev/null 2>&1 || make -C audio-resampler_b200 -j16 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -DART_B200_ABLATE" > /dev/null 2>&1
ART_B200_UPROF=1 ${CMD:-python bench.py --steps 2 --warmup 1 --launches-per-step 8 --no-configs --no-cpu --no-e2e} 2>&1 | grep "\[art\]"
