#!/bin/bash
# build a measurement copy of the float library (-DART_B200_ABLATE) on the GPU box and time the metric launch under a few
# work-skipping masks (results are wrong by construction; the point is what the kernel would cost if a role were free):
# 1 converters request no samples, 2 no filter copies, 4 epilogue reads no accumulators and stores nothing, 8 no MMAs, 32 no output stores,
# 64 converters skip the proxy fence, 128 converters skip the digit split and its stores (MASKS="..." overrides the list)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
rm -rf audio-resampler_b200/build
make -C audio-resampler_b200 -j16 NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v -DART_B200_ABLATE" > /dev/null 2>&1
for m in ${MASKS:-0 1 2 4 8 32 12 13}; do
  ART_B200_UDBG=$m python bench.py --steps 1 --launches-per-step 30 --no-configs --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('udbg $m kernel_ms', round(d['roofline']['kernel_ms_per_launch'],4))" || echo "udbg $m failed"
done
