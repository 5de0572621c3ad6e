#!/bin/bash
# profiles/udbg.sh -- role ablations of the tensor-core kernel (ART_B200_UDBG bits: 1 no conversion, 2 no filter copies,
# 4 no epilogue, 8 no MMAs); results are wrong by construction, only the timing matters.
for d in "$@"; do
  echo "== UDBG=$d"
  ART_B200_UDBG=$d ART_B200_UPROF=1 timeout 200 python bench.py --no-cpu --steps 10 --warmup 3 --e2e-streams 1 2>&1 | grep "umma cycles" | tail -1
done
