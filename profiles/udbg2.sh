#!/bin/bash
# profiles/udbg2.sh -- kernel time of the tensor-core kernel under role ablations, WITHOUT the role counters
# (ART_B200_UDBG bits: 1 no conversion, 2 no filter copies, 4 no epilogue, 8 no MMAs)
for d in "$@"; do
  ART_B200_UDBG=$d timeout 200 python bench.py --no-cpu --steps 10 --warmup 3 --e2e-streams 1 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('UDBG=$d kernel ms', round(r['kernel_ms_per_launch'], 4), 'cycles/tile', round(r['kernel_ms_per_launch'] * 1.965e6 / 13), 'step ms', round(d['ms_per_step'], 4))"
done
