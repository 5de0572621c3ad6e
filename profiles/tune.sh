#!/bin/bash
# profiles/tune.sh -- A/B runs of bench.py under different kernel-geometry knobs (one gpurun call).
run() { env "$@" ART_B200_TRACE=1 timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu --e2e-streams 1 2> /tmp/err.txt | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$*', '| value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'kernel ms', round(r['kernel_ms_per_launch'], 3), 'share', round(r['kernel_share_of_step'], 3))
"; grep -m1 "\[art\]" /tmp/err.txt; }
for spec in "$@"; do run $spec; done
