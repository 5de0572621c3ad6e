#!/bin/bash
# end-to-end leg of bench.py under the three host wait modes (ART_B200_WAIT) and two run lengths; 1 GPU
for mode in spin block yield; do
  for steps in 8 256; do
    ART_B200_WAIT=$mode python bench.py --steps 1 --launches-per-step 8 --no-configs --no-cpu --e2e-steps $steps 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', $steps, round(d['e2e']['value']), round(d['e2e']['batched_value']))"
  done
done
