#!/bin/bash
# Host ceiling and end-to-end throughput at N GPUs (all ranks at once): the PCIe duplex probe, then bench.py's e2e leg under
# the three host wait modes.  usage: profiles/e2e_scale_probe.sh N
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
[ "$N" = 1 ] && RUN=python
nproc
$RUN profiles/pcie_ngpu_probe.py 2>/dev/null | grep pcie_ngpu
for mode in spin block yield; do
  ART_B200_WAIT=$mode $RUN bench.py --gpus $N --steps 1 --launches-per-step 8 --no-configs --no-cpu --e2e-steps 128 2>/dev/null | \
    python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$mode N=$N e2e', round(d['e2e']['value']), 'batched', round(d['e2e']['batched_value']), 'device', round(d['value']))"
done
