"""Host ceiling of the end-to-end path at N GPUs: every rank moves the bench workload's bytes (2 MiB up, 2.28 MiB down per
stereo stream, pinned host memory, both directions at once on two streams) for ~1 s, ALL RANKS AT THE SAME TIME.  No kernel
runs: what comes out is what the host's memory system and the PCIe links give N concurrent duplex copies -- the number the
library's end-to-end throughput at N GPUs has to be judged against.

  python profiles/pcie_ngpu_probe.py                                  (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/pcie_ngpu_probe.py

Prints one JSON line on rank 0: per-rank and aggregate GB/s in each direction and the implied ceiling in output Gsamples/s
(44.1 -> 48 kHz: 4 bytes down and 3.675 bytes up per output sample).
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = torch.device("cuda", local)
nin, nout = 262144 * 2, 285350 * 2              # floats per stream: one bench stream up, its output down
K = 64                                          # streams per step, as bench.py's e2e leg
hin = torch.empty((K, nin), dtype=torch.float32).pin_memory()
hout = torch.empty((K, nout), dtype=torch.float32).pin_memory()
hin.uniform_(-0.5, 0.5)
gin = torch.empty((K, nin), dtype=torch.float32, device=d)
gout = torch.zeros((K, nout), dtype=torch.float32, device=d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(K):
            if up:
                with torch.cuda.stream(s1):
                    gin[i].copy_(hin[i], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout[i].copy_(gout[i], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=d, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / reps            # seconds per step of K streams, slowest rank


run(True, True, 2)
reps = 120                                      # ~0.4 s per one-way test, ~0.4 s duplex: sustained, not a burst
t_up, t_down, t_both = run(True, False, reps), run(False, True, reps), run(True, True, reps)
up_b, down_b = K * nin * 4, K * nout * 4
if rank == 0:
    print(json.dumps({
        "probe": "pcie_ngpu", "n_gpus": world, "host_cpus": os.cpu_count(),
        "bytes_per_rank_per_step": {"h2d": up_b, "d2h": down_b},
        "h2d_alone_GBps_per_rank": up_b / t_up / 1e9, "d2h_alone_GBps_per_rank": down_b / t_down / 1e9,
        "duplex_GBps_per_rank": (up_b + down_b) / t_both / 1e9,
        "duplex_GBps_aggregate": world * (up_b + down_b) / t_both / 1e9,
        "e2e_ceiling_Gsamples_per_s": world * K * nout / t_both / 1e9,
        "note": "ceiling = output samples all ranks could move per second if the resampling itself were free"}), flush=True)
if world > 1:
    dist.destroy_process_group()
