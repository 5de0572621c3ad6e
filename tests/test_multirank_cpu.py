"""CPU test of the N>1 path (world_size 2, gloo): streams are sharded across ranks with no data-path
collective; the only collectives are the MAX of the elapsed time and the SUM of the work (bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = bench.shard_streams(1024, world, rank)              # BASELINE config 4: 1024 stereo streams
    # every rank "processes" its own streams: pretend rank r needs (r + 1) ms and yields 100 frames per stream
    ms, units = bench.reduce_over_ranks(dist, torch.device("cpu"), float(rank + 1), 100.0 * count)
    out[rank] = (first, count, ms, units)
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_reduction_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][:2] == (0, 512) and res[1][:2] == (512, 512)
    for r in range(world):
        assert res[r][2] == 2.0              # MAX over ranks of the elapsed time
        assert res[r][3] == 100.0 * 1024     # SUM over ranks of the work


@pytest.mark.parametrize("total,world", [(1024, 8), (10, 4), (3, 8), (64, 1)])
def test_shards_partition_the_streams(total, world):
    spans = [bench.shard_streams(total, world, r) for r in range(world)]
    assert spans[0][0] == 0 and sum(c for _, c in spans) == total
    for (f0, c0), (f1, _) in zip(spans, spans[1:]):
        assert f0 + c0 == f1
    assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
