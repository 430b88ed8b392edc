"""Host logic of the multi-GPU path under gloo (world_size 2 and 3, CPU): shard ranges and the slab gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cvgpuspeedup_b200 import sharding


def test_shard_ranges_partition_everything():
    for n in [0, 1, 7, 8, 50, 8192, 8191]:
        for world in [1, 2, 3, 4, 8]:
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.full((n, 3, 4, 5), float("nan"))
        lo, hi = sharding.shard_range(n, rank, world)
        # stand-in for the kernel: plane z gets the value z (written in place in this rank's slab only)
        for z in range(lo, hi):
            full[z] = float(z)
        sharding.gather_slabs(full, n)
        want = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 3, 4, 5)
        assert torch.equal(full, want), f"rank {rank}: gathered tensor is wrong"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8), (2, 7), (3, 10)])
def test_gather_slabs_gloo(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n), nprocs=world, join=True)
