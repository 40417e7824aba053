"""The N > 1 host path on CPU: world_size-2 gloo processes shard independent pairs with no overlap and no gap, and
the job time is the max over ranks (what bench.py does with NCCL on GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vslam_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.block_partition(n_units, world, rank)
    owned = torch.zeros(n_units, dtype=torch.int64)
    owned[list(mine)] = 1
    dist.all_reduce(owned)                                    # every unit owned exactly once
    seeds = torch.tensor(list(sharding.weak_seeds(5, rank)))
    gathered = [torch.zeros_like(seeds) for _ in range(world)]
    dist.all_gather(gathered, seeds)
    t = sharding.max_over_ranks(10.0 + rank)
    total = sharding.sum_over_ranks(float(len(mine)))
    dist.barrier()
    if rank == 0:
        out.put((owned.tolist(), torch.cat(gathered).tolist(), t, total))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [4096, 7, 2])
def test_two_rank_block_partition_over_gloo(n_units):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, out)) for r in range(2)]
    for p in procs:
        p.start()
    owned, seeds, t, total = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert owned == [1] * n_units
    assert sorted(seeds) == list(range(10)) and len(set(seeds)) == 10
    assert t == 11.0 and total == n_units


def test_block_partition_matches_floor_rule():
    for n, world in ((4096, 8), (10, 4), (3, 8), (100, 3)):
        for i in range(n):
            owner = i * world // n
            assert i in sharding.block_partition(n, world, owner)
        assert sum(len(sharding.block_partition(n, world, r)) for r in range(world)) == n
    assert [sharding.sequence_owner(s, 8) for s in range(8)] == list(range(8))
