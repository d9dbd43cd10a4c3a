"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous item sharding and the
counter reduction bench.py uses (sum of units, max of time).  No data-path collective exists."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions_exactly():
    from lsml_b200.core.sharding import shard_range
    for n in (0, 1, 7, 8, 1024, 16384, 16385):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_items, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lsml_b200.core.sharding import reduce_counters, shard_range
    lo, hi = shard_range(n_items, rank, world)
    units = sum(10 + i for i in range(lo, hi))            # "band voxels" of each item
    total, ms = reduce_counters(units, 5.0 + 3.0 * rank, device=torch.device("cpu"))
    owned = torch.zeros(n_items, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    out[rank] = (total, ms, owned.tolist())
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world, n_items = 2, 9
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29533, n_items, out), nprocs=world, join=True)
        res = dict(out)
    want_total = sum(10 + i for i in range(n_items))
    for rank in range(world):
        total, ms, owned = res[rank]
        assert total == want_total
        assert ms == 8.0                                    # max over ranks
        assert owned == [1] * n_items                       # every item owned exactly once
