"""CPU (gloo, world_size 3) test of the slab decomposition's communication layer: the
TorchCollective a torchrun job uses must deliver exactly what the in-process LocalCollective
(used by the single-GPU test tier) delivers -- neighbour halo planes for several arrays in one
round of messages, and SUM all-reduces of the integer statistics."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_plane_ranges_cover_the_axis():
    from lsml_b200.core.slab import plane_ranges
    for n0 in (8, 37, 128, 1024):
        for world in (1, 2, 3, 8):
            rr = plane_ranges(n0, world)
            assert rr[0][0] == 0 and rr[-1][1] == n0
            assert all(b == c for (_, b), (c, _) in zip(rr, rr[1:]))
            sizes = [b - a for a, b in rr]
            assert max(sizes) - min(sizes) <= 1


def _sends(rank, world):
    """Two arrays (float64 'T', uint8 'state'), two planes each way, values encode the sender."""
    t = torch.full((2, 3, 4), float(rank), dtype=torch.float64)
    m = torch.full((2, 3, 4), rank, dtype=torch.uint8)
    lo = (t + 0.25, m + 10) if rank > 0 else (None, None)
    hi = (t + 0.5, m + 20) if rank + 1 < world else (None, None)
    return [(lo[0], hi[0]), (lo[1], hi[1])]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lsml_b200.core.slab import TorchCollective
    coll = TorchCollective()
    recv = coll.exchange([_sends(rank, world)])[0]
    stats = torch.arange(8, dtype=torch.int64).reshape(1, 8) * (rank + 1)
    total = coll.allreduce_sum([stats])[0]
    out[rank] = ([[None if x is None else x.numpy().copy() for x in pair] for pair in recv],
                 total.numpy().copy())
    dist.destroy_process_group()


def test_torch_collective_matches_local_collective():
    from lsml_b200.core.slab import LocalCollective
    world = 3
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, 29541, out), nprocs=world, join=True)
        res = dict(out)
    local = LocalCollective(world)
    want = local.exchange([_sends(r, world) for r in range(world)])
    want_total = local.allreduce_sum([torch.arange(8, dtype=torch.int64).reshape(1, 8) * (r + 1)
                                      for r in range(world)])[0].numpy()
    for r in range(world):
        got, total = res[r]
        assert np.array_equal(total, want_total)
        for a in range(2):
            for side in range(2):
                w = want[r][a][side]
                if w is None:
                    assert got[a][side] is None
                else:
                    assert np.array_equal(got[a][side], w.numpy()), (r, a, side)
    # rank 1 received rank 0's "to upper" planes from below and rank 2's "to lower" from above
    assert float(res[1][0][0][0][0, 0, 0]) == 0.5 and float(res[1][0][0][1][0, 0, 0]) == 2.25
