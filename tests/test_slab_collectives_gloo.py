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


def test_slab_layout_halo_exchange_is_consistent():
    """Pure index bookkeeping of the slab decomposition (no GPU): what a rank sends is exactly
    what its neighbour expects, in volume coordinates; the Dirichlet planes sit just outside the
    receiver's evaluated range; a simulated exchange fills every halo with the owner's planes."""
    from lsml_b200.core.slab import HALO_STATE, LocalCollective, SlabLayout
    for n0, world, gauss, overlap in [(40, 3, 6, 8), (37, 4, 6, 8), (24, 2, 0, 8), (1024, 8, 8, 8),
                                      (128, 4, 12, 16), (64, 2, 0, 0), (30, 1, 8, 8), (96, 3, 0, 3)]:
        lays = [SlabLayout(n0, r, world, gauss, overlap) for r in range(world)]
        assert lays[0].A == 0 and lays[-1].B == n0
        for r, L in enumerate(lays):
            assert 0 <= L.elo <= L.lo < L.hi <= L.ehi <= L.nloc
            assert L.origin0 + L.lo == L.A and L.origin0 + L.hi == L.B
            if world > 1:
                assert L.hw >= L.W + HALO_STATE + 1 and L.hw >= gauss
            if r > 0:
                assert L.lo - L.elo == L.W and L.elo >= HALO_STATE      # room for Dirichlet planes
            if r + 1 < world:
                assert L.ehi - L.hi == L.W and L.nloc - L.ehi >= HALO_STATE
        for wide in (False, True):
            for r in range(world - 1):
                lo_rank, hi_rank = lays[r], lays[r + 1]
                send_up = lo_rank.send_slices(wide)[1]          # r -> r+1
                recv_lo = hi_rank.recv_slices(wide)[0]
                send_dn = hi_rank.send_slices(wide)[0]          # r+1 -> r
                recv_hi = lo_rank.recv_slices(wide)[1]
                g = lambda L, s: (L.origin0 + s.start, L.origin0 + s.stop)
                assert g(lo_rank, send_up) == g(hi_rank, recv_lo), (n0, world, wide, r)
                assert g(hi_rank, send_dn) == g(lo_rank, recv_hi), (n0, world, wide, r)
                # the sender owns what it sends
                assert lo_rank.lo <= send_up.start and send_up.stop <= lo_rank.hi
                assert hi_rank.lo <= send_dn.start and send_dn.stop <= hi_rank.hi
                if not wide:
                    assert recv_lo.stop == hi_rank.elo and recv_hi.start == lo_rank.ehi
        # simulated exchange of a volume whose plane p holds the value p
        coll = LocalCollective(world)
        vol = torch.arange(n0, dtype=torch.float64)[:, None].repeat(1, 3)
        local = []
        for L in lays:
            a = vol[L.origin0:L.origin0 + L.nloc].clone()
            a[:L.lo] = -1.0
            a[L.hi:] = -1.0                                     # halos unknown
            local.append(a)
        sends = [[tuple(None if s is None else a[s] for s in L.send_slices(True))]
                 for L, a in zip(lays, local)]
        recvs = coll.exchange(sends)
        for L, a, rc in zip(lays, local, recvs):
            for src, sl in zip(rc[0], L.recv_slices(True)):
                if sl is not None:
                    a[sl] = src
            assert torch.equal(a, vol[L.origin0:L.origin0 + L.nloc]), (n0, world)


def test_slab_volume_exchange_plumbing_with_stub_engine(monkeypatch):
    """SlabVolume._exchange / SlabRank._halo_sends / _halo_store on CPU tensors: the engine (which
    needs a GPU) is replaced by a stub that only owns the arrays, so that the exchange code the
    GPU tests and torchrun jobs run is also exercised in the CPU tier."""
    from lsml_b200.core import slab

    class StubEngine:
        def __init__(self, shape, batch, dx=None, band=3.0, specs=(), device="cpu"):
            self.shape, self.device = tuple(shape), torch.device("cpu")
            self.total = int(np.prod(shape))
            self.u = torch.zeros((1,) + self.shape, dtype=torch.float64)
            self.mask = torch.zeros((1,) + self.shape, dtype=torch.uint8)

    monkeypatch.setattr(slab, "BandEngine", StubEngine)
    shape, world = (48, 5, 6), 3
    vol = slab.SlabVolume(shape, world=world, specs=[("image_sample", 2.0)], device="cpu")
    full_u = torch.arange(48, dtype=torch.float64)[:, None, None].expand(shape).contiguous()
    full_m = (torch.arange(48) % 7).to(torch.uint8)[:, None, None].expand(shape).contiguous()
    for r in vol.ranks:
        sl = r.local_slice()
        r.eng.u[0].copy_(full_u[sl])
        r.eng.mask[0].copy_(full_m[sl])
        r.eng.u[0, :r.lo] = -1
        r.eng.u[0, r.hi:] = -1                      # halos unknown
    flags = vol._exchange(lambda r: r.eng.u[0], wide=True)
    for r, f in zip(vol.ranks, flags):
        assert torch.equal(r.eng.u[0], full_u[r.local_slice()])
        assert f.tolist() == [1 if r.rank > 0 else 0, 1 if r.rank + 1 < world else 0]
    # a second exchange changes nothing
    assert all(f.tolist() == [0, 0] for f in vol._exchange(lambda r: r.eng.u[0], wide=True))
    # Dirichlet planes of two arrays in one round: they land just outside the evaluated range
    for r in vol.ranks:
        r.eng.u[0, :r.elo] = -5
        r.eng.u[0, r.ehi:] = -5
        r.eng.mask[0, :r.elo] = 99
        r.eng.mask[0, r.ehi:] = 99
    vol._exchange(lambda r: r.eng.u[0], lambda r: r.eng.mask[0])
    for r in vol.ranks:
        g = r.local_slice()
        if r.rank > 0:
            assert torch.equal(r.eng.u[0, r.elo - 2:r.elo], full_u[g][r.elo - 2:r.elo])
            assert torch.equal(r.eng.mask[0, r.elo - 2:r.elo], full_m[g][r.elo - 2:r.elo])
            if r.elo > 2:
                assert (r.eng.u[0, :r.elo - 2] == -5).all()       # nothing else was touched
        if r.rank + 1 < world:
            assert torch.equal(r.eng.u[0, r.ehi:r.ehi + 2], full_u[g][r.ehi:r.ehi + 2])
            assert torch.equal(r.eng.mask[0, r.ehi:r.ehi + 2], full_m[g][r.ehi:r.ehi + 2])
