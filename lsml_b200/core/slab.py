"""Slab decomposition of ONE large 3-D volume over several GPUs.

The volume is cut along axis 0 (the slowest axis, so halo faces are contiguous planes).  Rank r
owns the planes ``[A_r, B_r)`` and keeps a local grid ``[A_r - h, B_r + h)`` whose extra planes
are *halo copies* of the neighbours' voxels.  Per evolution iteration (lsml/core/model.py:378-398):

====================  =========================================================================
features              exact integer ``{u>0}`` statistics and band sums are all-reduced (SUM)
                      before the per-item feature values are formed; everything else is local
velocity, update      local (the upwind stencil reads one halo plane of ``u``)
halo exchange of u    the two planes next to every cut, neighbour to neighbour
band rebuild          the single-GPU sweeps (csrc/fmm.cu) restricted to the owned planes:
                      ``begin -> { march -> exchange halo planes of T and state -> requeue }``
                      until no rank has queued work, then ``finish``.  The sweeps converge to the
                      unique fixed point of the marcher's replay operator, so the band mask and
                      the distances are IDENTICAL to the single-GPU result.
compaction            of the owned planes; global band order = rank order x local order
====================  =========================================================================

Overlap (alternating Schwarz): a rebuild needs one exchange round every time a dependency chain
of the marcher crosses a cut, and those chains run ALONG the interface.  Each rank therefore also
evaluates ``overlap`` ghost planes beyond its cuts; the Dirichlet data it receives are the two
planes just outside that evaluated range, taken from the neighbour's solution.  On the overlap
both ranks solve the same strip problem with the same boundary data, whose fixed point is unique,
so at convergence they agree there and the owned parts glue to the global solution -- but a chain
now has to leave the overlap before it costs another round.

Communication goes through a small ``Collective`` interface with two implementations:
``TorchCollective`` (one process per GPU, ``torch.distributed``: NCCL over NVLink on a B200 box,
gloo in CPU-side tests of the bookkeeping) and ``LocalCollective`` (all ranks emulated in ONE
process on one device, neighbour copies instead of messages) which lets the single-GPU test
tier exercise exactly the same phase code.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .engine import BandEngine, gaussian_weights
from .regressors import export_regressor

import os

HALO_STATE = 2          # planes of T / state exchanged per cut (replay stencil reach)
# default ghost planes evaluated redundantly beyond every cut (LSML_SLAB_OVERLAP overrides: probes)
OVERLAP = int(os.environ.get("LSML_SLAB_OVERLAP", "8"))


def plane_ranges(n0, world):
    """Contiguous, near-equal plane ranges ``[(A_0, B_0), ...]`` of axis 0."""
    return [(r * n0 // world, (r + 1) * n0 // world) for r in range(world)]


class SlabLayout:
    """Index bookkeeping of one rank's local grid (pure integers; tested on the CPU).

    Local plane q of rank r is plane ``origin0 + q`` of the volume.  ``[lo, hi)`` are the owned
    planes, ``[elo, ehi)`` the planes the band rebuild evaluates (owned + ``W`` ghost planes
    beyond each cut), ``hw`` the halo width kept on each side of an interior cut."""

    def __init__(self, n0, rank, world, halo_needed=0, overlap=8):
        self.n0, self.rank, self.world = int(n0), int(rank), int(world)
        ranges = plane_ranges(self.n0, self.world)
        self.A, self.B = ranges[rank]
        thick = min(b - a for a, b in ranges)
        if world > 1 and thick < HALO_STATE:
            raise ValueError("every slab needs at least %d planes" % HALO_STATE)
        self.W = max(0, min(int(overlap), thick - HALO_STATE - 1)) if world > 1 else 0
        # ghost planes + Dirichlet planes + one more plane of u, or the widest axis-0 Gaussian
        self.hw = max(self.W + HALO_STATE + 1, int(halo_needed)) if world > 1 else 0
        if world > 1 and thick < self.hw:
            raise ValueError("slabs of %d planes are thinner than the %d-plane halo they have to "
                             "supply to their neighbours: use fewer ranks" % (thick, self.hw))
        self.hlo = min(self.hw, self.A)
        self.hhi = min(self.hw, self.n0 - self.B)
        self.lo, self.hi = self.hlo, self.hlo + (self.B - self.A)
        self.origin0 = self.A - self.hlo
        self.nloc = self.hi + self.hhi
        self.elo = self.lo - self.W if rank > 0 else 0
        self.ehi = self.hi + self.W if rank + 1 < world else self.nloc

    def send_slices(self, wide):
        """(slice sent to the lower neighbour, slice sent to the upper neighbour), local planes.
        ``wide``: the whole halo width (u); otherwise the two Dirichlet planes just outside the
        neighbour's evaluated range, which lie ``W`` planes inside this rank's owned range."""
        n, off = (self.hw, 0) if wide else (HALO_STATE, self.W)
        lo = slice(self.lo + off, self.lo + off + n) if self.rank > 0 else None
        hi = slice(self.hi - off - n, self.hi - off) if self.rank + 1 < self.world else None
        return lo, hi

    def recv_slices(self, wide):
        """(slice filled from the lower neighbour, slice filled from the upper neighbour)."""
        if wide:
            lo = slice(0, self.lo) if self.rank > 0 else None
            hi = slice(self.hi, self.nloc) if self.rank + 1 < self.world else None
        else:
            lo = slice(self.elo - HALO_STATE, self.elo) if self.rank > 0 else None
            hi = slice(self.ehi, self.ehi + HALO_STATE) if self.rank + 1 < self.world else None
        return lo, hi


class LocalCollective:
    """All ranks live in this process: ``parts`` arguments are lists with one entry per rank."""

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def allreduce_sum(self, parts):
        total = parts[0].clone()
        for p in parts[1:]:
            total += p.to(total.device)
        return [total.to(p.device) for p in parts]

    def exchange(self, sends):
        """sends[r] = list over arrays of (to_lower, to_upper) tensors (None at the ends of the
        volume); returns recvs[r] = list over arrays of (from_lower, from_upper)."""
        out = []
        for r in range(self.world):
            per_array = []
            for a in range(len(sends[r])):
                lo = sends[r - 1][a][1] if r > 0 else None
                hi = sends[r + 1][a][0] if r + 1 < self.world else None
                per_array.append((None if lo is None else lo.clone(),
                                  None if hi is None else hi.clone()))
            out.append(per_array)
        return out


class TorchCollective:
    """One rank per process; lists have exactly one entry."""

    def __init__(self, rank=None, world=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.local_ranks = [self.rank]

    def allreduce_sum(self, parts):
        t = parts[0].clone()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [t]

    def exchange(self, sends):
        """All arrays' halo planes in ONE batch of point-to-point operations."""
        ops, out = [], []
        for to_lo, to_hi in sends[0]:
            from_lo = from_hi = None
            if self.rank > 0:
                from_lo = torch.empty_like(to_lo)
                ops.append(self.dist.P2POp(self.dist.isend, to_lo.contiguous(), self.rank - 1))
                ops.append(self.dist.P2POp(self.dist.irecv, from_lo, self.rank - 1))
            if self.rank + 1 < self.world:
                from_hi = torch.empty_like(to_hi)
                ops.append(self.dist.P2POp(self.dist.isend, to_hi.contiguous(), self.rank + 1))
                ops.append(self.dist.P2POp(self.dist.irecv, from_hi, self.rank + 1))
            out.append((from_lo, from_hi))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        return [out]


class SlabRank:
    """One rank's share of the volume: a BandEngine over the local (owned + halo) grid."""

    def __init__(self, shape, rank, world, dx, band, specs, device, overlap=OVERLAP):
        if len(shape) != 3:
            raise ValueError("slab decomposition is for 3-D volumes")
        self.gshape = tuple(int(s) for s in shape)
        self.rank, self.world = rank, world
        dx = np.ones(3) if dx is None else np.asarray(dx, dtype=np.float64)
        # halo wide enough for the axis-0 Gaussians of the image planes (one-time) and for the
        # ghost + Dirichlet planes of the band rebuild (every iteration)
        gauss = 0
        for spec in specs:
            if spec[0] == "com_ray":
                raise NotImplementedError(
                    "COMRaySample reads the image along a ray through the whole volume and is "
                    "not available on slab-decomposed volumes")
            if spec[0] in ("boundary_size", "isoperimetric"):
                raise NotImplementedError(
                    "BoundarySize / IsoperimetricRatio are not available on slab-decomposed "
                    "volumes (the marching-cubes area of the cut planes would need a halo pass)")
            if spec[0] == "moments" and any(o > 2 for o in spec[2]):
                raise NotImplementedError(
                    "Moments of order > 2 are not available on slab-decomposed volumes")
            if spec[0] in ("smooth_u", "user_plane"):
                raise NotImplementedError(
                    "feature spec %r is not available on slab-decomposed volumes (it needs a "
                    "Gaussian halo of u every iteration / the whole image on one host)" % (spec[0],))
            if spec[0] in ("image_sample", "image_edge", "band_mean", "band_std") and spec[1] > 0:
                gauss = max(gauss, gaussian_weights(float(spec[1]) / dx[0], 0)[1])
        self.layout = lay = SlabLayout(self.gshape[0], rank, world, gauss, overlap)
        self.A, self.B, self.W, self.hw = lay.A, lay.B, lay.W, lay.hw
        self.hlo, self.hhi, self.lo, self.hi = lay.hlo, lay.hhi, lay.lo, lay.hi
        self.origin0, self.elo, self.ehi = lay.origin0, lay.elo, lay.ehi
        lshape = (lay.nloc,) + self.gshape[1:]
        self.eng = BandEngine(lshape, 1, dx=dx, band=band, specs=specs, device=device)
        self.plane = self.gshape[1] * self.gshape[2]
        self.device = self.eng.device
        self.pending = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._band_is_rebuilt = False

    # -------------------------------------------------------------------------------- views
    def local_slice(self):
        """Planes of the volume held locally (owned + halo)."""
        return slice(self.origin0, self.origin0 + self.eng.shape[0])

    def T_view(self):
        e = self.eng
        return e._ws_fmm[:e.total * 8].view(torch.float64).view(e.shape)

    def _halo_sends(self, arr, wide=False):
        """(planes for the lower neighbour, planes for the upper neighbour) of a local array."""
        lo, hi = self.layout.send_slices(wide)
        return (None if lo is None else arr[lo], None if hi is None else arr[hi])

    def _halo_store(self, arr, recv, wide=False):
        """Write received planes into the halo; returns an int64 tensor [2] = (lower halo
        changed, upper halo changed)."""
        changed = torch.zeros(2, dtype=torch.int64, device=self.device)
        bits = torch.uint8 if arr.dtype == torch.uint8 else torch.int64      # bitwise compare
        for side, src, sl in zip((0, 1), recv, self.layout.recv_slices(wide)):
            if src is None or sl is None:
                continue
            dst = arr[sl]
            src = src.to(self.device)
            changed[side] = (dst.view(bits) != src.view(bits)).any().to(torch.int64)
            dst.copy_(src)
        return changed

    # -------------------------------------------------------------------------------- stats
    def shift_stats(self, stats, offset):
        """Exact integer statistics re-expressed for axis-0 indices shifted by ``offset``:
        sum (i+o) = sum i + o n,  sum (i+o)^2 = sum i^2 + 2 o sum i + o^2 n."""
        s = stats.clone()
        n, s1 = stats[0, 0], stats[0, 1]
        s[0, 1] = s1 + offset * n
        s[0, 4] = stats[0, 4] + 2 * offset * s1 + offset * offset * n
        return s

    def count_owned(self):
        """Dense recount of the owned planes -> eng.stats in LOCAL plane coordinates."""
        e = self.eng
        owned = e.u[0, self.lo:self.hi]
        shape_c = _lib.int_array(owned.shape)
        e.stats.zero_()
        _lib.check(_lib.lib.lsml_global_stats(
            ctypes.c_int(3), shape_c, _lib.c_i64(1), _lib.ptr(owned), _lib.ptr(e.stats),
            e._stream()), "global_stats")
        e.stats.copy_(self.shift_stats(e.stats, self.lo))

    def global_frame_stats(self):
        return self.shift_stats(self.eng.stats, self.origin0)

    # -------------------------------------------------------------------------------- rebuild
    def _fmm_args(self):
        e = self.eng
        return (ctypes.c_int(3), e._shape_c, e._dx_c, ctypes.c_int(self.elo), ctypes.c_int(self.ehi))

    def rebuild_begin(self, from_band):
        e = self.eng
        if e._ws_fmm is None:
            nbytes = int(_lib.lib.lsml_fmm_workspace_bytes(_lib.c_i64(e.total)))
            e._ws_fmm = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            e.mask.zero_()
            _lib.check(_lib.lib.lsml_fmm_workspace_init(
                _lib.c_i64(e.total), _lib.ptr(e._ws_fmm), e._stream()), "fmm_init")
        from_band = from_band and self._band_is_rebuilt
        # Dirichlet copies of the previous rebuild are stale: far until the first exchange says
        # otherwise (the ghost planes inside the evaluated range are cleared by the kernels)
        T, m = self.T_view(), e.mask[0]
        if self.elo:
            T[:self.elo].fill_(float("inf")); m[:self.elo].zero_()
        if self.ehi < e.shape[0]:
            T[self.ehi:].fill_(float("inf")); m[self.ehi:].zero_()
        # u moved on the owned band (searched from the band) and, through the exchange, anywhere
        # in the ghost planes and next to them (searched densely)
        dense = _lib.int_array((self.elo, self.lo + 1 if self.rank > 0 else self.elo,
                                self.hi - 1 if self.rank + 1 < self.world else self.ehi, self.ehi))
        _lib.check(_lib.lib.lsml_fmm_slab_begin(
            *self._fmm_args(), _lib.ptr(e.u), _lib.ptr(e.mask), _lib.ptr(e._ws_fmm),
            _lib.ptr(e.band_idx if from_band else None),
            _lib.ptr(e.n_band_dev if from_band else None), dense, e._stream()), "fmm_slab_begin")

    def rebuild_march(self):
        e = self.eng
        _lib.check(_lib.lib.lsml_fmm_slab_march(
            *self._fmm_args(), _lib.ptr(e.u), ctypes.c_double(e.band), _lib.ptr(e.mask),
            _lib.ptr(e._ws_fmm), e._stream()), "fmm_slab_march")

    def rebuild_requeue(self, sides):
        """sides: device int64[2], non-zero = the lower / upper halo changed."""
        e = self.eng
        _lib.check(_lib.lib.lsml_fmm_slab_requeue(
            *self._fmm_args(), ctypes.c_double(e.band), _lib.ptr(sides),
            _lib.ptr(e.mask), _lib.ptr(e._ws_fmm), _lib.ptr(self.pending), e._stream()),
            "fmm_slab_requeue")

    def rebuild_finish(self):
        e = self.eng
        _lib.check(_lib.lib.lsml_fmm_slab_finish(
            *self._fmm_args(), ctypes.c_double(e.band), _lib.ptr(e.mask), _lib.c_p(0),
            _lib.ptr(e._ws_fmm), e._stream()), "fmm_slab_finish")
        self._band_is_rebuilt = True

    def clear_band(self):
        """No zero contour anywhere in the volume: empty band (distance_transform.py:42-47)."""
        self.eng.mask.zero_()
        self._band_is_rebuilt = False

    def compact_owned(self):
        e = self.eng
        e.band_idx, e._band_idx_prev = e._band_idx_prev, e.band_idx
        if e.band_idx.numel() < e.total:
            e.band_idx = torch.empty(e.total, dtype=torch.int64, device=self.device)
        owned = e.mask[0, self.lo:self.hi]
        _lib.check(_lib.lib.lsml_compact_mask_range(
            _lib.c_i64(owned.numel()), _lib.ptr(owned), _lib.ptr(e.band_idx),
            _lib.ptr(e.n_band_dev), _lib.c_i64(self.lo * self.plane), _lib.ptr(e._ws_compact),
            e._stream()), "compact_mask_range")
        e.n_band = int(e.n_band_dev.item())
        e.item_offsets[0] = 0
        e.item_offsets[1] = e.n_band
        e._ensure_band_buffers()

    # -------------------------------------------------------------------------------- features
    def band_sums(self, mode, mean):
        """Raw sums of every image plane over the owned band (mode 0: x, 1: (x - mean)^2)."""
        e = self.eng
        npl = max(len(e.program.planes), 1)
        out = torch.zeros((1, npl), dtype=torch.float64, device=self.device)
        if e.n_band and e.program.needs_band_stats:
            _lib.check(_lib.lib.lsml_band_sums(
                _lib.c_i64(e.voxels), _lib.c_i64(1), _lib.ptr(e.planes), _lib.c_i64(e.total),
                ctypes.c_int(len(e.program.planes)), _lib.ptr(e.band_idx),
                _lib.ptr(e.item_offsets), ctypes.c_int(mode), _lib.ptr(mean), _lib.ptr(out),
                _lib.ptr(e._ws_band), e._stream()), "band_sums")
        return out

    def assemble(self, stats_global):
        e, pr = self.eng, self.eng.program
        npl = len(pr.planes)
        _lib.check(_lib.lib.lsml_item_features(
            ctypes.c_int(pr.nfeat), pr.kind, pr.a, pr.b, ctypes.c_int(3), e._shape_c,
            _lib.c_i64(1), e._dx_c, ctypes.c_int(max(npl, 1)), _lib.ptr(stats_global),
            _lib.ptr(e.band_mean), _lib.ptr(e.band_std), _lib.ptr(e.boundary),
            _lib.ptr(e.item_feat), _lib.ptr(e.com), e._stream()), "item_features")

    def feature_matrix(self):
        """The materialised (n_band, F) matrix of the owned band (tests / training only; the
        evolution step evaluates the columns inside the velocity kernel)."""
        e, pr = self.eng, self.eng.program
        e._ensure_X(e.n_band)
        if e.n_band:
            _lib.check(_lib.lib.lsml_assemble_features_origin(
                ctypes.c_int(pr.nfeat), pr.kind, pr.a, pr.b, ctypes.c_int(3), e._shape_c,
                _lib.c_i64(1), e._dx_c, _lib.ptr(e.band_idx), _lib.ptr(e.n_band_dev),
                _lib.ptr(e.planes), _lib.c_i64(e.total), _lib.ptr(e.item_feat), _lib.ptr(e.com),
                _lib.ptr(e.X), ctypes.c_int(self.origin0), e._stream()), "assemble_features")
        return e.X[:e.n_band]

    def predict_and_update(self, regressor, step):
        e = self.eng
        if not e.n_band:
            return
        reg = export_regressor(regressor, self.device)
        reg.predict_band(e.band_source(origin0=self.origin0), e.n_band_dev, e.vel)
        _lib.check(_lib.lib.lsml_band_update_stats(
            ctypes.c_int(3), e._shape_c, _lib.c_i64(1), e._dx_c, _lib.ptr(e.u), _lib.ptr(e.vel),
            _lib.ptr(e.band_idx), _lib.ptr(e.n_band_dev), ctypes.c_double(float(step)),
            _lib.ptr(e.new_u), _lib.c_p(0), _lib.ptr(e.stats), e._stream()), "band_update")


class SlabVolume:
    """A single volume evolved by ``world`` ranks.  ``collective=None`` emulates all ranks in this
    process on ``device`` (LocalCollective); pass a TorchCollective under torchrun."""

    def __init__(self, shape, world=None, dx=None, band=3.0, specs=(), device="cuda",
                 collective=None, overlap=OVERLAP):
        self.coll = collective if collective is not None else LocalCollective(int(world))
        self.world = self.coll.world
        self.shape = tuple(int(s) for s in shape)
        self.specs = [tuple(s) for s in specs]
        self.band = float(band)
        self.ranks = [SlabRank(self.shape, r, self.world, dx, band, self.specs, device, overlap)
                      for r in self.coll.local_ranks]
        self.n_band = 0
        self.rebuild_rounds = 0

    # ------------------------------------------------------------------ setup
    def load_image(self, img, normalize=True):
        """``img``: the whole volume (numpy / tensor, every rank passes the same array or at
        least its own ``local_slice()`` planes filled in).  Normalisation uses the mean and
        population std of the WHOLE image (model.py:328-329), formed from all-reduced sums."""
        img_t = img if torch.is_tensor(img) else torch.from_numpy(np.ascontiguousarray(img))
        self.load_image_parts([img_t[r.local_slice()] for r in self.ranks], normalize)

    def load_image_parts(self, parts, normalize=True):
        """Same, from the planes each local rank holds (``SlabRank.local_slice()``), so that no
        process ever has to materialise the whole volume."""
        parts = [(p if torch.is_tensor(p) else torch.from_numpy(np.ascontiguousarray(p))).to(
            r.device, dtype=torch.float64) for p, r in zip(parts, self.ranks)]
        if normalize:
            n = float(np.prod(self.shape))
            own = [p[r.lo:r.hi] for p, r in zip(parts, self.ranks)]
            sums = self.coll.allreduce_sum([o.sum().reshape(1) for o in own])
            means = [s / n for s in sums]
            sq = self.coll.allreduce_sum([((o - m) ** 2).sum().reshape(1) for o, m in zip(own, means)])
            parts = [(p - m) / torch.sqrt(s / n) for p, m, s in zip(parts, means, sq)]
        for r, p in zip(self.ranks, parts):
            r.eng.load_images(p[None], normalize=False)

    def set_level_sets(self, u0):
        u_t = u0 if torch.is_tensor(u0) else torch.from_numpy(np.ascontiguousarray(u0))
        self.set_level_sets_parts([u_t[r.local_slice()] for r in self.ranks])

    def set_level_sets_parts(self, parts):
        for r, p in zip(self.ranks, parts):
            e = r.eng
            p = p if torch.is_tensor(p) else torch.from_numpy(np.ascontiguousarray(p))
            e.u[0].copy_(p.to(r.device, dtype=torch.float64))
            r.count_owned()
            r._band_is_rebuilt = False
        self._rebuild(from_band=False)

    # ------------------------------------------------------------------ collectives
    def _global_stats(self):
        return self.coll.allreduce_sum([r.global_frame_stats() for r in self.ranks])

    def _exchange(self, *getters, wide=False):
        """Halo exchange of one or more local arrays per rank in a single round of messages;
        returns, per rank, the sum over arrays of the (lower, upper) 'changed' flags."""
        arrs = [[g(r) for g in getters] for r in self.ranks]
        recvs = self.coll.exchange([[r._halo_sends(a, wide) for a in al]
                                    for r, al in zip(self.ranks, arrs)])
        out = []
        for r, al, rc in zip(self.ranks, arrs, recvs):
            flags = [r._halo_store(a, c, wide) for a, c in zip(al, rc)]
            out.append(sum(flags[1:], flags[0]))
        return out

    def _rebuild(self, from_band):
        gstats = self._global_stats()
        s = gstats[0][0].cpu().numpy().astype(np.int64)
        total = int(np.prod(self.shape))
        has_contour = int(s[7]) == total or not (int(s[0]) == 0 or int(s[0]) == total)
        if not has_contour:
            for r in self.ranks:
                r.clear_band()
                r.compact_owned()
            self.n_band = 0
            return
        for r in self.ranks:
            r.rebuild_begin(from_band)
        rounds = 0
        while True:
            for r in self.ranks:
                r.rebuild_march()
            changed = self._exchange(lambda r: r.T_view(), lambda r: r.eng.mask[0])
            for r, c in zip(self.ranks, changed):
                r.rebuild_requeue(c)
            rounds += 1                               # (one host read per round: the line below)
            if int(self.coll.allreduce_sum([r.pending for r in self.ranks])[0].item()) == 0:
                break
            if rounds > 10000:
                raise RuntimeError("slab band rebuild did not converge")
        for r in self.ranks:
            r.rebuild_finish()
            r.compact_owned()
        self.rebuild_rounds = rounds
        self.n_band = int(self.coll.allreduce_sum(
            [torch.tensor([r.eng.n_band], dtype=torch.int64, device=r.device)
             for r in self.ranks])[0].item())

    # ------------------------------------------------------------------ one iteration
    def step(self, regressor, step):
        nb = self.n_band
        if nb == 0:                       # `if mask.any()` (model.py:381)
            return 0
        gstats = self._global_stats()
        pr = self.ranks[0].eng.program
        if pr.needs_band_stats:
            cnt = float(nb)
            sums = self.coll.allreduce_sum([r.band_sums(0, None) for r in self.ranks])
            for r, sm in zip(self.ranks, sums):
                r.eng.band_mean.copy_(sm / cnt)
            sq = self.coll.allreduce_sum([r.band_sums(1, r.eng.band_mean) for r in self.ranks])
            for r, sm in zip(self.ranks, sq):
                r.eng.band_std.copy_(torch.sqrt(sm / cnt))
        for r, gs in zip(self.ranks, gstats):
            r.assemble(gs.to(r.device))
            r.predict_and_update(regressor, step)
        self._exchange(lambda r: r.eng.u[0], wide=True)
        self._rebuild(from_band=True)
        return nb

    # ------------------------------------------------------------------ results
    def gather(self, what="u"):
        """The owned planes of every LOCAL rank as one array (the whole volume when emulated)."""
        outs = []
        for r in self.ranks:
            a = r.eng.u[0] if what == "u" else r.eng.mask[0]
            outs.append(a[r.lo:r.hi].cpu())
        return torch.cat(outs).numpy()
