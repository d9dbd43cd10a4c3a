"""A CPU stand-in for ``lsml_b200.core.engine.BandEngine`` built on the oracle.

TEST INFRASTRUCTURE: lets the HOST logic of ``LevelSetMachineLearning.fit`` / ``segment`` (dataset
split, automatic step, balancing, regressor training order, score bookkeeping, early exit,
sparse reconstruction of the iterates) run on a machine without a GPU, with the oracle's
numerics in place of the kernels, so that it can be compared with the reference's own ``fit``
(tests/golden/fit.npz).  Only the attributes and methods ``lsml_b200/core/model.py`` uses exist.
"""
import numpy as np
import torch

from oracle import oracle as O


class OracleEngine:
    def __init__(self, shape, batch, dx=None, band=3.0, specs=(), device="cpu"):
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.batch = int(batch)
        self.voxels = int(np.prod(self.shape))
        self.total = self.voxels * self.batch
        self.dx = np.ones(self.ndim) if dx is None else np.asarray(dx, dtype=np.float64).copy()
        self.band = float(band)
        self.specs = [tuple(s) for s in specs]
        self.device = torch.device("cpu")
        self.u = torch.zeros((self.batch,) + self.shape, dtype=torch.float64)
        self.mask = torch.zeros((self.batch,) + self.shape, dtype=torch.uint8)
        self._dist = np.zeros((self.batch,) + self.shape)
        self.n_band = 0
        self.band_idx = torch.zeros(0, dtype=torch.int64)
        self.item_offsets = torch.zeros(self.batch + 1, dtype=torch.int64)
        self._last_idx = self.new_u = None

    def load_images(self, imgs, normalize=True):
        imgs = np.asarray(imgs, dtype=np.float64)
        self.img = torch.from_numpy(np.stack([O.normalize_image(i) if normalize else i
                                              for i in imgs]))

    def _compact(self):
        m = self.mask.numpy().astype(bool)
        self.band_idx = torch.from_numpy(np.flatnonzero(m.ravel()))
        counts = m.reshape(self.batch, -1).sum(axis=1)
        self.item_offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
        self.n_band = int(counts.sum())

    def _rebuild(self, b):
        dist, mask = O.distance_transform(self.u[b].numpy(), self.band, self.dx)
        self._dist[b] = np.where(np.isfinite(dist), dist, 0.0)
        self.mask[b] = torch.from_numpy(mask.astype(np.uint8))

    def set_level_sets(self, u0, mask=None):
        t = u0 if torch.is_tensor(u0) else torch.from_numpy(np.ascontiguousarray(u0))
        self.u.copy_(t.reshape(self.u.shape))
        for b in range(self.batch):
            self._rebuild(b)
        self._compact()

    def check_converged(self):
        pass

    def last_record(self, nb):
        return self._last_idx[:nb], self.new_u[:nb]

    def distance(self):
        return torch.from_numpy(self._dist.copy())

    def compute_features(self):
        rows = []
        for b in range(self.batch):
            m = self.mask[b].numpy().astype(bool)
            rows.append(O.feature_matrix(self.specs, self.u[b].numpy(), self.img[b].numpy(),
                                         self._dist[b], m, self.dx))
        return torch.from_numpy(np.concatenate(rows, axis=0))

    def step(self, regressor, step):
        nb = self.n_band
        if nb == 0:
            return 0
        self._last_idx = self.band_idx.clone()
        for b in range(self.batch):
            m = self.mask[b].numpy().astype(bool)
            u, dist, mask, _ = O.evolve_step(self.u[b].numpy(), self.img[b].numpy(), self._dist[b], m,
                                             self.dx, self.specs, regressor, step, self.band,
                                             rebuild=False)
            self.u[b] = torch.from_numpy(u)
        self.new_u = self.u.reshape(-1)[self._last_idx].clone()
        for b in range(self.batch):
            self._rebuild(b)
        self._compact()
        return nb


def ball_u0(eng, centres, radii):
    """numpy stand-in for lsml_b200.core.model._ball_u0 (initializer/provided/ball.py:13-31)."""
    out = np.empty((eng.batch,) + eng.shape)
    sl = (slice(None),) + (None,) * eng.ndim
    for b in range(eng.batch):
        ind = np.indices(eng.shape, dtype=float)
        ind *= eng.dx[sl]
        ind -= np.asarray(centres[b], dtype=float)[sl]
        out[b] = 2.0 * ((radii[b] - np.sqrt((ind ** 2).sum(axis=0))) > 0) - 1.0
    return torch.from_numpy(out)


def overlap_scores(eng, segs_dev, scorer=None):
    """numpy stand-in for lsml_b200.core.model.overlap_scores (score_functions.py:4-30)."""
    seg = segs_dev.numpy().astype(bool)
    fn = O.jaccard if scorer is None or scorer.__name__ == "jaccard" else scorer
    return np.array([fn(eng.u[b].numpy(), seg[b]) for b in range(eng.batch)])
