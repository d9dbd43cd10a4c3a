"""Device-resident level-set evolution engine: the hot path of lsml on one B200.

One :class:`BandEngine` holds a *batch* of independent items (images or volumes of identical
shape) entirely in HBM and advances all of them one iteration per :meth:`step`:

    features = FeatureMap(u, img, dist, mask, dx)          lsml/feature/feature_map.py:53-80
    v[mask]  = regressor.predict(features[mask])           lsml/core/model.py:388
    gmag     = gmag_os(u, v, mask, dx)                     lsml/core/model.py:390
    u[mask] += step * v[mask] * gmag[mask]                 lsml/core/model.py:394
    dist, mask = distance_transform(u, band, dx)           lsml/core/model.py:397

Data layout in HBM (all float64, C order):
    u        (batch, *shape)            level sets, updated in place on the band
    planes   (P, batch, *shape)         u-independent image planes, computed ONCE per image:
                                        smoothed image / edge magnitude per sigma
    mask     (batch, *shape) uint8      band mask (= the reference's ``mask``)
    band_idx (B,) int64                 ascending flat indices item*voxels+local (numpy order)
    X        (B, F)                     feature matrix, row order = band order
    vel, new_u (B,)                     velocity and updated level-set value per band voxel

All numerical work is done by the CUDA kernels behind ``liblsml_b200.so``; this class only
owns buffers (torch tensors) and sequences launches on the current stream.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .regressors import export_regressor

FK_PLANE, FK_DIST_COM, FK_BAND_MEAN, FK_BAND_STD = 0, 1, 2, 3
FK_SIZE, FK_BOUNDARY, FK_ISOPERIMETRIC, FK_MOMENT = 4, 5, 6, 7
FK_COM_RAY = 8


def gaussian_kernel1d(sigma, order, radius):
    """scipy.ndimage._filters._gaussian_kernel1d restated (same numpy operations, so the
    weights are bit-identical to the ones scipy's gaussian_filter1d uses)."""
    if order < 0:
        raise ValueError("order must be non-negative")
    exponent_range = np.arange(order + 1)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    if order == 0:
        return phi_x
    q = np.zeros(order + 1)
    q[0] = 1
    D = np.diag(exponent_range[1:], 1)
    P = np.diag(np.ones(order) / -sigma2, -1)
    Q_deriv = D + P
    for _ in range(order):
        q = Q_deriv.dot(q)
    q = (x[:, None] ** exponent_range).dot(q)
    return q * phi_x


def gaussian_weights(sigma, order, truncate=4.0):
    """(reversed weights, radius) exactly as scipy.ndimage.gaussian_filter1d builds them."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    return np.ascontiguousarray(gaussian_kernel1d(sd, order, lw)[::-1]), lw


class FeatureProgram:
    """Feature specs (see oracle/oracle.py for the tuple vocabulary) compiled to the column
    program of ``lsml_assemble_features`` plus the list of image planes it samples."""

    def __init__(self, specs, ndim):
        self.specs = [tuple(s) for s in specs]
        self.ndim = ndim
        self.planes = []          # ("sample" | "edge", sigma)
        kind, a, b = [], [], []
        self.needs_boundary = False
        self.hi_moments = []      # (column, axis, order) of the Moments columns of order > 2
        self.needs_band_stats = False
        self.needs_u_planes = False
        self.user_planes = {}     # plane index -> callable(img, dx)
        for spec in self.specs:
            k = spec[0]
            if k == "image_sample":
                kind.append(FK_PLANE); a.append(self._plane("sample", spec[1])); b.append(0)
            elif k == "image_edge":
                kind.append(FK_PLANE); a.append(self._plane("edge", spec[1])); b.append(0)
            elif k == "band_mean":
                kind.append(FK_BAND_MEAN); a.append(self._plane("sample", spec[1])); b.append(0)
                self.needs_band_stats = True
            elif k == "band_std":
                kind.append(FK_BAND_STD); a.append(self._plane("sample", spec[1])); b.append(0)
                self.needs_band_stats = True
            elif k == "size":
                kind.append(FK_SIZE); a.append(0); b.append(0)
            elif k == "boundary_size":
                kind.append(FK_BOUNDARY); a.append(0); b.append(0)
                self.needs_boundary = True
            elif k == "isoperimetric":
                kind.append(FK_ISOPERIMETRIC); a.append(0); b.append(0)
                self.needs_boundary = True
            elif k == "moments":
                for axis in spec[1]:
                    for order in spec[2]:
                        if int(order) != order or not 1 <= order <= 64:
                            raise ValueError("Moments order should be an integer in 1..64")
                        if order > 2:       # exact per-axis histograms + lsml_axis_moments
                            self.hi_moments.append((len(kind), int(axis), int(order)))
                        kind.append(FK_MOMENT); a.append(int(axis)); b.append(int(order))
            elif k == "dist_to_com":
                kind.append(FK_DIST_COM); a.append(0); b.append(0)
            elif k == "smooth_u":
                # gaussian_filter(u, sigma) sampled on the band (examples/gestalt2d/
                # prev_iter_feature.py:21-22): a plane recomputed from u every iteration
                kind.append(FK_PLANE); a.append(self._plane("usmooth", spec[1])); b.append(0)
                self.needs_u_planes = True
            elif k == "user_plane":
                # ("user_plane", key, fn): fn(img, dx) -> dense float64 array, evaluated ONCE per
                # image on the host by the user's own code (u-independent image features such as
                # examples/gestalt2d/corner_feature.py), then sampled like any image plane
                if len(spec) != 3 or not callable(spec[2]):
                    raise ValueError("user_plane spec must be ('user_plane', key, callable)")
                idx = self._plane("user", str(spec[1]))
                self.user_planes[idx] = spec[2]
                kind.append(FK_PLANE); a.append(idx); b.append(0)
            elif k == "com_ray":
                # COMRaySample: 2*n_samples columns reading the RAW image (image.py:147-171)
                n_samples = int(spec[1])
                if ndim not in (2, 3):
                    raise ValueError("`ndim` must be either 2 or 3")
                if not 1 <= n_samples < 32768:
                    raise ValueError("COMRaySample: n_samples must be in 1..32767")
                plane = self._plane("sample", 0.0)
                for j in range(2 * n_samples):
                    kind.append(FK_COM_RAY); a.append(plane); b.append(j | (n_samples << 16))
            else:
                raise NotImplementedError(
                    "no device kernel registered for feature spec %r (there is no CPU "
                    "fallback)" % (spec,))
        if self.needs_boundary and ndim not in (2, 3):
            raise ValueError("Boundary size is only defined for dimensions 2 and 3; "
                             "ndim provided = {}".format(ndim))
        self.nfeat = len(kind)
        if self.nfeat > 64:
            raise ValueError("at most 64 feature columns")
        self.kind = _lib.int_array(kind)
        self.a = _lib.int_array(a)
        self.b = _lib.int_array(b)

    def _plane(self, what, sigma):
        key = (what, sigma if what == "user" else float(sigma))
        if key not in self.planes:
            self.planes.append(key)
        return self.planes.index(key)


class _Phase:
    def __init__(self, eng, name):
        self.eng, self.name = eng, name

    def __enter__(self):
        if self.eng.timers is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.eng.timers is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.eng.timers.setdefault(self.name, []).append((self.a, b))
        return False


class BandEngine:
    def __init__(self, shape, batch, dx=None, band=3.0, specs=(), device="cuda"):
        _lib.require_cuda()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("BandEngine runs on CUDA devices only")
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        assert 1 <= self.ndim <= 3, "Only dimensions 1-3 supported."
        self.batch = int(batch)
        self.voxels = int(np.prod(self.shape))
        self.total = self.voxels * self.batch
        self.dx = (np.ones(self.ndim) if dx is None else np.asarray(dx, dtype=np.float64)).copy()
        if self.dx.shape != (self.ndim,):
            raise ValueError("`dx` has incorrect number of elements.")
        self.band = float(band)
        self.program = FeatureProgram(specs, self.ndim)
        self._shape_c = _lib.int_array(self.shape)
        self._dx_c = _lib.double_array(self.dx)
        d, f64, i64 = self.device, torch.float64, torch.int64
        full = (self.batch,) + self.shape
        with torch.cuda.device(d):
            self.u = torch.zeros(full, dtype=f64, device=d)
            self.mask = torch.zeros(full, dtype=torch.uint8, device=d)
            self.planes = torch.zeros((max(len(self.program.planes), 1),) + full, dtype=f64, device=d)
            self.stats = torch.zeros((self.batch, 8), dtype=torch.int64, device=d)
            # band size / indices: two buffers each, swapped on every compaction, so that the
            # band an iteration ran on (its sparse record, and the start set of the next rebuild)
            # stays valid without a copy
            self.n_band_dev = torch.zeros(1, dtype=i64, device=d)
            self._n_band_prev_dev = torch.zeros(1, dtype=i64, device=d)
            self.units_dev = torch.zeros(1, dtype=i64, device=d)   # sum of band sizes stepped
            self.item_offsets = torch.zeros(self.batch + 1, dtype=i64, device=d)
            self.band_idx = torch.zeros(1, dtype=i64, device=d)
            self._band_idx_prev = torch.zeros(1, dtype=i64, device=d)
            npl = max(len(self.program.planes), 1)
            self.band_mean = torch.zeros((self.batch, npl), dtype=f64, device=d)
            self.band_std = torch.zeros((self.batch, npl), dtype=f64, device=d)
            self.boundary = torch.zeros(self.batch, dtype=f64, device=d)
            self.item_feat = torch.zeros((self.batch, max(self.program.nfeat, 1)), dtype=f64, device=d)
            self.com = torch.zeros((self.batch, 3), dtype=f64, device=d)
            self._ws_compact = torch.empty(
                int(_lib.lib.lsml_compact_workspace_bytes(_lib.c_i64(self.total))),
                dtype=torch.uint8, device=d)
            _lib.lib.lsml_band_stats_workspace_bytes.restype = _lib.c_i64
            self._ws_band = torch.empty(max(8, int(_lib.lib.lsml_band_stats_workspace_bytes(
                _lib.c_i64(self.voxels), _lib.c_i64(self.batch), ctypes.c_int(npl)))),
                dtype=torch.uint8, device=d)
            if self.program.needs_boundary and self.ndim == 2:
                _lib.lib.lsml_contour_workspace_bytes.restype = _lib.c_i64
                self._ws_contour = torch.empty(int(_lib.lib.lsml_contour_workspace_bytes(
                    ctypes.c_int(self.shape[0]), ctypes.c_int(self.shape[1]),
                    _lib.c_i64(self.batch))), dtype=torch.uint8, device=d)
            elif self.program.needs_boundary:
                _lib.lib.lsml_surface_area_workspace_bytes.restype = _lib.c_i64
                self._ws_contour = torch.empty(max(8, int(_lib.lib.lsml_surface_area_workspace_bytes(
                    ctypes.c_int(3), self._shape_c, _lib.c_i64(self.batch)))),
                    dtype=torch.uint8, device=d)
            if self.program.hi_moments:
                _lib.lib.lsml_axis_moments_workspace_bytes.restype = _lib.c_i64
                self._ws_moments = torch.empty(int(_lib.lib.lsml_axis_moments_workspace_bytes(
                    ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch))),
                    dtype=torch.uint8, device=d)
                hm = np.asarray(self.program.hi_moments, dtype=np.int32)
                self._hm_cols, self._hm_axes, self._hm_orders = (
                    torch.from_numpy(np.ascontiguousarray(hm[:, q])).to(d) for q in range(3))
            _lib.lib.lsml_fmm_workspace_bytes.restype = _lib.c_i64
            self._ws_fmm = None       # allocated on the first band rebuild
        self._weights = {}            # (sigma, order) -> (device weights, radius)
        self._tmp_plane = None
        self._n_band_host = 0         # host copy of n_band_dev, None = not read back yet
        self.X = self.vel = self.new_u = None
        self._src = None              # lsml_band_source handed to the fused velocity kernels
        self.history = None           # list of (band_idx, new_u) per iteration when recording
        self.last_fmm_counters = None
        self._band_is_rebuilt = False
        self.timers = None            # dict while bench.py collects per-phase device times

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return _lib.current_stream()

    def _phase(self, name):
        """CUDA-event bracket around one phase of an iteration, active only while
        ``self.timers`` is a dict (bench.py's per-kernel evidence pass); a no-op otherwise."""
        return _Phase(self, name)

    def phase_times_ms(self):
        """{phase: (launch groups, total ms)} of the brackets recorded so far (synchronises)."""
        torch.cuda.synchronize(self.device)
        return {k: (len(v), float(sum(a.elapsed_time(b) for a, b in v)))
                for k, v in (self.timers or {}).items()}

    # The band size lives on the device (n_band_dev); every kernel of an iteration reads it
    # there, so `step` never waits for the host.  The host copy is fetched on demand.
    @property
    def n_band(self):
        if self._n_band_host is None:
            self._n_band_host = int(self.n_band_dev.item())
        return self._n_band_host

    @n_band.setter
    def n_band(self, value):
        self._n_band_host = None if value is None else int(value)

    @property
    def band_voxel_iterations(self):
        """Sum of the band sizes of all iterations stepped so far (one host read)."""
        return int(self.units_dev.item())

    def _ensure_band_buffers(self):
        """Per-band-voxel velocity and updated value: sized for the largest possible band (the
        whole grid: band == 0, or u == 0 everywhere), so that no iteration has to know its band
        size on the host."""
        if self.vel is None:
            self.vel = torch.zeros(self.total, dtype=torch.float64, device=self.device)
            self.new_u = torch.empty(self.total, dtype=torch.float64, device=self.device)

    def _ensure_X(self, n):
        """The materialised feature matrix (FeatureMap.__call__, fit): only on request."""
        n = max(int(n), 1)
        if self.X is None or self.X.shape[0] < n:
            cap = min(self.total, max(int(n * 1.25) + 1024, 4096))
            self.X = torch.empty((max(cap, n), max(self.program.nfeat, 1)), dtype=torch.float64,
                                 device=self.device)

    # ------------------------------------------------------------------ images
    def load_images(self, imgs, normalize=True):
        """imgs: (batch, *shape) float64 numpy array or CUDA tensor.  Computes every image plane
        the feature program samples (one-time, u-independent)."""
        with torch.cuda.device(self.device):
            t = imgs if torch.is_tensor(imgs) else torch.from_numpy(np.ascontiguousarray(imgs))
            t = t.to(self.device, dtype=torch.float64, non_blocking=True).contiguous()
            if tuple(t.shape) != (self.batch,) + self.shape:
                raise ValueError("`imgs` has shape {} but must be shape {}".format(
                    tuple(t.shape), (self.batch,) + self.shape))
            if normalize:
                out = torch.empty_like(t)
                _lib.lib.lsml_normalize_workspace_bytes.restype = _lib.c_i64
                ws = torch.empty(int(_lib.lib.lsml_normalize_workspace_bytes(
                    _lib.c_i64(self.voxels), _lib.c_i64(self.batch))), dtype=torch.uint8,
                    device=self.device)
                _lib.check(_lib.lib.lsml_normalize_f64(
                    _lib.c_i64(self.voxels), _lib.c_i64(self.batch), _lib.ptr(t), _lib.ptr(out),
                    _lib.ptr(ws), self._stream()), "normalize")
                t = out
            self.img = t
            self._compute_planes()

    def _gauss(self, src, dst, sigma, order, axis, mode):
        key = (float(sigma), int(order))
        if key not in self._weights:              # uploaded once per (sigma, order)
            w, radius = gaussian_weights(sigma, order)
            self._weights[key] = (torch.from_numpy(w).to(self.device), radius)
        wd, radius = self._weights[key]
        _lib.check(_lib.lib.lsml_gauss1d_f64(
            ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), _lib.ptr(src),
            _lib.ptr(dst), _lib.ptr(wd), ctypes.c_int(radius), ctypes.c_int(1 if order else 0),
            ctypes.c_int(axis), ctypes.c_int(mode), self._stream()), "gauss1d")

    def _smooth_chain(self, src, dst, sigmas):
        """dst = gaussian_filter1d applied along every axis in turn (scipy's order and
        arithmetic); axes with sigma <= 1e-15 are skipped as scipy.ndimage.gaussian_filter does."""
        axes = [i for i in range(self.ndim) if sigmas[i] > 1e-15]
        if not axes:
            dst.copy_(src)
            return
        if self.ndim == 2 and src.data_ptr() != dst.data_ptr():
            # both passes in one kernel when an item fits in shared memory twice
            wr = []
            for i in range(2):
                if i in axes:
                    key = (float(sigmas[i]), 0)
                    if key not in self._weights:
                        w, radius = gaussian_weights(sigmas[i], 0)
                        self._weights[key] = (torch.from_numpy(w).to(self.device), radius)
                    wr.append(self._weights[key])
                else:
                    wr.append((None, -1))
            done = ctypes.c_int(0)
            _lib.check(_lib.lib.lsml_gauss2d_f64(
                self._shape_c, _lib.c_i64(self.batch), _lib.ptr(src), _lib.ptr(dst),
                _lib.ptr(wr[0][0]), ctypes.c_int(wr[0][1]), _lib.ptr(wr[1][0]), ctypes.c_int(wr[1][1]),
                ctypes.byref(done), self._stream()), "gauss2d")
            if done.value:
                return
        if len(axes) > 1 and self._tmp_plane is None:
            self._tmp_plane = torch.empty_like(self.u)
        tmp = self._tmp_plane
        bufs = [tmp, dst] if len(axes) % 2 == 0 else [dst, tmp]
        for n, i in enumerate(axes):
            out = bufs[n % 2]
            self._gauss(src, out, sigmas[i], 0, i, 0)
            src = out
        assert src.data_ptr() == dst.data_ptr()

    def _refresh_u_planes(self):
        """Planes that depend on the level set (`smooth_u`): recomputed before every feature pass."""
        for p, (what, sigma) in enumerate(self.program.planes):
            if what == "usmooth":                # scipy.ndimage.gaussian_filter(u, sigma): no /dx
                self._smooth_chain(self.u, self.planes[p], [float(sigma)] * self.ndim)

    def _compute_planes(self):
        img_host = None
        for p, (what, sigma) in enumerate(self.program.planes):
            dst = self.planes[p]
            if what == "usmooth":
                continue
            if what == "user":
                # the user's own u-independent image feature, once per image, on the host
                if img_host is None:
                    img_host = self.img.cpu().numpy()
                fn = self.program.user_planes[p]
                for b in range(self.batch):
                    plane = np.asarray(fn(img_host[b], self.dx.copy()), dtype=np.float64)
                    if plane.shape != self.shape:
                        raise ValueError("user plane {!r} has shape {} but must be shape {}".format(
                            sigma, plane.shape, self.shape))
                    dst[b].copy_(torch.from_numpy(np.ascontiguousarray(plane)))
                continue
            if what == "sample":
                if sigma == 0:
                    dst.copy_(self.img)                              # image.py:22-23
                else:                                                # image.py:25-31
                    self._smooth_chain(self.img, dst, [sigma / self.dx[i] for i in range(self.ndim)])
            else:
                if sigma == 0:                                       # image.py:48-49
                    _lib.check(_lib.lib.lsml_np_gradient_mag_f64(
                        ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch),
                        _lib.ptr(self.img), _lib.ptr(dst), self._dx_c, self._stream()),
                        "np_gradient_mag")
                else:                                                # image.py:51-61
                    for a in range(self.ndim):
                        if self.ndim == 1:
                            mode = 4
                        elif a == 0:
                            mode = 1
                        elif a == self.ndim - 1:
                            mode = 3
                        else:
                            mode = 2
                        self._gauss(self.img, dst, sigma / self.dx[a], 1, a, mode)

    # ------------------------------------------------------------------ level sets / band
    def set_level_sets(self, u0, mask=None):
        """u0: (batch, *shape).  With ``mask`` None the band is rebuilt from u0
        (InitializerBase.__call__, initializer_base.py:56-61); otherwise the given band mask is
        used as is (for feature / update calls with a caller-supplied mask)."""
        with torch.cuda.device(self.device):
            t = u0 if torch.is_tensor(u0) else torch.from_numpy(np.ascontiguousarray(u0))
            self.u.copy_(t.to(self.device, dtype=torch.float64).reshape(self.u.shape))
            self._global_stats()
            if mask is None:
                self._rebuild_band()
            else:
                m = mask if torch.is_tensor(mask) else torch.from_numpy(np.ascontiguousarray(mask))
                self.mask.copy_(m.to(self.device).reshape(self.mask.shape).to(torch.uint8))
                self._ws_fmm = None   # the marcher's sparse state no longer matches `mask`
                self._band_is_rebuilt = False
                self._compact()

    def _global_stats(self):
        _lib.check(_lib.lib.lsml_global_stats(
            ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), _lib.ptr(self.u),
            _lib.ptr(self.stats), self._stream()), "global_stats")

    def _rebuild_band(self, dist_out=None, from_band=False):
        """distance_transform(u, band, dx) -> mask, band_idx.  ``from_band``: u has changed only
        on the current band since the rebuild that produced it (the evolution loop), so the
        interface is located from that band instead of a dense pass over u."""
        from_band = from_band and self._band_is_rebuilt and self._ws_fmm is not None
        if self._ws_fmm is None:
            nbytes = int(_lib.lib.lsml_fmm_workspace_bytes(_lib.c_i64(self.total)))
            self._ws_fmm = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.lib.lsml_fmm_counters_offset.restype = _lib.c_i64
            _lib.lib.lsml_fmm_status_offset.restype = _lib.c_i64
            off = (int(_lib.lib.lsml_fmm_counters_offset(_lib.c_i64(self.total))) +
                   int(_lib.lib.lsml_fmm_status_offset()))
            # int64 view {failed_rebuilds, barriers, live sweep, rix, begin, end} (lsml_b200.h)
            self._fmm_status = self._ws_fmm[off:off + 48].view(torch.int64)
            self.mask.zero_()
            _lib.check(_lib.lib.lsml_fmm_workspace_init(
                _lib.c_i64(self.total), _lib.ptr(self._ws_fmm), self._stream()), "fmm_init")
        with self._phase("rebuild"):
            _lib.check(_lib.lib.lsml_fmm_rebuild_from_band(
                ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), self._dx_c,
                _lib.ptr(self.u), ctypes.c_double(self.band), _lib.ptr(self.stats),
                _lib.ptr(self.mask), _lib.ptr(dist_out), _lib.ptr(self._ws_fmm),
                _lib.ptr(self.band_idx if from_band else None),
                _lib.ptr(self.n_band_dev if from_band else None), self._stream()),
                "fmm_rebuild")
        self._compact()
        self._band_is_rebuilt = True

    def check_converged(self):
        """Raise if any band rebuild since the workspace was created ended at its sweep limit
        instead of its fixed point (the mask of such a rebuild is not the reference's).  One
        host read; call it once per segmentation, not per iteration."""
        if self._ws_fmm is None:
            return
        failed = int(self._fmm_status[0].item())
        if failed:
            raise _lib.LsmlError("band rebuild: %d rebuild(s) stopped at the sweep limit before "
                                 "reaching their fixed point" % failed)

    def fmm_counters(self):
        out = (ctypes.c_longlong * 8)()
        _lib.check(_lib.lib.lsml_fmm_counters(_lib.c_i64(self.total), _lib.ptr(self._ws_fmm), out,
                                              self._stream()), "fmm_counters")
        return dict(n_init=out[0], n_cand=out[1], tail_sweeps=out[2], sweeps=out[3],
                    converged=bool(out[4]), replays=out[7])

    def fmm_sweep_sizes(self):
        out = (ctypes.c_longlong * 64)()
        _lib.check(_lib.lib.lsml_fmm_sweep_sizes(_lib.c_i64(self.total), _lib.ptr(self._ws_fmm),
                                                 out, self._stream()), "fmm_sweep_sizes")
        return list(out)

    def distance(self):
        """The reference's ``dist`` (signed distance inside the band, 0 elsewhere)."""
        dist = torch.empty_like(self.u)
        with torch.cuda.device(self.device):
            self._rebuild_band(dist_out=dist)
        return dist

    def _compact(self):
        # two index / count buffers, swapped on every compaction: the previous band stays valid
        # (it is the sparse record of the iteration that just ran) without a copy
        self.band_idx, self._band_idx_prev = self._band_idx_prev, self.band_idx
        self.n_band_dev, self._n_band_prev_dev = self._n_band_prev_dev, self.n_band_dev
        if self.band_idx.numel() < self.total:
            # capacity: the band can be the whole grid (band == 0 or u == 0 everywhere)
            self.band_idx = torch.empty(self.total, dtype=torch.int64, device=self.device)
        with self._phase("compact"):
            self._compact_launch()
        self._n_band_host = None                       # read back only when somebody asks
        self._ensure_band_buffers()

    def _compact_launch(self):
        _lib.check(_lib.lib.lsml_compact_mask(
            _lib.c_i64(self.total), _lib.c_i64(self.voxels), _lib.ptr(self.mask),
            _lib.ptr(self.band_idx), _lib.ptr(self.n_band_dev), _lib.ptr(self.item_offsets),
            _lib.ptr(self._ws_compact), self._stream()), "compact_mask")

    # ------------------------------------------------------------------ features
    def _item_level_features(self):
        """Everything a feature column needs besides the voxel itself: level-set dependent
        planes, band statistics, boundary measure, the per-item scalars and centres of mass."""
        pr = self.program
        npl = len(pr.planes)
        stride = self.total
        if pr.needs_u_planes:
            self._refresh_u_planes()
        if pr.needs_band_stats:
            _lib.check(_lib.lib.lsml_band_stats(
                _lib.c_i64(self.voxels), _lib.c_i64(self.batch), _lib.ptr(self.planes),
                _lib.c_i64(stride), ctypes.c_int(npl), _lib.ptr(self.band_idx),
                _lib.ptr(self.item_offsets), _lib.ptr(self.band_mean), _lib.ptr(self.band_std),
                _lib.ptr(self._ws_band), self._stream()), "band_stats")
        if pr.needs_boundary:
            fn = _lib.lib.lsml_contour_length if self.ndim == 2 else _lib.lib.lsml_surface_area
            _lib.check(fn(
                ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), _lib.ptr(self.u),
                self._dx_c, _lib.ptr(self.boundary), _lib.ptr(self._ws_contour), self._stream()),
                "boundary_size")
        _lib.check(_lib.lib.lsml_item_features(
            ctypes.c_int(pr.nfeat), pr.kind, pr.a, pr.b, ctypes.c_int(self.ndim), self._shape_c,
            _lib.c_i64(self.batch), self._dx_c, ctypes.c_int(max(npl, 1)), _lib.ptr(self.stats),
            _lib.ptr(self.band_mean), _lib.ptr(self.band_std), _lib.ptr(self.boundary),
            _lib.ptr(self.item_feat), _lib.ptr(self.com), self._stream()), "item_features")
        if pr.hi_moments:
            _lib.check(_lib.lib.lsml_axis_moments(
                ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), self._dx_c,
                _lib.ptr(self.u), _lib.ptr(self.stats), _lib.ptr(self.com),
                ctypes.c_int(len(pr.hi_moments)), _lib.ptr(self._hm_cols), _lib.ptr(self._hm_axes),
                _lib.ptr(self._hm_orders), ctypes.c_int(pr.nfeat), _lib.ptr(self.item_feat),
                _lib.ptr(self._ws_moments), self._stream()), "axis_moments")

    def compute_features(self):
        """FeatureMap on the current band -> X[:n_band] (band order), materialised: for
        ``FeatureMap.__call__`` and for the regression matrix of ``fit``.  The evolution loop
        does not come here (see :meth:`step`)."""
        pr = self.program
        nb = self.n_band
        if nb == 0 or pr.nfeat == 0:
            self._ensure_X(1)
            return self.X[:0]
        self._ensure_X(nb)
        self._item_level_features()
        _lib.check(_lib.lib.lsml_assemble_features(
            ctypes.c_int(pr.nfeat), pr.kind, pr.a, pr.b, ctypes.c_int(self.ndim), self._shape_c,
            _lib.c_i64(self.batch), self._dx_c, _lib.ptr(self.band_idx), _lib.ptr(self.n_band_dev),
            _lib.ptr(self.planes), _lib.c_i64(self.total), _lib.ptr(self.item_feat),
            _lib.ptr(self.com), _lib.ptr(self.X), self._stream()), "assemble_features")
        return self.X[:nb]

    def band_source(self, origin0=0):
        """``lsml_band_source``: the feature program and the device buffers it reads, for the
        fused velocity kernels (no materialised matrix)."""
        pr = self.program
        if self._src is None:
            self._src = _lib.BandSource(
                nfeat=pr.nfeat, kind=ctypes.cast(pr.kind, ctypes.c_void_p),
                a=ctypes.cast(pr.a, ctypes.c_void_p), b=ctypes.cast(pr.b, ctypes.c_void_p),
                ndim=self.ndim, shape=ctypes.cast(self._shape_c, ctypes.c_void_p),
                batch=self.batch, dx=ctypes.cast(self._dx_c, ctypes.c_void_p),
                band_idx=0, planes=self.planes.data_ptr(), plane_stride=self.total,
                item_feat=self.item_feat.data_ptr(), com=self.com.data_ptr(), origin0=0)
        self._src.band_idx = self.band_idx.data_ptr()       # swapped on every compaction
        self._src.origin0 = int(origin0)
        return self._src

    # ------------------------------------------------------------------ one iteration
    def step(self, regressor, step, rebuild=True, record=False):
        """One evolution iteration of every item (lsml/core/model.py:381-397).  Nothing in here
        waits for the device: the band size stays in HBM, the velocity kernel evaluates the
        feature columns itself (no feature matrix), and an empty band (`if mask.any()`,
        model.py:381) simply makes every launch a no-op.  ``record``: keep a copy of this
        iteration's sparse record (band indices, new values) in ``self.history`` -- that one
        needs the band size on the host."""
        with torch.cuda.device(self.device):
            reg = export_regressor(regressor, self.device)
            if reg.n_features != self.program.nfeat:
                raise ValueError("regressor expects {} features, feature map has {}".format(
                    reg.n_features, self.program.nfeat))
            self._ensure_band_buffers()
            with self._phase("features"):
                self._item_level_features()
            with self._phase("predict"):
                reg.predict_band(self.band_source(), self.n_band_dev, self.vel)
            # u[mask] += step*v*|Du| and, in the same scatter, the exact {u > 0} statistics of the
            # new u (no dense recount)
            with self._phase("update"):
                _lib.check(_lib.lib.lsml_band_update_stats(
                    ctypes.c_int(self.ndim), self._shape_c, _lib.c_i64(self.batch), self._dx_c,
                    _lib.ptr(self.u), _lib.ptr(self.vel), _lib.ptr(self.band_idx),
                    _lib.ptr(self.n_band_dev), ctypes.c_double(float(step)),
                    _lib.ptr(self.new_u), _lib.c_p(0), _lib.ptr(self.stats), self._stream()),
                    "band_update")
            self.units_dev.add_(self.n_band_dev)
            if record:
                nb = self.n_band
                if self.history is None:
                    self.history = []
                self.history.append((self.band_idx[:nb].clone(), self.new_u[:nb].clone())
                                    if nb else None)
            if rebuild:
                self._rebuild_band(from_band=True)
            else:
                self._band_is_rebuilt = False     # u moved on; the band no longer brackets it

    def last_record(self, nb):
        """(band indices, new values) of the iteration :meth:`step` just ran, ``nb`` being the
        band size BEFORE that step; views, valid until the next step."""
        return self._band_idx_prev[:nb], self.new_u[:nb]
