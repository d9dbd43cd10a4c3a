"""CPU oracle for the lsml level-set evolution hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package ``lsml_b200``
never imports anything under ``oracle/`` and fails loudly without its CUDA library.

Every function restates one piece of the reference (paths relative to ``/root/reference``)
with numpy / scipy / the C library ``oracle/liblsml_oracle.so`` and deliberately keeps the
reference's *work pattern* (full-grid temporaries, the Gaussian recomputed on every call, ...)
so that timing it is an honest stand-in for the reference's CPU path on a box where
``/root/reference`` is absent.  Where the reference's own C file could be compiled
(``oracle/_ref/masked_gradient_ref.so``) it is used for the gradient kernels.

Pinned against the reference itself by ``tests/test_oracle_golden.py`` (the committed fixtures
under ``tests/golden/``, produced by ``oracle/gen_golden.py`` from the imported reference, and the
reference's own C file compiled into ``oracle/_ref``), ``tests/test_host_mirror_vs_reference.py``
and the ``reference`` leg of ``tests/test_reference_suite.py`` (both run only where
``/root/reference`` exists).

PARITY UNPINNED pieces (third-party code absent from ``/root/reference``): the fast-marching
distance (scikit-fmm; ``skfmm_literal.c`` is an upstream-literal restatement used as a measuring
stick), the contour length and the marching-cubes area (scikit-image); see
the ``lsml_oracle.c`` header.
"""
import ctypes
import itertools
import os
import pathlib
import subprocess

import numpy as np
from scipy.ndimage import gaussian_filter1d

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None
_REF = None

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)


def build():
    """Compile the C checker (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-C", str(_HERE), "--no-print-directory"],
                          stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = _HERE / "liblsml_oracle.so"
        if not path.exists():
            build()
        _LIB = ctypes.CDLL(str(path))
        _LIB.orc_fmm_distance.restype = ctypes.c_long
        _LIB.orc_skfmm_literal_distance.restype = ctypes.c_long
        _LIB.orc_marching_squares_length.restype = ctypes.c_double
        _LIB.orc_marching_cubes_area.restype = ctypes.c_double
    return _LIB


def reference_lib():
    """The reference's own compiled C file (None when it was never built)."""
    global _REF
    if _REF is None:
        path = _HERE / "_ref" / "masked_gradient_ref.so"
        if not path.exists():
            return None
        _REF = ctypes.CDLL(str(path))
    return _REF


def _p(a, typ):
    return a.ctypes.data_as(typ)


def _prep(arr, mask, dx):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    mask = (np.ones(arr.shape, dtype=bool) if mask is None
            else np.ascontiguousarray(mask, dtype=bool))
    dx = (np.ones(arr.ndim) if dx is None else np.asarray(dx, dtype=np.float64)).copy()
    shape = np.asarray(arr.shape, dtype=np.int32)
    return arr, mask, dx, shape


# --------------------------------------------------------------------------------------------
# a1 / a2: masked stencils
# --------------------------------------------------------------------------------------------
def gmag_os(arr, nu, mask=None, dx=None, use_reference=False):
    """gradient_magnitude_osher_sethian, lsml/gradient/masked_gradient.py:155-237."""
    arr, mask, dx, shape = _prep(arr, mask, dx)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    out = np.zeros_like(arr)
    ref = reference_lib() if use_reference else None
    if ref is not None:
        fn = getattr(ref, "gmag_os%dd" % arr.ndim)
        fn.restype = None
        args = [ctypes.c_int(int(s)) for s in arr.shape]
        args += [_p(arr, _c_dbl_p), _p(mask.view(np.uint8), _c_u8_p), _p(nu, _c_dbl_p),
                 _p(out, _c_dbl_p)]
        args += [ctypes.c_double(float(d)) for d in dx]
        fn(*args)
    else:
        _lib().orc_gmag_os(ctypes.c_int(arr.ndim), _p(shape, _c_int_p), _p(arr, _c_dbl_p),
                           _p(mask.view(np.uint8), _c_u8_p), _p(nu, _c_dbl_p),
                           _p(out, _c_dbl_p), _p(dx, _c_dbl_p))
    return out


def gradient_centered(arr, mask=None, dx=None, normalize=False, use_reference=False):
    """gradient_centered, lsml/gradient/masked_gradient.py:71-152. Returns ([g...], gmag)."""
    arr, mask, dx, shape = _prep(arr, mask, dx)
    grads = np.zeros((arr.ndim,) + arr.shape)
    gmag = np.zeros_like(arr)
    ref = reference_lib() if use_reference else None
    if ref is not None:
        fn = getattr(ref, "gradient_centered%dd" % arr.ndim)
        fn.restype = None
        args = [ctypes.c_int(int(s)) for s in arr.shape]
        args += [_p(arr, _c_dbl_p), _p(mask.view(np.uint8), _c_u8_p)]
        args += [_p(grads[a], _c_dbl_p) for a in range(arr.ndim)]
        args += [_p(gmag, _c_dbl_p)]
        args += [ctypes.c_double(float(d)) for d in dx]
        args += [ctypes.c_int(int(normalize))]
        fn(*args)
    else:
        _lib().orc_gradient_centered(ctypes.c_int(arr.ndim), _p(shape, _c_int_p),
                                     _p(arr, _c_dbl_p), _p(mask.view(np.uint8), _c_u8_p),
                                     _p(grads, _c_dbl_p), _p(gmag, _c_dbl_p), _p(dx, _c_dbl_p),
                                     ctypes.c_int(int(normalize)))
    return [grads[a] for a in range(arr.ndim)], gmag


# --------------------------------------------------------------------------------------------
# a14: distance transform / band mask
# --------------------------------------------------------------------------------------------
def fmm_distance(phi, dx=None, narrow=0.0, literal=False):
    """skfmm.distance restatement: returns (dist, frozen) with dist 0 where not frozen.
    ``literal``: the upstream-literal marcher of oracle/skfmm_literal.c (float64 heap keys with
    scikit-fmm's own heap mechanics, no causality test, break at the band edge, value2 kept)
    instead of the marcher the product is bit-exact to -- the measuring stick of
    tools/band_gap_report.py, never a golden source."""
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    dx = (np.ones(phi.ndim) if dx is None else np.asarray(dx, dtype=np.float64)).copy()
    if dx.ndim == 0:
        dx = np.full(phi.ndim, float(dx))
    shape = np.asarray(phi.shape, dtype=np.int32)
    dist = np.zeros_like(phi)
    frozen = np.zeros(phi.shape, dtype=np.uint8)
    fn = _lib().orc_skfmm_literal_distance if literal else _lib().orc_fmm_distance
    n = fn(ctypes.c_int(phi.ndim), _p(shape, _c_int_p), _p(phi, _c_dbl_p),
           _p(dx, _c_dbl_p), ctypes.c_double(float(narrow)),
           _p(dist, _c_dbl_p), _p(frozen, _c_u8_p))
    if n < 0:
        raise ValueError("the array phi contains no zero contour (no zero level set)")
    return dist, frozen.astype(bool)


def skfmm_distance(phi, dx=1.0, narrow=0.0, **_):
    """Drop-in for ``skfmm.distance`` (masked array when anything is masked)."""
    phi = np.asarray(phi, dtype=np.float64)
    dxa = np.full(phi.ndim, float(dx)) if np.ndim(dx) == 0 else np.asarray(dx, dtype=float)
    dist, frozen = fmm_distance(phi, dxa, narrow)
    if frozen.all():
        return dist
    return np.ma.MaskedArray(dist, ~frozen)


def distance_transform(arr, band, dx, literal=False):
    """lsml/util/distance_transform.py:5-59 (edge cases :36-47, band==0 :54-57)."""
    arr = np.asarray(arr, dtype=np.float64)
    if (arr == 0).all():
        return np.zeros_like(arr), np.ones(arr.shape, dtype=bool)
    n_pos = (arr > 0).sum()
    if n_pos == arr.size or n_pos == 0:
        mask = np.zeros(arr.shape, dtype=bool)
        sign = np.sign(arr.ravel()[0])
        return sign * np.full(arr.shape, np.inf), mask
    dist, frozen = fmm_distance(arr, dx, band, literal=literal)
    return dist, frozen


# --------------------------------------------------------------------------------------------
# a5-a7: image features (lsml/feature/provided/image.py)
# --------------------------------------------------------------------------------------------
def smooth(img, sigma, dx):
    """image.py:25-29: chain of gaussian_filter1d(sigma/dx[i]) along every axis."""
    out = img.copy()
    for i in range(img.ndim):
        out = gaussian_filter1d(out, sigma=sigma / dx[i], axis=i)
    return out


def image_sample_plane(img, sigma, dx):
    """ImageSample before band sampling (image.py:19-33)."""
    return img if sigma == 0 else smooth(img, sigma, dx)


def image_edge_plane(img, sigma, dx):
    """ImageEdgeSample before band sampling (image.py:45-66)."""
    if sigma == 0:
        grads = np.gradient(img, *dx)
        if img.ndim == 1:
            grads = [grads]
    else:
        grads = [gaussian_filter1d(img, sigma=sigma / dx[a], order=1, axis=a)
                 for a in range(img.ndim)]
    acc = np.zeros_like(img)
    for g in grads:
        acc = acc + g ** 2
    return acc ** 0.5


def center_of_mass(u, dx):
    """Moments._compute_center_of_mass (shape.py:201-212): order-1 moments over {u>0}."""
    return np.array([moment(u, dx, a, 1) for a in range(u.ndim)])


def moment(u, dx, axis, order):
    """Moments._compute_moment (shape.py:214-230)."""
    indices = np.indices(u.shape, dtype=float)
    mesh = indices[axis] * dx[axis]
    if order > 1:
        mesh -= center_of_mass(u, dx)[axis]
    measure = (u > 0).sum() * np.prod(dx)
    return (mesh ** order)[u > 0].sum() * np.prod(dx) / measure


def dist_to_com_plane(u, dx):
    """DistanceToCenterOfMass (shape.py:250-267) on the full grid."""
    com = center_of_mass(u, dx)
    sl = (slice(None),) + (None,) * u.ndim
    mesh = np.indices(u.shape, dtype=float) * dx[sl]
    return np.linalg.norm(mesh - com[sl], axis=0)


def com_ray_columns(u, img, mask, dx, n_samples):
    """COMRaySample.compute_feature (image.py:147-171) on the band: (B, 2*n_samples).

    Same expressions as the reference (numpy.linspace, ``t*com + (1-t)*(coord*dx)``, scipy's
    RegularGridInterpolator with fill_value 0 on the RAW image); the only change of work
    pattern is one interpolator call for all band voxels instead of one per voxel -- the
    interpolation is elementwise, so the values are the same bit for bit
    (tests/test_oracle_golden.py pins this against the reference's own output)."""
    from scipy.interpolate import RegularGridInterpolator
    com = center_of_mass(u, dx)
    t = np.linspace(-1, 1, 2 * n_samples + 2)[1:-1, None]
    interpolator = RegularGridInterpolator(
        points=[np.arange(s, dtype=float) * delta for s, delta in zip(u.shape, dx)],
        values=img, bounds_error=False, fill_value=0.0)
    coords = np.array(np.where(mask)).T
    points = t[None] * com[None, None] + (1 - t)[None] * (coords * dx)[:, None, :]
    return interpolator(points.reshape(-1, u.ndim)).reshape(coords.shape[0], 2 * n_samples)


def boundary_size(u, dx):
    """BoundarySize (shape.py:56-89): marching-squares length (2-D), marching-cubes area (3-D;
    orc_marching_cubes_area -- scikit-image's Lewiner variant is absent: parity unpinned)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    dx = np.asarray(dx, dtype=np.float64).copy()
    if u.ndim == 3:
        return float(_lib().orc_marching_cubes_area(
            ctypes.c_int(u.shape[0]), ctypes.c_int(u.shape[1]), ctypes.c_int(u.shape[2]),
            _p(u, _c_dbl_p), _p(dx, _c_dbl_p)))
    if u.ndim != 2:
        raise ValueError("Boundary size is only defined for dimensions 2 and 3")
    return float(_lib().orc_marching_squares_length(
        ctypes.c_int(u.shape[0]), ctypes.c_int(u.shape[1]), _p(u, _c_dbl_p), _p(dx, _c_dbl_p),
        ctypes.c_int(1)))


# Feature "specs": plain tuples, shared vocabulary between tests, bench and the product's
# feature classes (lsml_b200.feature.*.spec()).
#   ("image_sample", sigma) ("image_edge", sigma) ("band_mean", sigma) ("band_std", sigma)
#   ("size",) ("boundary_size",) ("isoperimetric",) ("moments", axes, orders) ("dist_to_com",)
#   ("com_ray", n_samples) ("smooth_u", sigma) ("user_plane", key, fn(img, dx) -> plane)
def spec_width(spec):
    if spec[0] == "com_ray":
        return 2 * int(spec[1])
    return len(spec[1]) * len(spec[2]) if spec[0] == "moments" else 1


def feature_columns(spec, u, img, dist, mask, dx):
    """One feature, reference work pattern: returns (B, width) float64 in band order."""
    kind = spec[0]
    dx = np.asarray(dx, dtype=float)
    nb = int(mask.sum())
    if kind == "image_sample":
        return image_sample_plane(img, spec[1], dx)[mask][:, None]
    if kind == "image_edge":
        return image_edge_plane(img, spec[1], dx)[mask][:, None]
    if kind == "band_mean":      # InteriorImageAverage: mean over the BAND (image.py:78-93)
        return np.full((nb, 1), image_sample_plane(img, spec[1], dx)[mask].mean())
    if kind == "band_std":       # InteriorImageVariation (image.py:106-121), ddof=0
        return np.full((nb, 1), image_sample_plane(img, spec[1], dx)[mask].std())
    if kind == "size":           # shape.py:26-32
        return np.full((nb, 1), (u > 0).sum() * np.prod(dx))
    if kind == "boundary_size":
        return np.full((nb, 1), boundary_size(u, dx))
    if kind == "isoperimetric":  # shape.py:123-153
        size = (u > 0).sum() * np.prod(dx)
        bs = boundary_size(u, dx)
        if u.ndim == 2:
            return np.full((nb, 1), 4 * np.pi * size / bs ** 2)
        return np.full((nb, 1), 36 * np.pi * size ** 2 / bs ** 3)
    if kind == "moments":        # shape.py:232-238
        cols = [moment(u, dx, a, o) for a, o in itertools.product(spec[1], spec[2])]
        return np.tile(np.asarray(cols)[None, :], (nb, 1))
    if kind == "dist_to_com":
        return dist_to_com_plane(u, dx)[mask][:, None]
    if kind == "com_ray":
        return com_ray_columns(u, img, mask, dx, int(spec[1]))
    if kind == "smooth_u":       # examples/gestalt2d/prev_iter_feature.py:21-22
        from scipy.ndimage import gaussian_filter
        return gaussian_filter(u, sigma=spec[1])[mask][:, None]
    if kind == "user_plane":     # a user's u-independent image feature (e.g. corner_feature.py)
        return np.asarray(spec[2](img, dx), dtype=np.float64)[mask][:, None]
    raise ValueError("unknown feature spec %r" % (spec,))


def feature_matrix(specs, u, img, dist, mask, dx):
    """FeatureMap.__call__(...)[mask] (feature_map.py:53-80): (B, F) float64, band order."""
    cols = [feature_columns(s, u, img, dist, mask, dx) for s in specs]
    return np.concatenate(cols, axis=1) if cols else np.zeros((int(mask.sum()), 0))


# --------------------------------------------------------------------------------------------
# a13 / a15: update and the segment loop (lsml/core/model.py:378-398)
# --------------------------------------------------------------------------------------------
def evolve_step(u, img, dist, mask, dx, specs, regressor, step, band, use_reference=False,
                rebuild=True):
    """One iteration.  Returns (u_new, dist_new, mask_new, velocity_on_band)."""
    u_new = u.copy()
    if not mask.any():
        return u_new, dist, mask, np.zeros(0)
    feats = feature_matrix(specs, u, img, dist, mask, dx)
    v = np.zeros(u.shape)
    vb = np.asarray(regressor.predict(feats), dtype=np.float64)
    v[mask] = vb
    g = gmag_os(u, v, mask, dx, use_reference=use_reference)
    u_new[mask] += step * v[mask] * g[mask]
    if rebuild:
        dist, mask = distance_transform(u_new, band, dx)
    return u_new, dist, mask, vb


def normalize_image(img):
    """model.py:328-329."""
    return (img - img.mean()) / img.std()


def ball_init(shape, radius, dx, location=None):
    """BallInitializer.initialize (initializer/provided/ball.py:13-31) -> u0 = 2*m-1."""
    dx = np.asarray(dx, dtype=float)
    loc = 0.5 * np.array(shape) if location is None else np.asarray(location, dtype=float)
    sl = (slice(None),) + (None,) * len(shape)
    ind = np.indices(shape, dtype=float)
    ind *= dx[sl]
    ind -= (loc * dx)[sl]
    m = (radius - np.sqrt((ind ** 2).sum(axis=0))) > 0
    return 2 * m.astype(float) - 1


def segment(img, u0, dx, specs, regressors, step, band, normalize=True, use_reference=False):
    """model.py:283-411 with the initial level set given.  Returns (us, masks, velocities)."""
    img_ = normalize_image(img) if normalize else img
    dx = np.asarray(dx, dtype=float)
    dist, mask = distance_transform(u0, band, dx)
    us, masks, vels = [u0.copy()], [mask], []
    u = u0
    for reg in regressors:
        u, dist, mask, vb = evolve_step(u, img_, dist, mask, dx, specs, reg, step, band,
                                        use_reference=use_reference)
        us.append(u)
        masks.append(mask)
        vels.append(vb)
    return np.array(us), masks, vels


def jaccard(u, seg):
    """lsml/score_functions.py:4-23."""
    t = u > 0
    union = float((t | seg).sum())
    return 1.0 if union == 0 else float((t & seg).sum()) / union
