"""CPU-only: the engine's PYTHON orchestration of the image / level-set planes
(BandEngine._compute_planes, _smooth_chain, _refresh_u_planes: which kernel is called on which
buffer along which axis with which weights, the ping-pong buffers, the weights cache, user
planes) with the two plane kernels replaced by scipy / numpy stand-ins that work on the same raw
pointers.  The kernels themselves are checked bit for bit on the GPU (tests/test_engine_gpu.py,
tests/test_golden_gpu.py); this test pins the host logic around them against the oracle's planes
where no GPU exists."""
import ctypes
import types

import numpy as np
import pytest
import torch
from scipy.ndimage import correlate1d


def _view(ptr, shape, dtype=np.float64):
    n = int(np.prod(shape))
    ctype = ctypes.c_double if dtype == np.float64 else ctypes.c_int
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,)).reshape(shape)


class FakeLib:
    """lsml_gauss1d_f64 / lsml_np_gradient_mag_f64 with the C ABI's argument lists
    (include/lsml_b200.h), computed by scipy / numpy on host memory."""

    def __init__(self):
        self.calls = []

    def lsml_gauss1d_f64(self, ndim, shape, batch, src, dst, w, radius, antisym, axis, mode, stream):
        ndim, batch, radius, axis, mode = ndim.value, batch.value, radius.value, axis.value, mode.value
        full = (batch,) + tuple(shape[i] for i in range(ndim))
        x = _view(src, full).copy()
        weights = _view(w, (2 * radius + 1,)).copy()
        out = _view(dst, full)
        t = correlate1d(x, weights, axis=axis + 1, mode="reflect")      # what gaussian_filter1d runs
        if mode == 0:
            out[...] = t
        elif mode == 1:
            out[...] = t * t
        elif mode == 2:
            out[...] = out + t * t
        elif mode == 3:
            out[...] = np.sqrt(out + t * t)
        else:
            out[...] = np.sqrt(t * t)
        self.calls.append(("gauss1d", axis, mode, radius, antisym.value))
        return 0

    def lsml_gauss2d_f64(self, shape, batch, src, dst, w0, r0, w1, r1, done, stream):
        """Both passes of a small 2-D item in one call (axis 0 first, then axis 1); r < 0 skips."""
        batch, r0, r1 = batch.value, r0.value, r1.value
        full = (batch, shape[0], shape[1])
        t = _view(src, full).copy()
        if r0 >= 0:
            t = correlate1d(t, _view(w0, (2 * r0 + 1,)).copy(), axis=1, mode="reflect")
        if r1 >= 0:
            t = correlate1d(t, _view(w1, (2 * r1 + 1,)).copy(), axis=2, mode="reflect")
        _view(dst, full)[...] = t
        done._obj.value = 1
        self.calls.append(("gauss2d", r0, r1))
        return 0

    def lsml_np_gradient_mag_f64(self, ndim, shape, batch, src, dst, dx, stream):
        ndim, batch = ndim.value, batch.value
        full = (batch,) + tuple(shape[i] for i in range(ndim))
        x, out = _view(src, full), _view(dst, full)
        spacing = [dx[i] for i in range(ndim)]
        for b in range(batch):
            grads = np.gradient(x[b], *spacing)
            grads = [grads] if ndim == 1 else grads
            acc = np.zeros_like(x[b])
            for g in grads:
                acc = acc + g ** 2
            out[b] = acc ** 0.5
        return 0


def make_engine(monkeypatch, shape, batch, dx, specs):
    """A BandEngine object without its CUDA constructor: only what the plane code touches."""
    from lsml_b200 import _lib
    from lsml_b200.core import engine as E
    fake = FakeLib()
    monkeypatch.setattr(_lib, "lib", fake)
    monkeypatch.setattr(_lib, "current_stream", lambda: ctypes.c_void_p(0))
    eng = E.BandEngine.__new__(E.BandEngine)
    eng.device = torch.device("cpu")
    eng.shape, eng.ndim, eng.batch = tuple(shape), len(shape), batch
    eng.dx = np.asarray(dx, dtype=np.float64)
    eng.program = E.FeatureProgram(specs, eng.ndim)
    eng._shape_c = _lib.int_array(eng.shape)
    eng._dx_c = _lib.double_array(eng.dx)
    full = (batch,) + eng.shape
    eng.u = torch.zeros(full, dtype=torch.float64)
    eng.planes = torch.zeros((max(len(eng.program.planes), 1),) + full, dtype=torch.float64)
    eng._weights, eng._tmp_plane = {}, None
    return eng, fake


@pytest.mark.parametrize("shape,dx", [((23,), (0.8,)), ((17, 21), (1.0, 2.0)), ((9, 12, 10), (1.0, 0.5, 1.6))])
def test_plane_orchestration_matches_oracle_planes(monkeypatch, oracle, shape, dx):
    from scipy.ndimage import gaussian_filter
    rs = np.random.RandomState(len(shape))
    batch = 2
    dx = np.asarray(dx)
    calls = []

    def user_plane(img, dxx):
        calls.append(img.shape)
        return img * 3.0 - 1.0

    specs = [("image_sample", 0.0), ("image_sample", 1.5), ("image_edge", 0.0), ("image_edge", 2.0),
             ("band_mean", 1.5), ("smooth_u", 0.0), ("smooth_u", 1.25), ("user_plane", "mine", user_plane)]
    eng, fake = make_engine(monkeypatch, shape, batch, dx, specs)
    imgs = rs.randn(batch, *shape)
    eng.img = torch.from_numpy(imgs.copy())
    eng._compute_planes()
    us = rs.randn(batch, *shape)
    eng.u.copy_(torch.from_numpy(us))
    eng._refresh_u_planes()
    planes = dict(zip(eng.program.planes, eng.planes.numpy()))
    assert list(planes) == [("sample", 0.0), ("sample", 1.5), ("edge", 0.0), ("edge", 2.0),
                            ("usmooth", 0.0), ("usmooth", 1.25), ("user", "mine")]
    for b in range(batch):
        assert np.array_equal(planes[("sample", 0.0)][b], imgs[b])
        assert np.array_equal(planes[("sample", 1.5)][b], oracle.image_sample_plane(imgs[b], 1.5, dx))
        assert np.array_equal(planes[("edge", 0.0)][b], oracle.image_edge_plane(imgs[b], 0.0, dx))
        assert np.array_equal(planes[("edge", 2.0)][b], oracle.image_edge_plane(imgs[b], 2.0, dx))
        assert np.array_equal(planes[("usmooth", 0.0)][b], us[b])
        assert np.array_equal(planes[("usmooth", 1.25)][b], gaussian_filter(us[b], sigma=1.25))
        assert np.array_equal(planes[("user", "mine")][b], imgs[b] * 3.0 - 1.0)
    assert calls == [tuple(shape)] * batch                    # once per image
    # a second refresh (next iteration) reuses the uploaded weights and overwrites the planes
    n_weights = len(eng._weights)
    eng.u.mul_(0.5)
    eng._refresh_u_planes()
    assert len(eng._weights) == n_weights
    assert np.array_equal(eng.planes.numpy()[5][0], gaussian_filter(us[0] * 0.5, sigma=1.25))
    assert np.array_equal(eng.planes.numpy()[1][0], oracle.image_sample_plane(imgs[0], 1.5, dx))
