"""CPU-only checks of COMRaySample:

* the oracle restatement against the golden vectors of the unmodified reference
  (tests/golden/com_ray.npz, oracle/gen_golden.py);
* the product's per-sample device function (lsml_b200/csrc/com_ray.cuh, ``__host__
  __device__``) compiled for the host with g++ and checked bit for bit against the same vectors
  -- the arithmetic the CUDA kernel runs, exercised where no GPU exists.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from golden_util import load

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def cases():
    g = load("com_ray.npz")
    for c in range(int(g["n_cases"])):
        yield (g["u%d" % c], g["img%d" % c], g["mask%d" % c], g["dx%d" % c],
               int(g["n_samples%d" % c]), g["X%d" % c])


def test_oracle_reproduces_reference(oracle):
    n = 0
    for u, img, mask, dx, n_samples, want in cases():
        got = oracle.feature_matrix([("com_ray", n_samples)], u, img, None, mask, dx)
        assert got.shape == want.shape == (int(mask.sum()), 2 * n_samples)
        assert np.array_equal(got, want)
        assert (want == 0).any() and (want != 0).any()     # rays leave the grid in every case
        n += 1
    assert n == 4


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("com_ray") / "com_ray_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "lsml_b200", "csrc"),
                           os.path.join(HERE, "native", "com_ray_host.cpp"), "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    lib.com_ray_host.restype = None
    return lib


def host_com_ray(lib, img, dx, com, mask, n_samples):
    ndim = img.ndim
    pad = 3 - ndim
    n = np.ones(3, dtype=np.int32); n[pad:] = img.shape
    dxp = np.ones(3); dxp[pad:] = dx
    comp = np.zeros(3); comp[pad:] = com
    coords = np.zeros((int(mask.sum()), 3), dtype=np.int64)
    coords[:, pad:] = np.array(np.where(mask)).T
    out = np.empty((coords.shape[0], 2 * n_samples))
    img = np.ascontiguousarray(img, dtype=np.float64)
    p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    lib.com_ray_host(p(img, ctypes.c_double), ctypes.c_int(ndim), p(n, ctypes.c_int),
                     p(dxp, ctypes.c_double), p(comp, ctypes.c_double),
                     p(coords, ctypes.c_longlong), ctypes.c_longlong(coords.shape[0]),
                     ctypes.c_int(n_samples), p(out, ctypes.c_double))
    return out


def test_device_function_on_host_bit_exact(oracle, host_lib):
    for u, img, mask, dx, n_samples, want in cases():
        com = oracle.center_of_mass(u, dx)
        got = host_com_ray(host_lib, img, dx, com, mask, n_samples)
        assert np.array_equal(got, want)


def test_device_function_edge_cases(oracle, host_lib):
    rs = np.random.RandomState(2)
    img = rs.randn(9, 11)
    mask = np.ones(img.shape, dtype=bool)
    dx = np.array([1.0, 0.5])
    # a centre of mass exactly on grid points / on the last grid line / outside the grid
    for com in ([4.0, 2.5], [8.0, 5.0], [0.0, 0.0], [20.0, -3.0], [3.3, 1.7]):
        com = np.asarray(com)
        u = np.zeros(img.shape)         # not used by the host function; oracle needs com only
        got = host_com_ray(host_lib, img, dx, com, mask, 3)
        # the same expression through scipy, with the centre of mass forced
        from scipy.interpolate import RegularGridInterpolator
        t = np.linspace(-1, 1, 8)[1:-1, None]
        f = RegularGridInterpolator([np.arange(s, dtype=float) * d for s, d in zip(img.shape, dx)],
                                    img, bounds_error=False, fill_value=0.0)
        coords = np.array(np.where(mask)).T
        want = np.array([f(t * com[None] + (1 - t) * (c * dx)[None]) for c in coords])
        assert np.array_equal(got, want), com
    # no {u > 0} voxel: the centre of mass is NaN and so is every sample (scipy: f(nan) = nan)
    got = host_com_ray(host_lib, img, dx, np.array([np.nan, np.nan]), mask, 2)
    assert np.isnan(got).all()
