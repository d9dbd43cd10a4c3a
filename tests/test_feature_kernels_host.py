"""CPU-only: the feature kernels' own source (lsml_b200/csrc/features.cu, marked region:
FeatProgram / ItemGeom, moment_from_stats, item_features_kernel, assemble_kernel -- with
com_ray.cuh) compiled for the host and run on one "thread", against the oracle's feature matrix
and the reference's golden vectors.  Covers what tests/test_com_ray_host.py cannot: the wiring
of a COMRaySample column inside assemble_kernel (plane pointer, packed sample index, voxel index
decode, per-item centre of mass from the exact integer statistics) on batches."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from golden_util import load

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "lsml_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    regions = []
    for name in ("band_features.cuh", "features.cu"):    # band_feature(), then the two kernels
        src = open(os.path.join(CSRC, name)).read()
        found = re.findall(r"// \[host-test:begin\][^\n]*\n(.*?)// \[host-test:end\]", src, re.S)
        assert len(found) == 1, name
        regions += found
    text = ('#include "cuda_host_shim.h"\n#include "com_ray.cuh"\nnamespace lsml {\n' +
            "\n".join(regions) +
            open(os.path.join(HERE, "native", "features_host_driver.inc")).read())
    out = tmp_path_factory.mktemp("features_host")
    cpp, so = str(out / "features_host.cpp"), str(out / "features_host.so")
    with open(cpp, "w") as f:
        f.write(text)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                           "-I", os.path.join(HERE, "native"), "-I", CSRC, cpp, "-o", so])
    lib = ctypes.CDLL(so)
    lib.features_host.restype = ctypes.c_int
    return lib


def run_feature_kernels(lib, oracle, specs, us, imgs, masks, dx):
    """The engine's host-side preparation in numpy (planes, exact statistics, band statistics),
    then the two kernels.  Returns X (n_band, F) in band order over the whole batch."""
    from lsml_b200.core.engine import FeatureProgram
    batch, shape = us.shape[0], us.shape[1:]
    ndim, pad = len(shape), 3 - len(shape)
    prog = FeatureProgram(specs, ndim)
    n = np.ones(3, dtype=np.int32); n[pad:] = shape
    dxp = np.ones(3); dxp[pad:] = dx
    voxels = int(np.prod(shape))
    npl = max(len(prog.planes), 1)
    planes = np.zeros((npl, batch, voxels))
    for p, (what, sigma) in enumerate(prog.planes):
        for b in range(batch):
            pl = (oracle.image_sample_plane(imgs[b], sigma, dx) if what == "sample"
                  else oracle.image_edge_plane(imgs[b], sigma, dx))
            planes[p, b] = pl.ravel()
    stats = np.zeros((batch, 8), dtype=np.uint64)
    for b in range(batch):
        pos = us[b] > 0
        idx = np.indices(shape)
        stats[b, 0] = pos.sum()
        for a in range(ndim):
            stats[b, 1 + pad + a] = int(idx[a][pos].sum())
            stats[b, 4 + pad + a] = int((idx[a][pos].astype(np.int64) ** 2).sum())
        stats[b, 7] = (us[b] == 0).sum()
    band_idx = np.flatnonzero(masks.reshape(batch, -1).ravel()).astype(np.int64)
    band_mean, band_std = np.zeros((batch, npl)), np.zeros((batch, npl))
    for b in range(batch):
        m = masks[b].ravel()
        for p in range(len(prog.planes)):
            band_mean[b, p], band_std[b, p] = planes[p, b][m].mean(), planes[p, b][m].std()
    item_feat = np.zeros((batch, prog.nfeat))
    com = np.zeros((batch, 3))
    X = np.zeros((len(band_idx), prog.nfeat))
    P = lambda arr, t: arr.ctypes.data_as(ctypes.POINTER(t))
    lib.features_host(
        ctypes.c_int(prog.nfeat), prog.kind, prog.a, prog.b, ctypes.c_int(ndim), P(n, ctypes.c_int),
        P(dxp, ctypes.c_double), ctypes.c_longlong(batch), ctypes.c_int(npl),
        P(stats, ctypes.c_ulonglong), P(band_mean, ctypes.c_double), P(band_std, ctypes.c_double),
        None, P(band_idx, ctypes.c_longlong), ctypes.c_longlong(len(band_idx)),
        P(planes, ctypes.c_double), ctypes.c_longlong(batch * voxels), P(item_feat, ctypes.c_double),
        P(com, ctypes.c_double), P(X, ctypes.c_double))
    return X


def test_com_ray_columns_inside_assemble_kernel(oracle, host_lib):
    g = load("com_ray.npz")
    for c in range(int(g["n_cases"])):
        u, img, mask, dx = g["u%d" % c], g["img%d" % c], g["mask%d" % c], g["dx%d" % c]
        n_samples = int(g["n_samples%d" % c])
        specs = [("image_sample", 0.0), ("com_ray", n_samples), ("size",)]
        X = run_feature_kernels(host_lib, oracle, specs, u[None], img[None], mask[None], dx)
        assert np.array_equal(X[:, 1:1 + 2 * n_samples], g["X%d" % c])
        assert np.array_equal(X[:, 0], img[mask])
        assert np.array_equal(X[:, -1], np.full(int(mask.sum()), (u > 0).sum() * np.prod(dx)))


def test_feature_program_on_a_batch_against_oracle(oracle, host_lib):
    rs = np.random.RandomState(5)
    shape, batch, n_samples = (14, 17, 15), 3, 3
    dx = np.array([1.0, 1.0, 1.0])
    gg = np.indices(shape).astype(float)
    us, imgs, masks = [], [], []
    for b in range(batch):
        c = np.array(shape) / 2.0 + rs.uniform(-2, 2, size=3)
        r = np.sqrt(sum((gg[a] - c[a]) ** 2 for a in range(3)))
        us.append(rs.uniform(3, 5) - r + 0.2 * rs.randn(*shape))
        imgs.append(rs.randn(*shape))
        masks.append(oracle.distance_transform(us[-1], 3.0, dx)[1])
    us, imgs, masks = np.array(us), np.array(imgs), np.array(masks)
    specs = [("image_sample", 1.0), ("com_ray", n_samples), ("dist_to_com",), ("image_edge", 0.0),
             ("band_mean", 1.0), ("band_std", 1.0), ("moments", (0, 1, 2), (1, 2)), ("size",)]
    X = run_feature_kernels(host_lib, oracle, specs, us, imgs, masks, dx)
    row = 0
    for b in range(batch):
        want = oracle.feature_matrix(specs, us[b], imgs[b], None, masks[b], dx)
        got = X[row:row + want.shape[0]]
        row += want.shape[0]
        ray = slice(1, 1 + 2 * n_samples)
        assert np.array_equal(got[:, ray], want[:, ray])                      # COM ray: bit exact
        assert np.array_equal(got[:, 0], want[:, 0])                          # plane sample
        assert np.array_equal(got[:, 1 + 2 * n_samples], want[:, 1 + 2 * n_samples])   # dist to COM
        assert np.allclose(got, want, rtol=1e-12, atol=1e-13)                 # moments: sum order
    assert row == X.shape[0]


def test_reference_com_ray_test_through_the_kernel_source(oracle, host_lib):
    """lsml/feature/provided/test/test_image.py:159-189 (non-dyadic dx, a single band voxel) on
    the kernels' source: what tests/test_reference_suite.py asserts on the GPU."""
    x, dx = np.linspace(-2, 2, 301, retstep=True)
    y, dy = np.linspace(-2, 2, 201, retstep=True)
    xx, yy = np.meshgrid(x, y)
    img = 1 - np.sqrt(xx ** 2 + yy ** 2)
    mask = abs(img) < 0.01
    ii, jj = np.where(mask)
    mask[ii[1:], jj[1:]] = False
    X = run_feature_kernels(host_lib, oracle, [("com_ray", 10)], img[None], img[None], mask[None],
                            np.array([dy, dx]))
    samples = X[0]
    _, dt = np.linspace(-1, 1, 22, retstep=True)
    ds = np.diff(samples)
    assert round(abs(ds.var()), 10) == 0
    assert round(abs((ds / dt).mean() - 1), 1) == 0
    want = oracle.feature_matrix([("com_ray", 10)], img, img, None, mask, np.array([dy, dx]))
    assert np.allclose(X, want, rtol=1e-12, atol=1e-13)
