"""GPU parity at the sizes BASELINE.json names ("GPU parity tests stop at toy sizes").

* C3: one iteration of the bench recipe on the full 256^3 volume against the oracle -- initial
  band (mask AND band indices), forest velocities, updated level set, rebuilt band.
* C2: 64 images of 101^2, the bench's C2 feature list (incl. the marching-squares features),
  RF-100, 10 iterations, every image against the oracle.
* C4: 512 crops of 64^3 in ONE engine, 30 iterations with the bench's forest; 16 crops spread
  over the batch are evolved by the oracle from the same images.
* `segment()` on an EXPANDING 3-D ball whose band outgrows its initial size several times
  (the sparse record must survive a buffer growth), all iterates against the oracle.

Bars: masks / band indices bit-exact; level sets 1e-9 absolute (identical expressions; only the
band mean / std reductions differ in summation order, 1e-12 relative, before the forest's
float32 cast).
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _forest(X, y, seed, n_estimators=100):
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    rs = np.random.RandomState(seed)
    sub = rs.choice(X.shape[0], size=min(X.shape[0], 20000), replace=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return Pipeline([("standardscaler", StandardScaler()),
                         ("regressor", RandomForestRegressor(n_estimators=n_estimators,
                                                             max_samples=300, random_state=rs))]
                        ).fit(X[sub], y[sub])


def test_c3_one_iteration_at_256_cubed(oracle):
    import bench
    from lsml_b200.core.engine import BandEngine
    O = oracle
    n = 256
    img, radius, c = bench.make_volume(n, 1234)
    dx = np.ones(3)
    eng = BandEngine((n, n, n), 1, dx=dx, band=bench.BAND, specs=bench.SPECS)
    eng.load_images(img[None], normalize=True)
    u0 = O.ball_init((n, n, n), bench.INIT_RADIUS, dx)
    eng.set_level_sets(u0[None])
    img_n = O.normalize_image(img)
    assert np.allclose(eng.img[0].cpu().numpy(), img_n, rtol=0, atol=1e-12)
    dist, mask = O.distance_transform(u0, bench.BAND, dx)
    # band mask and band INDICES (numpy boolean-index order, model.py:388)
    assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask)
    nb = eng.n_band
    assert nb == int(mask.sum())
    assert np.array_equal(eng.band_idx[:nb].cpu().numpy(), np.flatnonzero(mask.ravel()))
    Xg = eng.compute_features().cpu().numpy()
    Xo = O.feature_matrix(bench.SPECS, u0, img_n, dist, mask, dx)
    assert np.allclose(Xg, Xo, rtol=1e-9, atol=1e-9)
    y = bench.target_distance_on_band(np.flatnonzero(mask.ravel()), n, radius, c)
    reg = _forest(Xg, y, 1234)
    eng.step(reg, bench.STEP)
    u1, _, mask1, vb = O.evolve_step(u0, img_n, dist, mask, dx, bench.SPECS, reg, bench.STEP,
                                     bench.BAND)
    assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask1)
    nb1 = eng.n_band
    assert np.array_equal(eng.band_idx[:nb1].cpu().numpy(), np.flatnonzero(mask1.ravel()))
    ug = eng.u[0].cpu().numpy()
    assert np.array_equal(ug > 0, u1 > 0)
    assert np.allclose(ug, u1, rtol=1e-5, atol=1e-9), np.abs(ug - u1).max()
    eng.check_converged()


def test_c2_batch_of_101_squared_images(oracle):
    import bench
    import torch
    from lsml_b200.core.engine import BandEngine
    O = oracle
    shape, count, iters = (101, 101), 64, 10
    dx = np.ones(2)
    imgs = bench._blob_items(count, shape, 20000, 15, 35, 10)
    specs = bench.C2_SPECS
    eng = BandEngine(shape, count, dx=dx, band=bench.BAND, specs=specs)
    eng.load_images(imgs, normalize=True)
    g = np.indices(shape).astype(np.float64)
    rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(2)))
    u0 = np.where(rc < 10.0, 1.0, -1.0)
    eng.set_level_sets(np.broadcast_to(u0, (count,) + shape).copy())
    X = eng.compute_features().cpu().numpy()
    y = X[:, 1] + 0.2 * np.random.RandomState(1).standard_normal(X.shape[0])
    reg = _forest(X, y, 5)
    for _ in range(iters):
        eng.step(reg, 0.4)
    eng.check_converged()
    ug, mg = eng.u.cpu().numpy(), eng.mask.cpu().numpy().astype(bool)
    for b in range(count):
        img_n = O.normalize_image(imgs[b])
        u = u0.copy()
        dist, mask = O.distance_transform(u, bench.BAND, dx)
        for _ in range(iters):
            u, dist, mask, _ = O.evolve_step(u, img_n, dist, mask, dx, specs, reg, 0.4, bench.BAND)
        assert np.array_equal(mg[b], mask), (b, int((mg[b] != mask).sum()))
        assert np.array_equal(ug[b] > 0, u > 0), b
        assert np.allclose(ug[b], u, rtol=1e-5, atol=1e-9), (b, np.abs(ug[b] - u).max())


def test_c4_512_crops_30_iterations(oracle):
    import bench
    import torch
    from lsml_b200.core.engine import BandEngine
    O = oracle
    shape, count = (64, 64, 64), 512
    dx = np.ones(3)
    dev = torch.device("cuda")
    host_reg, reg = bench.c4_forest(dev)
    raw = bench.c4_crops_device(0, count, dev)
    eng = BandEngine(shape, count, dx=dx, band=bench.BAND, specs=bench.SPECS)
    eng.load_images(raw, normalize=True)
    u0_t = bench.c4_initial_level_set(dev)
    eng.set_level_sets(u0_t.expand((count,) + shape).contiguous())
    for _ in range(bench.C4_ITERS):
        eng.step(reg, 0.4)
    eng.check_converged()
    sample = [0, 1, 37, 64, 100, 129, 200, 255, 256, 300, 333, 384, 420, 470, 500, 511]
    u0 = u0_t.cpu().numpy()
    for b in sample:
        img_n = O.normalize_image(raw[b].cpu().numpy())
        u = u0.copy()
        dist, mask = O.distance_transform(u, bench.BAND, dx)
        for _ in range(bench.C4_ITERS):
            u, dist, mask, _ = O.evolve_step(u, img_n, dist, mask, dx, bench.SPECS, host_reg, 0.4,
                                             bench.BAND)
        mg = eng.mask[b].cpu().numpy().astype(bool)
        ug = eng.u[b].cpu().numpy()
        assert np.array_equal(mg, mask), (b, int((mg != mask).sum()))
        assert np.array_equal(ug > 0, u > 0), b
        assert np.allclose(ug, u, rtol=1e-5, atol=1e-9), (b, np.abs(ug - u).max())


def test_segment_iterates_survive_a_growing_band(oracle):
    """An expanding ball: the band grows from ~3.5 k to > 20 k voxels, far beyond any capacity
    derived from the first band; every iterate `segment` returns equals the oracle's."""
    from sklearn.linear_model import LinearRegression
    from lsml_b200 import LevelSetMachineLearning
    from lsml_b200.feature import Size, get_basic_image_features
    from lsml_b200.initializer import BallInitializer
    O = oracle
    n, iters = 56, 12
    rs = np.random.RandomState(3)
    g = np.indices((n, n, n)).astype(float) - n / 2
    img = (np.sqrt((g ** 2).sum(0)) < 22).astype(float) + 0.1 * rs.randn(n, n, n)
    feats = get_basic_image_features(ndim=3, sigmas=[0, 2]) + [Size(ndim=3)]
    X = rs.randn(32, 9)
    reg = LinearRegression().fit(X, 2.5 + 0.0 * X[:, 0])       # constant outward velocity
    model = LevelSetMachineLearning.from_regressors(feats, BallInitializer(radius=6), [reg] * iters,
                                                    step=0.5, band=3)
    specs = [tuple(s) for s in model.feature_map.specs()]
    assert len(specs) == 9
    us = model.segment(img, verbose=False, iterate_until_validation_max=False)
    u0 = O.ball_init((n, n, n), 6.0, np.ones(3))
    want, _, _ = O.segment(img, u0, np.ones(3), specs, [reg] * iters, 0.5, 3.0)
    assert us.shape == want.shape
    assert (want[-1] > 0).sum() > 8 * (want[0] > 0).sum(), "the ball did not expand"
    for i in range(us.shape[0]):
        assert np.array_equal(us[i] > 0, want[i] > 0), i
        assert np.allclose(us[i], want[i], rtol=1e-5, atol=1e-9), (i, np.abs(us[i] - want[i]).max())
