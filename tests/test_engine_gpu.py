"""GPU parity of the per-iteration pipeline against the CPU oracle (features, regressors,
update, band rebuild).  Every comparison states its bar:

* integer / index / mask results: bit-exact
* image planes, local features, tree-ensemble velocities, updated level sets given identical
  inputs: bit-exact (same float64 expressions in the same order)
* global features that are reductions (band mean/std, 2nd moments, contour length):
  relative 1e-12 (summation order differs from numpy's pairwise sum)
"""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12

IMG_SPECS = [("image_sample", 0), ("image_sample", 2), ("image_edge", 0), ("image_edge", 2),
             ("band_mean", 0), ("band_mean", 2), ("band_std", 0), ("band_std", 2)]


def shape_specs(ndim, boundary):
    if boundary:
        return [("boundary_size",), ("dist_to_com",), ("isoperimetric",),
                ("moments", (0, 1), (1, 2)), ("size",)]
    return [("dist_to_com",), ("moments", (0, 1), (1, 2)), ("size",)]


def make_item(rs, shape, radius_frac=0.3):
    """A noisy blob image and a smooth, non-distance level set around it."""
    ndim = len(shape)
    grids = np.indices(shape).astype(float)
    c = [s / 2 + rs.uniform(-2, 2) for s in shape]
    r = np.sqrt(sum((grids[a] - c[a]) ** 2 for a in range(ndim)))
    rad = radius_frac * min(shape)
    img = (r < rad).astype(float) + 0.2 * rs.randn(*shape)
    u = (rad * 0.8 - r) * (1.0 + 0.3 * np.sin(grids[0] / 3.0)) + 0.05 * rs.randn(*shape)
    return img, u


def build_engine(shape, batch, dx, specs, band=3.0, seed=0, normalize=True):
    from lsml_b200.core.engine import BandEngine
    rs = np.random.RandomState(seed)
    imgs, us = zip(*[make_item(rs, shape) for _ in range(batch)])
    imgs, us = np.stack(imgs), np.stack(us)
    eng = BandEngine(shape, batch, dx=dx, band=band, specs=specs)
    eng.load_images(imgs, normalize=normalize)
    eng.set_level_sets(us)
    return eng, imgs, us


@pytest.mark.parametrize("shape,dx", [((41, 37), (1.0, 1.0)), ((41, 37), (0.7, 1.3)),
                                      ((23, 30, 27), (1.0, 1.0, 1.0)),
                                      ((23, 30, 27), (1.2, 0.8, 1.0)), ((64,), (0.5,))])
def test_image_planes_bit_exact(oracle, shape, dx):
    specs = [("image_sample", 0), ("image_sample", 1), ("image_sample", 2.5),
             ("image_edge", 0), ("image_edge", 1), ("image_edge", 3)]
    eng, imgs, us = build_engine(shape, 2, dx, specs, normalize=False)
    dxa = np.asarray(dx)
    for p, (what, sigma) in enumerate(eng.program.planes):
        got = eng.planes[p].cpu().numpy()
        for b in range(2):
            want = (oracle.image_sample_plane(imgs[b], sigma, dxa) if what == "sample"
                    else oracle.image_edge_plane(imgs[b], sigma, dxa))
            assert np.array_equal(got[b], want), (what, sigma, np.abs(got[b] - want).max())


@pytest.mark.parametrize("shape,dx", [((5, 40, 48), (1.0, 1.0, 1.0)), ((70, 33, 64), (1.0, 0.5, 1.0)),
                                      ((3, 130), (1.0, 1.0))])
def test_image_planes_bit_exact_marched_axes(oracle, shape, dx):
    """The register-march Gaussian (strided axes, features.cu: gauss1d_march_kernel) at its
    corners: a column shorter than the filter radius (several reflections), a column longer than
    one ring length, radii 4 / 8 / 16 and the fall-back beyond 16, both marched axes of a volume."""
    specs = [("image_sample", 1), ("image_sample", 2), ("image_sample", 4), ("image_edge", 2),
             ("image_edge", 4.4)]
    eng, imgs, us = build_engine(shape, 2, dx, specs, normalize=False)
    dxa = np.asarray(dx)
    for p, (what, sigma) in enumerate(eng.program.planes):
        got = eng.planes[p].cpu().numpy()
        for b in range(2):
            want = (oracle.image_sample_plane(imgs[b], sigma, dxa) if what == "sample"
                    else oracle.image_edge_plane(imgs[b], sigma, dxa))
            assert np.array_equal(got[b], want), (what, sigma, np.abs(got[b] - want).max())


def test_normalize_matches_numpy(oracle):
    eng, imgs, us = build_engine((33, 45), 3, (1, 1), [("image_sample", 0)], normalize=True)
    got = eng.img.cpu().numpy()
    for b in range(3):
        assert np.allclose(got[b], oracle.normalize_image(imgs[b]), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("shape,dx,band", [((51, 51), (1.0, 1.0), 3.0),
                                           ((40, 56), (0.8, 1.25), 2.5),
                                           ((30, 33, 29), (1.0, 1.0, 1.0), 3.0),
                                           ((30, 33, 29), (1.0, 0.75, 1.5), 2.0)])
def test_band_rebuild_mask_bit_exact_and_distance(oracle, shape, dx, band):
    """distance_transform(u, band, dx): mask bit-exact, distances bit-exact."""
    eng, imgs, us = build_engine(shape, 3, dx, [("size",)], band=band, normalize=False)
    dist = eng.distance().cpu().numpy()
    mask = eng.mask.cpu().numpy().astype(bool)
    ctr = eng.fmm_counters()
    assert ctr["converged"], ctr
    for b in range(3):
        wd, wm = oracle.distance_transform(us[b], band, np.asarray(dx))
        assert np.array_equal(mask[b], wm), (b, (mask[b] != wm).sum(), ctr)
        assert np.array_equal(dist[b], wd), np.abs(dist[b] - wd).max()
    idx = eng.band_idx[:eng.n_band].cpu().numpy()
    assert np.array_equal(idx, np.flatnonzero(mask.ravel()))


def test_band_rebuild_reference_known_answers(oracle):
    """lsml/util/test/test_distance_transform.py:11-53 through the device path."""
    from lsml_b200.core.engine import BandEngine

    def run(arr, band, dx):
        eng = BandEngine(arr.shape, 1, dx=dx, band=band, specs=[])
        eng.set_level_sets(arr[None])
        return eng.distance().cpu().numpy()[0], eng.mask.cpu().numpy()[0].astype(bool)

    d, m = run(np.r_[-1, -1, 1, -1, -1.], 1, [1.])
    assert (d == np.r_[0., -0.5, 0.5, -0.5, 0.]).all()
    assert (m == np.r_[False, True, True, True, False]).all()
    for sign in (1.0, -1.0):
        d, m = run(sign * np.ones((4, 5, 6)), 0, [1, 1, 1.])
        assert m.sum() == 0
    arr = np.ones((4, 5, 6))
    arr[2] = -1
    d, m = run(arr, 0, [1, 1, 1.])
    assert m.sum() == arr.size
    wd, wm = oracle.distance_transform(arr, 0, [1, 1, 1.])
    assert np.array_equal(d, wd)
    d, m = run(np.zeros((4, 5, 6)), 0, [1, 1, 1.])
    assert m.sum() == 120 and (d == 0).all()


@pytest.mark.parametrize("shape,dx", [((51, 51), (1.0, 1.0)), ((40, 56), (0.8, 1.25)),
                                      ((30, 33, 29), (1.0, 1.0, 1.0)),
                                      ((30, 33, 29), (1.0, 0.75, 1.5))])
def test_feature_matrix_against_oracle(oracle, shape, dx):
    ndim = len(shape)
    specs = IMG_SPECS + shape_specs(ndim, boundary=(ndim == 2))
    eng, imgs, us = build_engine(shape, 3, dx, specs, normalize=False)
    X = eng.compute_features().cpu().numpy()
    mask = eng.mask.cpu().numpy().astype(bool)
    off = eng.item_offsets.cpu().numpy()
    dxa = np.asarray(dx)
    cols = []
    for s in specs:
        cols += [s] * oracle.spec_width(s)
    for b in range(3):
        want = oracle.feature_matrix(specs, us[b], imgs[b], None, mask[b], dxa)
        got = X[off[b]:off[b + 1]]
        assert got.shape == want.shape
        for f, s in enumerate(cols):
            exact = s[0] in ("image_sample", "image_edge", "size") or \
                (s[0] == "dist_to_com" and np.all(dxa == 1.0))
            if exact:
                assert np.array_equal(got[:, f], want[:, f]), \
                    (s, np.abs(got[:, f] - want[:, f]).max())
            else:
                assert np.allclose(got[:, f], want[:, f], rtol=RTOL, atol=1e-13), \
                    (s, np.abs(got[:, f] - want[:, f]).max())


def _train_models(X, y, seed=0):
    from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
    from sklearn.linear_model import LinearRegression
    from sklearn.neural_network import MLPRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from sklearn.tree import DecisionTreeRegressor
    rs = np.random.RandomState(seed)
    models = {
        "linear": Pipeline([("s", StandardScaler()), ("r", LinearRegression())]),
        "forest": Pipeline([("s", StandardScaler()),
                            ("r", RandomForestRegressor(n_estimators=20, max_samples=300,
                                                        random_state=rs))]),
        "forest_noscale": RandomForestRegressor(n_estimators=7, max_depth=6, random_state=1),
        "extra": ExtraTreesRegressor(n_estimators=5, random_state=2),
        "tree": DecisionTreeRegressor(max_depth=1, random_state=0),
        "mlp": Pipeline([("s", StandardScaler()),
                         ("r", MLPRegressor(hidden_layer_sizes=(16, 8), max_iter=50,
                                            random_state=0))]),
    }
    for m in models.values():
        m.fit(X, y)
    return models


@pytest.mark.parametrize("variant", ["ranked", "global"])
def test_regressors_against_sklearn(variant, monkeypatch):
    """``variant``: which forest kernel the exported model walks -- rank-quantised 4-byte nodes
    from shared memory (the default whenever the forest fits that format) or 32-byte super
    nodes from global memory; ExtraTrees with full-depth trees of 4000 samples do not fit the
    ranked format and take the super-node path in both runs."""
    import torch
    from lsml_b200.core.regressors import export_regressor
    monkeypatch.setenv("LSML_FOREST_VARIANT", variant)
    rs = np.random.RandomState(5)
    X = rs.randn(4000, 12) * rs.rand(12) * 5 + rs.randn(12)
    y = np.sin(X[:, 0]) + 0.5 * X[:, 3] - (X[:, 5] > 0.3) + 0.1 * rs.randn(4000)
    Xt = rs.randn(5000, 12) * rs.rand(12) * 5 + rs.randn(12)
    Xt[:300] = X[:300]                      # training points sit exactly on split values
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        models = _train_models(X, y)
    Xd = torch.from_numpy(Xt).cuda()
    for name, m in models.items():
        want = m.predict(Xt)
        reg = export_regressor(m)
        got = reg.predict(Xd).cpu().numpy()
        if name in ("forest", "forest_noscale", "extra", "tree"):
            assert np.array_equal(got, want), (name, np.abs(got - want).max())
            if name in ("forest", "forest_noscale", "tree"):
                assert (reg._ranked_layout() is not None) == (variant == "ranked"), name
        else:
            assert np.allclose(got, want, rtol=1e-11, atol=1e-12), \
                (name, np.abs(got - want).max())


def test_mlp_tensor_core_kernel_against_sklearn():
    """The FP64 tensor-core (DMMA) MLP kernel: widths that are not multiples of 4 / 8, two hidden
    layers, every activation, a batch that is not a multiple of the 8-voxel tile."""
    import torch
    from sklearn.neural_network import MLPRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from lsml_b200.core.regressors import export_regressor
    rs = np.random.RandomState(3)
    for F, hidden, act in ((16, (100,), "relu"), (13, (37, 9), "tanh"), (5, (10,), "logistic"),
                           (22, (64, 33, 7), "identity")):
        X = rs.randn(600, F) * (1 + 3 * rs.rand(F)) + rs.randn(F)
        y = np.tanh(X[:, 0]) + 0.3 * X[:, F - 1] + 0.05 * rs.randn(600)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = Pipeline([("s", StandardScaler()),
                          ("r", MLPRegressor(hidden_layer_sizes=hidden, activation=act, max_iter=30,
                                             random_state=0))]).fit(X, y)
        Xt = rs.randn(1003, F) * 2
        got = export_regressor(m).predict(torch.from_numpy(Xt).cuda()).cpu().numpy()
        want = m.predict(Xt)
        assert np.allclose(got, want, rtol=1e-11, atol=1e-12), (F, hidden, act, np.abs(got - want).max())


def test_ranked_forest_shapes():
    """The shared-memory forest kernel at its corners: more trees than one staged group holds,
    64 features (one CTA per SM), a batch that is not a multiple of the chunk, NaN features."""
    import torch
    from sklearn.ensemble import RandomForestRegressor
    from lsml_b200.core.regressors import export_regressor
    rs = np.random.RandomState(11)
    for F, n_trees, n in ((5, 150, 2051), (64, 30, 1500), (33, 40, 513)):
        X = rs.randn(1500, F)
        y = X[:, 0] - 2 * X[:, F - 1] + np.cos(X[:, F // 2]) + 0.1 * rs.randn(1500)
        m = RandomForestRegressor(n_estimators=n_trees, max_samples=400, random_state=0).fit(X, y)
        reg = export_regressor(m)
        assert reg._ranked_layout() is not None
        Xt = rs.randn(n, F)
        Xt[:200] = X[:200]
        got = reg.predict(torch.from_numpy(Xt).cuda()).cpu().numpy()
        assert np.array_equal(got, m.predict(Xt)), (F, n_trees)
    Xn = np.full((3, 33), np.nan)
    got = reg.predict(torch.from_numpy(Xn).cuda()).cpu().numpy()
    want = 0.0
    for e in m.estimators_:
        t, i = e.tree_, 0
        while t.children_left[i] != -1:
            i = t.children_right[i]
        want += t.value[i, 0, 0]
    assert np.array_equal(got, np.full(3, want / len(m.estimators_)))


@pytest.mark.parametrize("shape,dx", [((51, 51), (1.0, 1.0)), ((30, 33, 29), (1.0, 0.75, 1.5))])
def test_full_step_against_oracle(oracle, shape, dx):
    """features -> forest velocity -> upwind update -> band rebuild, two iterations."""
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    ndim = len(shape)
    specs = IMG_SPECS + shape_specs(ndim, boundary=False)
    eng, imgs, us = build_engine(shape, 2, dx, specs, normalize=False, seed=3)
    dxa = np.asarray(dx)
    mask0 = eng.mask.cpu().numpy().astype(bool)
    # train on the oracle's features of item 0 with a synthetic target
    F0 = oracle.feature_matrix(specs, us[0], imgs[0], None, mask0[0], dxa)
    rs = np.random.RandomState(0)
    y = F0[:, 0] - F0[:, 1] + 0.1 * rs.randn(F0.shape[0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        reg = Pipeline([("s", StandardScaler()),
                        ("r", RandomForestRegressor(n_estimators=10, max_samples=200,
                                                    random_state=rs))]).fit(F0, y)
    step = 0.4
    u_o = [us[b].copy() for b in range(2)]
    m_o = [mask0[b] for b in range(2)]
    for it in range(2):
        eng.step(reg, step)
        u_g = eng.u.cpu().numpy()
        m_g = eng.mask.cpu().numpy().astype(bool)
        for b in range(2):
            u_o[b], _, m_o[b], vb = oracle.evolve_step(u_o[b], imgs[b], None, m_o[b], dxa, specs,
                                                       reg, step, eng.band)
            assert np.array_equal(m_g[b], m_o[b]), (it, b, (m_g[b] != m_o[b]).sum())
            assert np.allclose(u_g[b], u_o[b], rtol=1e-12, atol=1e-12), \
                np.abs(u_g[b] - u_o[b]).max()


@pytest.mark.parametrize("shape,dx", [((61, 47), (1.0, 1.0)), ((26, 31, 28), (1.0, 1.0, 1.0)),
                                      ((64, 72, 80), (1.0, 0.9, 1.1))])
def test_incremental_statistics_and_sparse_rebuild_match_dense(shape, dx):
    """After several iterations the {u>0} statistics maintained from sign changes in the scatter
    kernel equal a dense recount, and the band found from the previous band (sparse interface
    search, hinted schedule) equals the band of a from-scratch dense rebuild of the same u."""
    from sklearn.linear_model import LinearRegression
    specs = [("image_sample", 0.0), ("size",)]
    eng, imgs, us = build_engine(shape, 3, dx, specs, normalize=True, seed=5)
    reg = LinearRegression().fit(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]),
                                 np.array([0.3, 1.3, 0.3]))          # v = img + 0.3
    for it in range(6):
        eng.step(reg, 0.45)
        stats_inc = eng.stats.cpu().numpy().copy()
        mask_inc = eng.mask.cpu().numpy().copy()
        idx_inc = eng.band_idx[:eng.n_band].cpu().numpy().copy()
        u_now = eng.u.cpu().numpy()
        want = np.zeros_like(stats_inc)
        for b in range(3):
            pos = u_now[b] > 0
            coords = np.nonzero(pos)
            want[b, 0] = pos.sum()
            for a in range(len(shape)):
                want[b, 1 + 3 - len(shape) + a] = coords[a].sum()
                want[b, 4 + 3 - len(shape) + a] = (coords[a].astype(np.int64) ** 2).sum()
            want[b, 7] = (u_now[b] == 0).sum()
        assert np.array_equal(stats_inc, want), (it, stats_inc, want)
        # dense rebuild of the same u on a fresh engine
        from lsml_b200.core.engine import BandEngine
        ref = BandEngine(shape, 3, dx=dx, band=eng.band, specs=[])
        ref.set_level_sets(u_now)
        assert np.array_equal(ref.mask.cpu().numpy(), mask_inc), it
        assert np.array_equal(ref.band_idx[:ref.n_band].cpu().numpy(), idx_inc), it


def test_vanishing_level_set_makes_later_steps_noops(oracle):
    """`if mask.any()` (model.py:381) without the host: when an update pushes the whole level set
    below zero the rebuilt band is empty (distance_transform.py:42-47) and every later step is a
    no-op -- the kernels read the band size from the device and find nothing to do."""
    from sklearn.linear_model import LinearRegression
    from lsml_b200.core.engine import BandEngine
    shape = (24, 26)
    g = np.indices(shape).astype(float)
    u0 = 3.0 - np.sqrt((g[0] - 12) ** 2 + (g[1] - 13) ** 2)          # a small disc
    img = np.random.RandomState(0).randn(*shape)
    specs = [("image_sample", 0.0), ("size",)]
    reg = LinearRegression().fit(np.random.RandomState(1).randn(8, 2), np.full(8, -30.0))   # v = -30
    eng = BandEngine(shape, 1, dx=np.ones(2), band=3.0, specs=specs)
    eng.load_images(img[None], normalize=False)
    eng.set_level_sets(u0[None])
    assert eng.n_band > 0
    eng.step(reg, 1.0)
    u1 = eng.u[0].cpu().numpy()
    assert (u1 <= 0).all() or (u1 > 0).sum() == 0
    assert eng.n_band == 0 and int(eng.mask.sum()) == 0
    dist, mask = oracle.distance_transform(u1, 3.0, np.ones(2))
    assert not mask.any()
    for _ in range(3):                      # no-ops: u unchanged, band stays empty, units unchanged
        units = eng.band_voxel_iterations
        eng.step(reg, 1.0)
        assert eng.n_band == 0
        assert eng.band_voxel_iterations == units
        assert np.array_equal(eng.u[0].cpu().numpy(), u1)
    eng.check_converged()
