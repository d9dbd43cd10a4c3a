"""GPU: the shape features added in round 2 -- 3-D BoundarySize / IsoperimetricRatio (marching
cubes area as a reduction, csrc/mcubes.cu) and Moments of any order (exact per-axis histograms) --
against the oracle.  Reductions: relative 1e-12 (summation order differs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _items(rs, shape, batch):
    g = np.indices(shape).astype(float)
    us, imgs = [], []
    for _ in range(batch):
        c = [s / 2 + rs.uniform(-2, 2) for s in shape]
        r = np.sqrt(sum((g[a] - c[a]) ** 2 for a in range(len(shape))))
        us.append(0.3 * min(shape) - r + 0.6 * rs.randn(*shape))
        imgs.append((r < 0.3 * min(shape)).astype(float) + 0.2 * rs.randn(*shape))
    return np.stack(us), np.stack(imgs)


@pytest.mark.parametrize("dx", [(1.0, 1.0, 1.0), (0.7, 1.3, 2.0)])
def test_basic_shape_features_3d(oracle, dx):
    """get_basic_shape_features(ndim=3) (shape.py:270-297): boundary size, distance to the centre
    of mass, sphericity, moments, volume -- the list that could not run in round 1."""
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.feature import FeatureMap, get_basic_shape_features
    rs = np.random.RandomState(5)
    shape, batch = (22, 27, 19), 3
    us, imgs = _items(rs, shape, batch)
    specs = FeatureMap(get_basic_shape_features(ndim=3)).specs()
    assert [s[0] for s in specs] == ["boundary_size", "dist_to_com", "isoperimetric", "moments", "size"]
    eng = BandEngine(shape, batch, dx=np.asarray(dx), band=3.0, specs=specs)
    eng.load_images(imgs, normalize=False)
    eng.set_level_sets(us)
    X = eng.compute_features().cpu().numpy()
    off = eng.item_offsets.cpu().numpy()
    masks = eng.mask.cpu().numpy().astype(bool)
    for b in range(batch):
        dist, mask = oracle.distance_transform(us[b], 3.0, np.asarray(dx))
        assert np.array_equal(mask, masks[b])
        want = oracle.feature_matrix(specs, us[b], imgs[b], dist, mask, np.asarray(dx))
        got = X[off[b]:off[b + 1]]
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()
        area = oracle.boundary_size(us[b], np.asarray(dx))
        assert abs(eng.boundary[b].item() - area) <= 1e-12 * area


@pytest.mark.parametrize("shape,axes", [((37, 45), (0, 1)), ((20, 24, 28), (0, 1, 2)),
                                        ((20, 24, 28), (2, 0))])
def test_moments_of_any_order(oracle, shape, axes):
    from lsml_b200.core.engine import BandEngine
    rs = np.random.RandomState(9)
    nd = len(shape)
    dx = np.array([0.8, 1.0, 1.25][:nd])
    us, imgs = _items(rs, shape, 2)
    specs = [("moments", axes, (1, 2, 3, 4, 5)), ("size",)]
    eng = BandEngine(shape, 2, dx=dx, band=3.0, specs=specs)
    eng.load_images(imgs, normalize=False)
    eng.set_level_sets(us)
    X = eng.compute_features().cpu().numpy()
    off = eng.item_offsets.cpu().numpy()
    for b in range(2):
        dist, mask = oracle.distance_transform(us[b], 3.0, dx)
        want = oracle.feature_matrix(specs, us[b], imgs[b], dist, mask, dx)
        got = X[off[b]:off[b + 1]]
        assert np.allclose(got, want, rtol=1e-11, atol=1e-9), np.abs(got - want).max()


def test_evolution_with_3d_shape_features(oracle):
    """A few evolution steps whose velocity depends on the surface area and a 3rd-order moment
    (fused velocity kernel reading the per-item scalars), against the oracle."""
    from sklearn.linear_model import LinearRegression
    from lsml_b200.core.engine import BandEngine
    rs = np.random.RandomState(2)
    shape = (24, 26, 22)
    dx = np.ones(3)
    us, imgs = _items(rs, shape, 2)
    specs = [("image_sample", 1.0), ("boundary_size",), ("isoperimetric",),
             ("moments", (0, 2), (2, 3)), ("size",)]
    eng = BandEngine(shape, 2, dx=dx, band=3.0, specs=specs)
    eng.load_images(imgs, normalize=True)
    eng.set_level_sets(us)
    X = eng.compute_features().cpu().numpy()
    reg = LinearRegression().fit(X, 0.5 * X[:, 0] + 1e-4 * X[:, 1] + 0.2 * X[:, 2] + 1e-3 * X[:, 4])
    for _ in range(3):
        eng.step(reg, 0.3)
    for b in range(2):
        img_n = oracle.normalize_image(imgs[b])
        u = us[b]
        dist, mask = oracle.distance_transform(u, 3.0, dx)
        for _ in range(3):
            u, dist, mask, _ = oracle.evolve_step(u, img_n, dist, mask, dx, specs, reg, 0.3, 3.0)
        assert np.array_equal(eng.mask[b].cpu().numpy().astype(bool), mask)
        assert np.allclose(eng.u[b].cpu().numpy(), u, rtol=1e-9, atol=1e-9)
