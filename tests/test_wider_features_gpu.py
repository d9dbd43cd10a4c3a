"""GPU: features added when widening past the first hot-path slice.

COMRaySample (lsml/feature/provided/image.py:124-171): the CUDA path against the golden vectors
of the unmodified reference (tests/golden/com_ray.npz) and against the oracle on a batch.  The
same per-sample function is checked on the host in tests/test_com_ray_host.py."""
import numpy as np
import pytest

from golden_util import load

pytestmark = pytest.mark.gpu


def cases():
    g = load("com_ray.npz")
    for c in range(int(g["n_cases"])):
        yield (g["u%d" % c], g["img%d" % c], g["mask%d" % c], g["dx%d" % c],
               int(g["n_samples%d" % c]), g["X%d" % c])


def test_com_ray_reproduces_reference_bit_exact():
    from lsml_b200.core.engine import BandEngine
    for u, img, mask, dx, n_samples, want in cases():
        eng = BandEngine(u.shape, 1, dx=dx, band=3.0, specs=[("com_ray", n_samples)])
        eng.load_images(img[None], normalize=False)
        eng.set_level_sets(u[None])
        assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask)
        X = eng.compute_features().cpu().numpy()
        assert X.shape == want.shape
        assert np.array_equal(X, want)


def test_com_ray_feature_class_and_column_layout():
    """Through the public feature classes, mixed with other columns (feature_map.py:42-51)."""
    from lsml_b200.feature import COMRaySample, FeatureMap, ImageSample, Size
    u, img, mask, dx, n_samples, want = next(cases())
    ray = COMRaySample(ndim=2, sigma=0, n_samples=n_samples)
    assert ray.size == 2 * n_samples and ray.name == "COM Ray (N=3, σ = 0.000)"
    dense = ray(u=u, img=img, dist=None, mask=mask, dx=dx)
    assert dense.shape == u.shape + (2 * n_samples,)
    assert np.array_equal(dense[mask], want) and not dense[~mask].any()
    fm = FeatureMap([ImageSample(ndim=2, sigma=0), ray, Size(ndim=2)])
    assert fm.n_features == 2 * n_samples + 2
    F = fm(u=u, img=img, dist=None, mask=mask, dx=dx)
    assert np.array_equal(F[mask][:, 0], img[mask])
    assert np.array_equal(F[mask][:, 1:1 + 2 * n_samples], want)
    assert np.array_equal(F[mask][:, -1], np.full(int(mask.sum()), (u > 0).sum() * np.prod(dx)))


def test_com_ray_batch_against_oracle(oracle):
    from lsml_b200.core.engine import BandEngine
    rs = np.random.RandomState(5)
    shape, batch, n_samples = (20, 24, 22), 3, 5
    dx = np.array([1.0, 1.0, 1.0])
    gg = np.indices(shape).astype(float)
    us, imgs = [], []
    for b in range(batch):
        c = np.array(shape) / 2.0 + rs.uniform(-4, 4, size=3)
        r = np.sqrt(sum((gg[a] - c[a]) ** 2 for a in range(3)))
        us.append(rs.uniform(4, 7) - r + 0.2 * rs.randn(*shape))
        imgs.append(rs.randn(*shape))
    us, imgs = np.array(us), np.array(imgs)
    specs = [("image_sample", 1.0), ("com_ray", n_samples), ("dist_to_com",)]
    eng = BandEngine(shape, batch, dx=dx, band=3.0, specs=specs)
    eng.load_images(imgs, normalize=False)
    eng.set_level_sets(us)
    X = eng.compute_features().cpu().numpy()
    off = eng.item_offsets.cpu().numpy()
    for b in range(batch):
        dist, mask = oracle.distance_transform(us[b], 3.0, dx)
        assert np.array_equal(eng.mask[b].cpu().numpy().astype(bool), mask)
        want = oracle.feature_matrix(specs, us[b], imgs[b], dist, mask, dx)
        got = X[off[b]:off[b + 1]]
        assert np.array_equal(got[:, 1:1 + 2 * n_samples], want[:, 1:1 + 2 * n_samples])
        assert np.array_equal(got[:, 0], want[:, 0]) and np.array_equal(got[:, -1], want[:, -1])


def test_com_ray_rejected_where_it_cannot_run():
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.slab import SlabVolume
    with pytest.raises(ValueError):
        BandEngine((32,), 1, specs=[("com_ray", 2)])
    with pytest.raises(NotImplementedError):
        SlabVolume((32, 16, 16), world=2, specs=[("com_ray", 2)])


# ---- user-feature registry: PrevIterFeature (examples/gestalt2d) and precomputed image planes ----
def test_smooth_u_reproduces_reference_example_bit_exact():
    from lsml_b200.core.engine import BandEngine
    g = load("prev_iter.npz")
    for c in range(int(g["n_cases"])):
        u, mask, dx = g["u%d" % c], g["mask%d" % c], g["dx%d" % c]
        specs = [("smooth_u", float(s)) for s in g["sigmas%d" % c]]
        eng = BandEngine(u.shape, 1, dx=dx, band=3.0, specs=specs)
        eng.load_images(np.zeros(u.shape)[None], normalize=False)
        eng.set_level_sets(u[None])
        assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask)
        assert np.array_equal(eng.compute_features().cpu().numpy(), g["X%d" % c])


def test_user_features_through_an_evolution(oracle):
    """smooth_u is refreshed from u every iteration; a precomputed user plane is sampled like an
    image plane: two evolution steps against the oracle."""
    from scipy.ndimage import gaussian_filter
    from sklearn.linear_model import LinearRegression
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.feature import FeatureMap, ImageSample, PrecomputedImageFeature, PrevIterFeature

    class Blurred(PrecomputedImageFeature):
        @property
        def name(self):
            return "blurred-squared-{:.2f}".format(self.sigma)

        def compute_plane(self, img, dx):
            return gaussian_filter(img, sigma=self.sigma) ** 2

    rs = np.random.RandomState(9)
    shape, dx = (40, 44), np.ones(2)
    gg = np.indices(shape).astype(float)
    u = 9.0 - np.sqrt((gg[0] - 19.3) ** 2 + (gg[1] - 22.6) ** 2) + 0.2 * rs.randn(*shape)
    img = rs.randn(*shape)
    fm = FeatureMap([ImageSample(2, 0), PrevIterFeature(2, 1.0), Blurred(2, 1.5), PrevIterFeature(2, 2.0)])
    specs = fm.specs()
    reg = LinearRegression().fit(rs.randn(32, 4), rs.randn(32))
    reg.coef_ = np.array([0.3, 0.05, -0.2, 0.02]); reg.intercept_ = 0.1
    eng = BandEngine(shape, 1, dx=dx, band=3.0, specs=specs)
    eng.load_images(img[None], normalize=False)
    eng.set_level_sets(u[None])
    dist, mask = oracle.distance_transform(u, 3.0, dx)
    for it in range(2):
        X = eng.compute_features().cpu().numpy()
        want = oracle.feature_matrix(specs, u, img, dist, mask, dx)
        if it == 0:
            assert np.array_equal(X, want)
        else:       # u itself now differs in the last bits (dot-product order of the regressor)
            assert np.allclose(X, want, rtol=1e-10, atol=1e-12), np.abs(X - want).max()
        eng.step(reg, 0.4)
        u, dist, mask, _ = oracle.evolve_step(u, img, dist, mask, dx, specs, reg, 0.4, 3.0)
        assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask)
        assert np.allclose(eng.u[0].cpu().numpy(), u, rtol=1e-12, atol=1e-12)
