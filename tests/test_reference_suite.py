"""The reference's own unit tests, restated against the drop-in package.

Every test below follows one test of the reference (file:line in the docstring): same inputs,
same calls, same assertions, with ``lsml.*`` replaced by a ``pkg`` fixture.  ``pkg`` is

* ``reference`` -- the UNMODIFIED reference imported through oracle/ref_shim.py (runs on the CPU,
  only where /root/reference exists): proves that the restated assertions are the reference's;
* ``b200``      -- ``lsml_b200`` (marked ``gpu``): the drop-in has to pass the same tests.

Tests that never reach a kernel (validation, equality, balance_mask, seeds) run against
``lsml_b200`` on the CPU as well.  Left out, with the reason: ``test_datasets_handler.py`` (HDF5
store, out of scope), the 3-D
``BoundarySize`` / ``IsoperimetricRatio`` tests (marching cubes: the b200 leg only,
asserted as such).  The 3-D grids of test_shape.py are subsampled by 2 per axis on the CPU
(reference) leg to keep it to seconds; the GPU leg uses the reference's sizes.
"""
import types

import numpy as np
import pytest


# --------------------------------------------------------------------------------------------
# the two packages under test
# --------------------------------------------------------------------------------------------
def _reference_pkg():
    from oracle import oracle as O
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference is not present on this machine")
    O.build()
    import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.feature import base_feature, feature_map
    from lsml.feature.provided import image, shape
    from lsml.gradient import masked_gradient
    from lsml.initializer.initializer_base import InitializerBase
    from lsml.initializer.provided.ball import BallInitializer, RandomBallInitializer
    from lsml.initializer.seed import center_of_mass_seeder
    from lsml.util.balance_mask import balance_mask
    from lsml.util.distance_transform import distance_transform
    return types.SimpleNamespace(
        name="reference", base_feature=base_feature, FeatureMap=feature_map.FeatureMap,
        image=image, shape=shape, mg=masked_gradient, InitializerBase=InitializerBase,
        BallInitializer=BallInitializer, RandomBallInitializer=RandomBallInitializer,
        center_of_mass_seeder=center_of_mass_seeder, balance_mask=balance_mask,
        distance_transform=distance_transform, sub3d=2, has_contours=False)


def _b200_pkg():
    from lsml_b200.feature import base_feature, feature_map
    from lsml_b200.feature.provided import image, shape
    from lsml_b200.gradient import masked_gradient
    from lsml_b200.initializer import (BallInitializer, InitializerBase, RandomBallInitializer,
                                       center_of_mass_seeder)
    from lsml_b200.util.balance_mask import balance_mask
    from lsml_b200.util.distance_transform import distance_transform
    return types.SimpleNamespace(
        name="b200", base_feature=base_feature, FeatureMap=feature_map.FeatureMap,
        image=image, shape=shape, mg=masked_gradient, InitializerBase=InitializerBase,
        BallInitializer=BallInitializer, RandomBallInitializer=RandomBallInitializer,
        center_of_mass_seeder=center_of_mass_seeder, balance_mask=balance_mask,
        distance_transform=distance_transform, sub3d=1, has_contours=True)


@pytest.fixture(scope="module", params=["reference", pytest.param("b200", marks=pytest.mark.gpu)])
def pkg(request):
    """Tests that run kernels: the reference on the CPU, the drop-in on the GPU."""
    return _reference_pkg() if request.param == "reference" else _b200_pkg()


@pytest.fixture(scope="module", params=["reference", "b200"])
def host_pkg(request):
    """Tests that never reach a kernel: both packages on the CPU."""
    return _reference_pkg() if request.param == "reference" else _b200_pkg()


def almost(a, b, places=7):
    """unittest.assertAlmostEqual"""
    assert round(abs(a - b), places) == 0, (a, b, places)


# --------------------------------------------------------------------------------------------
# lsml/feature/test/test_base_feature.py
# --------------------------------------------------------------------------------------------
def _mock_features(p):
    class SomeImageFeature(p.base_feature.BaseImageFeature):
        name = None
        locality = None

        def compute_feature(self, u, img, dist, mask, dx):
            pass

    class SomeShapeFeature(p.base_feature.BaseShapeFeature):
        name = None
        locality = None

        def compute_feature(self, u, dist, mask, dx):
            pass

    return (lambda ndim: SomeImageFeature(ndim=ndim)), (lambda ndim: SomeShapeFeature(ndim=ndim))


def test_base_feature_correct_type(host_pkg):
    """test_base_feature.py:34-43"""
    image_feature, shape_feature = _mock_features(host_pkg)
    bf = host_pkg.base_feature
    assert image_feature(1).type == bf.IMAGE_FEATURE_TYPE
    assert shape_feature(1).type == bf.SHAPE_FEATURE_TYPE


def test_base_feature_shape_and_dx_mismatches(host_pkg):
    """test_base_feature.py:45-115"""
    image_feature, shape_feature = _mock_features(host_pkg)
    image_feature, shape_feature = image_feature(2), shape_feature(2)
    u, img = np.ones((3, 2)), np.ones((3, 2))
    mask = np.ones((3, 2), dtype=bool)
    for dist in (np.ones((3, 1)), np.ones((3,))):           # :45-59, :82-96
        with pytest.raises(ValueError):
            image_feature(u=u, img=img, dist=dist, mask=mask)
        with pytest.raises(ValueError):
            shape_feature(u=u, dist=dist, mask=mask)
    with pytest.raises(ValueError):                          # :61-69
        image_feature(u=u, img=np.ones((3, 1)))
    with pytest.raises(ValueError):                          # :71-80
        image_feature(u=u, img=np.ones((3, 1)), mask=np.ones((3, 9), dtype=bool))
    with pytest.raises(ValueError):                          # :98-115: ndim = 2, one dx term
        image_feature(u=u, img=img, dist=np.ones((3, 2)), mask=mask, dx=[1])
    with pytest.raises(ValueError):
        shape_feature(u=u, dist=np.ones((3, 2)), mask=mask, dx=[1])


def test_base_feature_eq_neq_hash(host_pkg):
    """test_base_feature.py:117-154"""
    image_feature, shape_feature = _mock_features(host_pkg)
    assert image_feature(2) == image_feature(2)
    assert shape_feature(2) == shape_feature(2)
    i1, i2, s1, s2 = image_feature(2), image_feature(2), shape_feature(2), shape_feature(2)
    assert len({i1, i2}) == 1
    i1.name = 'other image feature'
    s1.name = 'other shape feature'
    assert i1 != i2 and s1 != s2
    assert len({s1, s2}) == 2


def test_base_feature_empty_mask_and_defaults(host_pkg):
    """test_base_feature.py:156-183"""
    image_feature, shape_feature = _mock_features(host_pkg)
    image_feature, shape_feature = image_feature(2), shape_feature(2)
    u, img, dist = np.ones((3, 2)), np.ones((3, 2)), np.ones((3, 2))
    mask = np.zeros((3, 2), dtype=bool)
    feature = image_feature(u=u, img=img, dist=dist, mask=mask)
    assert feature.shape == u.shape + (image_feature.size,) and feature.dtype == u.dtype
    feature = shape_feature(u=u, dist=dist, mask=mask)
    assert feature.shape == u.shape + (image_feature.size,) and feature.dtype == u.dtype
    image_feature(u=u, img=img)          # no mask, no dist
    shape_feature(u=u)


def test_base_feature_type_errors(host_pkg):
    """test_base_feature.py:185-228"""
    image_feature, shape_feature = _mock_features(host_pkg)
    image_feature, shape_feature = image_feature(2), shape_feature(2)
    u = np.ones((3, 2))
    with pytest.raises(TypeError):
        image_feature(u='not an array')
    with pytest.raises(TypeError):
        shape_feature(u='not an array')
    with pytest.raises(ValueError):
        image_feature(u=u)                                   # img missing
    with pytest.raises(TypeError):
        image_feature(u=u, img='not an array')
    with pytest.raises(TypeError):
        shape_feature(u=u, dist='not an array')
    with pytest.raises(TypeError):
        shape_feature(u=u, mask=u)                           # mask not bool


# --------------------------------------------------------------------------------------------
# lsml/util/test/test_balance_mask.py, lsml/initializer/test/test_seed.py
# --------------------------------------------------------------------------------------------
def test_balance_mask(host_pkg):
    """test_balance_mask.py:10-43 (+ what its first two tests meant to assert: equal counts)"""
    balance_mask = host_pkg.balance_mask
    random_state = np.random.RandomState(1234)
    arr = random_state.randn(100)
    mask = balance_mask(arr, random_state)
    assert len(arr[mask] > 0) == len(arr[mask] < 0)
    assert (arr[mask] > 0).sum() == (arr[mask] < 0).sum()
    random_state = np.random.RandomState(1234)
    arr = random_state.randn(100)
    arr[random_state.rand(arr.shape[0]) < 0.25] = 0.
    mask = balance_mask(arr, random_state)
    assert (arr[mask] > 0).sum() == (arr[mask] < 0).sum()
    assert mask[arr == 0].all()                              # zeros are always kept (:31-39)
    for arr in (np.ones(100), -np.ones(100)):
        assert balance_mask(arr, np.random.RandomState(1234)).sum() == arr.shape[0]


def _disk(r):
    g = np.arange(-r, r + 1)
    return (g[:, None] ** 2 + g[None, :] ** 2 <= r * r).astype(np.uint8)     # skimage.morphology.disk


def _ball(r):
    g = np.arange(-r, r + 1)
    return (g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2 <= r * r).astype(np.uint8)


def test_center_of_mass_seeder(host_pkg):
    """test_seed.py:12-135"""
    seeder = host_pkg.center_of_mass_seeder

    def seed_of(seg, dx):
        return seeder(dataset_example=types.SimpleNamespace(seg=seg, dx=np.asarray(dx, dtype=float),
                                                            index=None, key=None, img=None, dist=None))

    seg = np.pad(_disk(10), 10, 'constant')
    assert np.linalg.norm(seed_of(seg, [1., 1.]) - np.r_[20, 20]) < 1e-8
    assert np.linalg.norm(seed_of(seg, [1., 0.5]) - np.r_[20, 10]) < 1e-8
    assert np.linalg.norm(seed_of(seg[:, ::2], [1., 2.]) - np.r_[20, 20]) < 1e-8
    translation = np.r_[5, -3]
    moved = np.roll(np.roll(seg, translation[0], axis=0), translation[1], axis=1)
    assert np.linalg.norm(seed_of(moved, [1., 1.]) - (np.r_[20, 20] + translation)) < 1e-8
    dx = np.r_[1, 0.5]
    assert np.linalg.norm(seed_of(moved, dx) - (np.r_[20, 10] + translation * dx)) < 1e-8
    seg3 = np.pad(_ball(10), 10, 'constant')
    assert np.linalg.norm(seed_of(seg3, [1., 1., 1.]) - np.r_[20, 20, 20]) < 1e-8
    assert np.linalg.norm(seed_of(seg3[:, :, ::2], [1., 1., 2.]) - np.r_[20, 20, 20]) < 1e-8
    translation = np.r_[5, -3, 4.]
    moved = seg3
    for itrans, trans in enumerate(translation):
        moved = np.roll(moved, int(trans), axis=itrans)
    assert np.linalg.norm(seed_of(moved, [1., 1., 1.]) - (np.r_[20, 20, 20] + translation)) < 1e-8


# --------------------------------------------------------------------------------------------
# lsml/feature/provided/test/test_image.py
# --------------------------------------------------------------------------------------------
def _image_case():
    random_state = np.random.RandomState(123)
    img = random_state.randn(100, 200)
    u = random_state.randn(*img.shape)
    return img, u, u > 0


def test_image_sample2d(pkg):
    """test_image.py:10-43"""
    from scipy.ndimage import gaussian_filter1d
    img, u, mask = _image_case()
    feature = pkg.image.ImageSample(ndim=2, sigma=0)(u=u, img=img, mask=mask, dx=None)
    assert np.abs(feature[mask] - img[mask]).mean() <= 1e-8
    sigma = 1.0
    feature = pkg.image.ImageSample(ndim=2, sigma=sigma)(u=u, img=img, mask=mask, dx=[1, 2])
    smoothed_img = gaussian_filter1d(gaussian_filter1d(img, sigma, axis=0), 0.5 * sigma, axis=1)
    assert np.abs(feature[mask] - smoothed_img[mask]).mean() <= 1e-8


def test_image_edge_sample2d(pkg):
    """test_image.py:45-81"""
    from scipy.ndimage import gaussian_filter1d
    img, u, mask = _image_case()
    di, dj = np.gradient(img)
    gmag = np.sqrt(di ** 2 + dj ** 2)
    feature = pkg.image.ImageEdgeSample(ndim=2, sigma=0)(u=u, img=img, mask=mask, dx=None)
    assert np.abs(feature[mask] - gmag[mask]).mean() <= 1e-8
    sigma = 1
    feature = pkg.image.ImageEdgeSample(ndim=2, sigma=sigma)(u=u, img=img, mask=mask, dx=[1, 2])
    di = gaussian_filter1d(img, sigma=sigma, axis=0, order=1)
    dj = gaussian_filter1d(img, sigma=0.5 * sigma, axis=1, order=1)
    gmag = np.sqrt(di ** 2 + dj ** 2)
    assert np.abs(feature[mask] - gmag[mask]).mean() <= 1e-8


def test_interior_image_average_and_variation_2d(pkg):
    """test_image.py:83-157 (statistics over the MASK, despite the names)"""
    from scipy.ndimage import gaussian_filter1d
    img, u, mask = _image_case()
    flat = img.copy()
    flat[mask] = 2.0
    almost(2.0, pkg.image.InteriorImageAverage(ndim=2, sigma=0)(u=u, img=flat, mask=mask)[mask][0])
    almost(0, pkg.image.InteriorImageVariation(ndim=2, sigma=0)(u=u, img=flat, mask=mask)[mask][0])
    sigma = 10
    smoothed_image = gaussian_filter1d(gaussian_filter1d(img, sigma=sigma, axis=0),
                                       sigma=0.5 * sigma, axis=1)
    feature = pkg.image.InteriorImageAverage(ndim=2, sigma=sigma)(u=u, img=img, mask=mask, dx=[1, 2])
    almost(smoothed_image[mask].mean(), feature[mask][0])
    feature = pkg.image.InteriorImageVariation(ndim=2, sigma=sigma)(u=u, img=img, mask=mask, dx=[1, 2])
    almost(smoothed_image[mask].std(), feature[mask][0])


def test_com_ray_samples2d(pkg):
    """test_image.py:159-189: samples of a signed distance field along the ray are a line of
    slope 1."""
    x, dx = np.linspace(-2, 2, 301, retstep=True)
    y, dy = np.linspace(-2, 2, 201, retstep=True)
    xx, yy = np.meshgrid(x, y)
    img = 1 - np.sqrt(xx ** 2 + yy ** 2)
    mask = abs(img) < 0.01
    ii, jj = np.where(mask)
    mask[ii[1:], jj[1:]] = False
    com_ray = pkg.image.COMRaySample(sigma=0, n_samples=10)
    features = com_ray(u=img, img=img, mask=mask, dx=[dy, dx])
    samples = features[mask][0]
    _, dt = np.linspace(-1, 1, 2 * com_ray.n_samples + 2, retstep=True)
    ds = np.diff(samples)
    almost(ds.var(), 0, places=10)
    almost((ds / dt).mean(), 1, places=1)


# --------------------------------------------------------------------------------------------
# lsml/feature/provided/test/test_shape.py
# --------------------------------------------------------------------------------------------
def _grid(ns, sub):
    """np.linspace(-2, 2, n, retstep=True) per axis (every `sub`-th point on the CPU leg) and the
    meshgrid in the reference's (default 'xy') indexing."""
    axes, steps = [], []
    for n in ns:
        a, d = np.linspace(-2, 2, (n - 1) // sub + 1, retstep=True)
        axes.append(a)
        steps.append(d)
    return np.meshgrid(*axes), steps


def _first_voxel_mask(shape_):
    mask = np.zeros(shape_, dtype=bool)
    mask.ravel()[0] = True
    return mask


def test_size_2d_3d(pkg):
    """test_shape.py:10-41"""
    (xx, yy), (dx, dy) = _grid((701, 901), 1)
    z = 1 - np.sqrt(xx ** 2 + yy ** 2)
    area = pkg.shape.Size(ndim=2)(u=z, mask=_first_voxel_mask(xx.shape), dx=[dx, dy])
    almost(np.pi, area[0, 0], places=2)
    (xx, yy, zz), (dx, dy, dz) = _grid((401, 301, 501), pkg.sub3d)
    w = 1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)
    volume = pkg.shape.Size(ndim=3)(u=w, mask=_first_voxel_mask(xx.shape), dx=[dx, dy, dz])
    almost(4 * np.pi / 3, volume[0, 0, 0], places=2 if pkg.sub3d == 1 else 1)


def test_boundary_size_and_isoperimetric_2d(pkg):
    """test_shape.py:43-57, 76-90"""
    if not pkg.has_contours:
        pytest.skip("scikit-image (find_contours) is not installed: the reference leg cannot run")
    (xx, yy), (dx, dy) = _grid((501, 701), 1)
    z = 1 - np.sqrt(xx ** 2 + yy ** 2)
    curve_length = pkg.shape.BoundarySize(ndim=2)(u=z, mask=_first_voxel_mask(z.shape), dx=[dx, dy])
    almost(2 * np.pi, curve_length[0, 0], places=4)
    (xx, yy), (dx, dy) = _grid((701, 901), 1)
    z = 1 - np.sqrt(xx ** 2 + yy ** 2)
    ratio = pkg.shape.IsoperimetricRatio(ndim=2)(u=z, mask=_first_voxel_mask(xx.shape), dx=[dx, dy])
    almost(1, ratio[0, 0], places=3)


def test_boundary_size_and_isoperimetric_3d(pkg):
    """test_shape.py:59-74, 92-107 (sphere area 4 pi to 0 places, sphericity 1 to 2 places; the
    reference's meshgrid / dx ordering quirk is kept: arrays are (len(y), len(x), len(z)) while dx
    is passed as [dx, dy, dz])."""
    if pkg.name == "reference":
        pytest.skip("scikit-image (marching_cubes_lewiner) is not installed")
    sub = pkg.sub3d
    x, dx = np.linspace(-2, 2, (501 - 1) // sub + 1, retstep=True)
    y, dy = np.linspace(-2, 2, (401 - 1) // sub + 1, retstep=True)
    z, dz = np.linspace(-2, 2, (601 - 1) // sub + 1, retstep=True)
    xx, yy, zz = np.meshgrid(x, y, z)
    w = 1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)
    area = pkg.shape.BoundarySize(ndim=3)(u=w, mask=_first_voxel_mask(w.shape), dx=[dx, dy, dz])
    almost(4 * np.pi, area[0, 0, 0], places=0)
    x, dx = np.linspace(-2, 2, (201 - 1) // sub + 1, retstep=True)
    y, dy = np.linspace(-2, 2, (201 - 1) // sub + 1, retstep=True)
    z, dz = np.linspace(-2, 2, (901 - 1) // sub + 1, retstep=True)
    xx, yy, zz = np.meshgrid(x, y, z)
    w = 1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)
    ratio = pkg.shape.IsoperimetricRatio(ndim=3)(u=w, mask=_first_voxel_mask(w.shape),
                                                  dx=[dx, dy, dz])
    almost(1, ratio[0, 0, 0], places=2 if sub == 1 else 1)


def test_moments_2d(pkg):
    """test_shape.py:109-164"""
    (xx, yy), (dx, dy) = _grid((301, 501), 1)
    mask = _first_voxel_mask(xx.shape)
    moments = pkg.shape.Moments(ndim=2, axes=[0, 1], orders=[1])
    com = moments(u=1 - np.sqrt(xx ** 2 + yy ** 2), mask=mask, dx=[dy, dx])[mask].squeeze()
    almost(2.0, com[0], places=3)
    almost(2.0, com[1], places=3)
    z = 1 - np.sqrt((xx - 0.25) ** 2 + (yy + 0.25) ** 2)
    com = moments(u=z, mask=mask, dx=[dy, dx])[mask].squeeze()
    almost(1.75, com[0], places=3)
    almost(2.25, com[1], places=3)
    (xx, yy), (dx, dy) = _grid((301, 201), 1)
    mask = _first_voxel_mask(xx.shape)
    spread = pkg.shape.Moments(ndim=2, axes=[0, 1], orders=[2])(
        u=1 - np.sqrt(xx ** 2 + yy ** 2), mask=mask, dx=[dy, dx])[mask].squeeze()
    almost(0.25, spread[0], places=2)
    almost(0.25, spread[1], places=2)


def test_moments_3d(pkg):
    """test_shape.py:166-230"""
    (xx, yy, zz), (dx, dy, dz) = _grid((301, 501, 401), pkg.sub3d)
    mask = _first_voxel_mask(xx.shape)
    moments = pkg.shape.Moments(ndim=3, axes=[0, 1, 2], orders=[1])
    com = moments(u=1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2), mask=mask, dx=[dy, dx, dz])[mask].squeeze()
    for a in range(3):
        almost(2.0, com[a], places=3)
    w = 1 - np.sqrt((xx - 0.25) ** 2 + (yy + 0.25) ** 2 + (zz - 0.5) ** 2)
    com = moments(u=w, mask=mask, dx=[dy, dx, dz])[mask].squeeze()
    places = 3 if pkg.sub3d == 1 else 2
    almost(1.75, com[0], places=places)
    almost(2.25, com[1], places=places)
    almost(2.50, com[2], places=places)
    (xx, yy, zz), (dx, dy, dz) = _grid((301, 201, 401), pkg.sub3d)
    mask = _first_voxel_mask(xx.shape)
    spread = pkg.shape.Moments(ndim=3, axes=[0, 1, 2], orders=[2])(
        u=1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2), mask=mask, dx=[dy, dx, dz])[mask].squeeze()
    for a in range(3):
        almost(0.2, spread[a], places=2)


# --------------------------------------------------------------------------------------------
# lsml/feature/test/test_feature_map.py
# --------------------------------------------------------------------------------------------
def test_simple_feature_map(pkg):
    """test_feature_map.py:10-34 (smoke test of a mixed feature list)"""
    if not pkg.has_contours:
        pytest.skip("IsoperimetricRatio needs scikit-image on the reference leg")
    features = [pkg.image.ImageSample(sigma=0), pkg.image.ImageEdgeSample(sigma=3),
                pkg.shape.Size(), pkg.shape.IsoperimetricRatio(), pkg.shape.Moments()]
    feature_map = pkg.FeatureMap(features=features)
    random_state = np.random.RandomState(1234)
    img = random_state.randn(34, 67)
    u = random_state.randn(34, 67)
    mask = random_state.randn(34, 67) > 0
    out = feature_map(u=u, img=img, dist=u, mask=mask)
    assert out.shape == (34, 67, 8)


# --------------------------------------------------------------------------------------------
# lsml/initializer/provided/test/test_ball.py, lsml/initializer/test/test_initializer_base.py
# --------------------------------------------------------------------------------------------
def test_ball_initializers(pkg):
    """test_ball.py:11-74"""
    radius = 10
    u, _, _ = pkg.BallInitializer(radius=radius)(np.zeros((51, 51)))
    assert radius == (u[51 // 2] > 0).sum() // 2
    u, _, _ = pkg.BallInitializer(radius=radius)(np.zeros((31, 31, 31)))
    assert radius == (u[31 // 2, 31 // 2] > 0).sum() // 2
    u, _, _ = pkg.BallInitializer(radius=radius)(np.zeros((51, 51)), dx=[2, 0.5])
    assert radius * 2 ** -1 == (u[:, 25] > 0).sum() // 2
    assert radius * 0.5 ** -1 == (u[25, :] > 0).sum() // 2
    random_state = np.random.RandomState(1234)
    mask = np.pad(np.ones((5, 5), dtype=bool), 5, 'constant')
    image = np.zeros(mask.shape)
    image[mask] = 1 + 0.5 * random_state.randn(mask.sum())
    image[~mask] = -1 - 0.5 * random_state.rand((~mask).sum())
    u0, _, _ = pkg.RandomBallInitializer(random_state=random_state)(image)
    assert (u0 > 0).any()


def test_threshold_initializer_square_image(pkg):
    """initializer/provided/test/test_threshold.py:8-26"""
    if pkg.name == "reference":
        pytest.skip("scikit-image (threshold_otsu) is not installed: the reference leg cannot run")
    from lsml_b200.initializer import ThresholdInitializer
    random_state = np.random.RandomState(1234)
    mask = np.pad(np.ones((5, 5), dtype=bool), 5, 'constant')
    image = np.zeros(mask.shape)
    image[mask] = 1 + 0.5 * random_state.randn(mask.sum())
    image[~mask] = -1 - 0.5 * random_state.rand((~mask).sum())
    u0, _, _ = ThresholdInitializer(sigma=0)(image)
    assert (mask == (u0 > 0)).all()


def test_initializer_base_call(pkg):
    """test_initializer_base.py:8-27"""
    class MyInitializer(pkg.InitializerBase):
        def initialize(self, img, dx, seed):
            return img > 0

    image = np.random.RandomState(123).randn(41, 50)
    u0, _, _ = MyInitializer()(image)
    assert ((u0 > 0) == (image > 0)).all()


# --------------------------------------------------------------------------------------------
# lsml/util/test/test_distance_transform.py
# --------------------------------------------------------------------------------------------
def test_distance_transform(pkg):
    """test_distance_transform.py:11-53"""
    distance_transform = pkg.distance_transform
    dist, mask = distance_transform(np.r_[-1, -1, 1, -1, -1.], band=1, dx=[1.])
    assert (dist == np.r_[0., -0.5, 0.5, -0.5, 0.]).all()
    assert (mask == np.r_[False, True, True, True, False]).all()
    dist, mask = distance_transform(np.ones((4, 5, 6)), band=0, dx=[1, 1, 1.])
    assert mask.sum() == 0 and (dist == np.inf).all()
    dist, mask = distance_transform(-np.ones((4, 5, 6)), band=0, dx=[1, 1, 1.])
    assert mask.sum() == 0 and (dist == -np.inf).all()
    arr = np.ones((4, 5, 6))
    arr[2] = -1
    dist, mask = distance_transform(arr, band=0, dx=[1, 1, 1.])
    assert mask.sum() == arr.size
    dist, mask = distance_transform(np.zeros((4, 5, 6)), band=0, dx=[1, 1, 1.])
    assert mask.sum() == 120 and (dist == 0).all()


# --------------------------------------------------------------------------------------------
# lsml/gradient/test/test_masked_gradient.py
# --------------------------------------------------------------------------------------------
def test_gradient_centered(pkg):
    """test_masked_gradient.py:13-84"""
    mg = pkg.mg
    random_state = np.random.RandomState(1234)
    for variant in ("plain", "aniso", "mask"):
        for ndim in [1, 2, 3]:
            dims = random_state.randint(50, 101, size=ndim)
            arr = random_state.randn(*dims)
            dx = random_state.rand(ndim) if variant == "aniso" else None
            mask = arr > 0 if variant == "mask" else np.ones(arr.shape, dtype=bool)
            if variant == "aniso":
                grads, gmag = mg.gradient_centered(arr, dx=dx)
                numpy_grads = np.gradient(arr, *dx)
            elif variant == "mask":
                grads, gmag = mg.gradient_centered(arr, mask=mask)
                numpy_grads = np.gradient(arr)
            else:
                grads, gmag = mg.gradient_centered(arr)
                numpy_grads = np.gradient(arr)
            numpy_grads = [numpy_grads] if ndim == 1 else numpy_grads
            for i in range(ndim):
                assert np.abs(grads[i] - numpy_grads[i])[mask].mean() <= 1e-8
            numpy_gmag = np.linalg.norm(numpy_grads, axis=0)
            assert np.abs(gmag - numpy_gmag)[mask].mean() <= 1e-8
    for ndim in [1, 2, 3]:
        dims = random_state.randint(50, 101, size=ndim)
        arr = random_state.randn(*dims)
        grads = mg.gradient_centered(arr, normalize=True, return_gradient_magnitude=False)
        gmag_normalize = np.linalg.norm(grads, axis=0)
        assert np.abs(gmag_normalize - np.ones_like(gmag_normalize)).mean() <= 1e-8


def _manual_gmag_os(arr, nu, dx):
    """test_masked_gradient.py:86-109"""
    di = np.diff(arr, axis=0) / dx[0]
    dj = np.diff(arr, axis=1) / dx[1]
    fwdi = np.vstack([di, di[-1]])
    bcki = np.vstack([di[0], di])
    fwdj = np.c_[dj, dj[:, -1]]
    bckj = np.c_[dj[:, 0], dj]
    plus = np.sqrt(np.maximum(bcki, 0) ** 2 + np.minimum(fwdi, 0) ** 2 +
                   np.maximum(bckj, 0) ** 2 + np.minimum(fwdj, 0) ** 2)
    minus = np.sqrt(np.minimum(bcki, 0) ** 2 + np.maximum(fwdi, 0) ** 2 +
                    np.minimum(bckj, 0) ** 2 + np.maximum(fwdj, 0) ** 2)
    gmag = np.zeros_like(arr)
    gmag[nu > 0] = minus[nu > 0]
    gmag[nu < 0] = plus[nu < 0]
    return gmag


def test_gradient_magnitude_osher_sethian2d(pkg):
    """test_masked_gradient.py:111-121"""
    random_state = np.random.RandomState(1234)
    arr = random_state.randn(141, 112)
    nu = random_state.randn(141, 112)
    dx = random_state.rand(2)
    gmag = pkg.mg.gradient_magnitude_osher_sethian(arr, nu, dx=dx)
    assert np.abs(gmag - _manual_gmag_os(arr, nu, dx)).mean() <= 1e-8
