"""GPU: the CUDA path against the golden fixtures made from the unmodified reference."""
import numpy as np
import pytest

from golden_util import load, regressors_of, specs_of

pytestmark = pytest.mark.gpu


def test_dropin_stencils_reproduce_reference_outputs():
    from lsml_b200.gradient import masked_gradient as mg
    g = load("masked_gradient.npz")
    for ndim in (1, 2, 3):
        arr, nu, mask, dx = g["arr%d" % ndim], g["nu%d" % ndim], g["mask%d" % ndim], g["dx%d" % ndim]
        assert np.array_equal(mg.gradient_magnitude_osher_sethian(arr, nu, mask, dx),
                              g["gmag_os%d" % ndim])
        assert np.array_equal(mg.gradient_magnitude_osher_sethian(arr, nu), g["gmag_os_full%d" % ndim])
        for norm in (0, 1):
            grads, gm = mg.gradient_centered(arr, mask, dx, normalize=bool(norm))
            assert np.array_equal(np.stack(grads), g["gc%d_n%d_grads" % (ndim, norm)])
            assert np.array_equal(gm, g["gc%d_n%d_gmag" % (ndim, norm)])


def test_feature_matrix_reproduces_reference_feature_map():
    from lsml_b200.core.engine import BandEngine
    f = load("features.npz")
    cases = [(specs_of(f, "specs2"), f["u2_%d" % b], f["imgs"][b], f["mask2_%d" % b],
              f["dx2_%d" % b], f["X2_%d" % b]) for b in range(3)]
    cases.append((specs_of(f, "specs3"), f["u3"], f["img3"], f["mask3"], f["dx3"], f["X3"]))
    for specs, u, img, mask, dx, want in cases:
        eng = BandEngine(u.shape, 1, dx=dx, band=3.0, specs=specs)
        eng.load_images(img[None], normalize=False)
        eng.set_level_sets(u[None])
        # band rebuilt on the device == the reference's mask (bit-exact)
        assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), mask)
        X = eng.compute_features().cpu().numpy()
        assert X.shape == want.shape
        cols = []
        for s in specs:
            cols += [s[0]] * (len(s[1]) * len(s[2]) if s[0] == "moments" else 1)
        for j, kind in enumerate(cols):
            if kind in ("image_sample", "image_edge", "size") or \
                    (kind == "dist_to_com" and (dx == 1).all()):
                assert np.array_equal(X[:, j], want[:, j]), (kind, j)
            else:   # reductions: summation order differs from numpy's pairwise sum
                assert np.allclose(X[:, j], want[:, j], rtol=1e-12, atol=1e-13), (kind, j)


def test_initializers_on_device():
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.model import _ball_u0
    from lsml_b200.initializer import BallInitializer, RandomBallInitializer
    ini = load("initializers.npz")
    eng = BandEngine((51, 51), 1, dx=np.ones(2), band=3.0, specs=[])
    centre, radius = BallInitializer(radius=10).device_ball((51, 51), np.ones(2))
    u0 = _ball_u0(eng, centre[None], [radius])
    assert np.array_equal(u0[0].cpu().numpy(), ini["ball_u0"])
    eng.set_level_sets(u0)
    assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), ini["ball_mask"])
    dx = np.array([1.0, 0.8, 1.2])
    eng = BandEngine((20, 22, 18), 1, dx=dx, band=2.0, specs=[])
    centre, radius = BallInitializer(radius=4, location=[10, 12, 9]).device_ball((20, 22, 18), dx)
    u0 = _ball_u0(eng, centre[None], [radius])
    assert np.array_equal(u0[0].cpu().numpy(), ini["ball3_u0"])
    eng.set_level_sets(u0)
    assert np.array_equal(eng.mask[0].cpu().numpy().astype(bool), ini["ball3_mask"])
    # host-side mirror of the reference's RandomBallInitializer (RNG order preserved)
    u0, dist, mask = RandomBallInitializer(random_state=np.random.RandomState(99))(
        ini["rand_img"], band=3, dx=np.ones(2))
    assert np.array_equal(u0, ini["rand_u0"]) and np.array_equal(mask, ini["rand_mask"])


@pytest.mark.parametrize("kind", ["linear", "forest"])
def test_segment_reproduces_reference_segment(kind):
    """LevelSetMachineLearning.segment vs the `us` returned by the reference's own segment()."""
    from lsml_b200 import LevelSetMachineLearning
    from lsml_b200.core.regressors import ForestDeviceRegressor, LinearDeviceRegressor, _dev
    from lsml_b200.feature import (DistanceToCenterOfMass, Moments, Size,
                                   get_basic_image_features)
    from lsml_b200.initializer import BallInitializer
    import torch
    s = load("segment.npz")
    dev = torch.device("cuda")
    regs = []
    for i, r in enumerate(regressors_of(s, kind)):
        if kind == "linear":
            regs.append(LinearDeviceRegressor(r.coef, r.intercept, (r.mean, r.scale), dev))
        else:
            fr = ForestDeviceRegressor.__new__(ForestDeviceRegressor)
            from lsml_b200.core.regressors import DeviceRegressor
            DeviceRegressor.__init__(fr, len(r.mean), (r.mean, r.scale), dev)
            bases = np.asarray(r.bases, dtype=np.int32)
            if len(bases) == len(r.root):      # fixture written before the node-count sentinel
                bases = np.append(bases, np.int32(r.nodes.shape[0]))
            fr.nodes, fr.tree_base = _dev(r.nodes, dev), _dev(bases, dev)
            fr.root_is_leaf, fr.n_trees = _dev(r.root, dev), len(r.root)
            fr.n_nodes, fr.divisor = int(r.nodes.shape[0]), float(len(r.root))
            regs.append(fr)
    feats = (get_basic_image_features(ndim=2, sigmas=[0, 2]) +
             [DistanceToCenterOfMass(ndim=2), Moments(ndim=2), Size(ndim=2)])
    model = LevelSetMachineLearning.from_regressors(feats, BallInitializer(radius=10), regs,
                                                    step=float(s["step"]), band=float(s["band"]))
    us = model.segment(s["%s_img" % kind], verbose=False, iterate_until_validation_max=False)
    want = s["%s_us" % kind]
    assert us.shape == want.shape
    for i in range(us.shape[0]):
        # segmentation masks bit-exact; level sets within 1e-5 relative (north_star bar)
        assert np.array_equal(us[i] > 0, want[i] > 0), i
        assert np.allclose(us[i], want[i], rtol=1e-5, atol=1e-9), np.abs(us[i] - want[i]).max()
