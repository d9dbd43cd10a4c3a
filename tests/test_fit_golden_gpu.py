"""GPU: ``lsml_b200``'s ``fit`` + ``segment`` with the real device engine against the reference's
OWN ``fit`` (tests/golden/fit.npz; see tests/test_fit_host_logic_cpu.py for the same comparison
with the oracle's numerics, where it is bit-exact)."""
import numpy as np
import pytest

from golden_util import load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,balance,seed", [("bal", True, 77), ("nobal", False, 5),
                                              ("forest", True, 9), ("seedball", True, 21),
                                              ("aniso", True, 33)])
def test_fit_reproduces_the_references_fit_on_device(tag, balance, seed):
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from lsml_b200 import LevelSetMachineLearning
    from lsml_b200.feature import (DistanceToCenterOfMass, Moments, Size, get_basic_image_features)
    from lsml_b200.initializer import BallInitializer
    g = load("fit.npz")
    imgs, segs = g["imgs"], g["segs"]
    feats = (get_basic_image_features(ndim=2, sigmas=[0, 2]) +
             [DistanceToCenterOfMass(ndim=2), Moments(ndim=2), Size(ndim=2)])
    from test_fit_host_logic_cpu import seed_ball_initializer
    init = seed_ball_initializer() if tag == "seedball" else BallInitializer(radius=6)
    model = LevelSetMachineLearning(feats, init, band=3)
    rs = np.random.RandomState(seed)
    final = (RandomForestRegressor(n_estimators=10, max_samples=200, random_state=rs)
             if tag == "forest" else LinearRegression())
    model.fit(imgs=list(imgs[:14]), segs=list(segs[:14]), max_iters=7,
              random_state=rs, regression_model_class=Pipeline,
              regression_model_kwargs=dict(steps=[('s', StandardScaler()), ('r', final)]),
              dx=(np.tile(g["aniso_dx"], (14, 1)) if tag == "aniso" else None),
              balance_regression_targets=balance, validation_history_len=4,
              validation_history_tol=-1.0)
    assert rs.randint(1 << 30) == int(g[tag + "_rng_after"])
    for key in ("training", "validation", "testing"):
        assert list(model._split[key]) == list(g["%s_%s_idx" % (tag, key)])
    assert model.step == float(g[tag + "_step"])            # max |ground-truth distance| on the band
    assert model.iteration == int(g[tag + "_iterations"])
    for i in range(1, model.iteration + 1):
        reg = model.regression_model(i)
        # the training matrices agree to the last bits (band statistics are summed in another
        # order); least squares amplifies that by its conditioning
        assert np.allclose(reg.steps[0][1].mean_, g["%s_mean_%d" % (tag, i)], rtol=1e-9, atol=1e-12), i
        assert np.allclose(reg.steps[0][1].scale_, g["%s_scale_%d" % (tag, i)], rtol=1e-9, atol=1e-12), i
        # (the per-example features make the least-squares problem nearly rank deficient with a
        # dozen examples: last-bit differences of the training matrix are amplified)
        assert np.allclose(reg.predict(g["probe"]), g["%s_probe_%d" % (tag, i)], rtol=1e-4, atol=1e-4), i
        if tag != "forest":
            assert np.allclose(reg.steps[1][1].coef_, g["%s_coef_%d" % (tag, i)], rtol=1e-4, atol=1e-5), i
    for key in ("training", "validation", "testing"):
        assert np.allclose(getattr(model, key + "_scores"), g["%s_%s_scores" % (tag, key)],
                           rtol=0, atol=1e-9), key
    for flag, name in ((False, "_us"), (True, "_us_valmax")):
        us = model.segment(imgs[14], dx=(g["aniso_dx"] if tag == "aniso" else None), verbose=False,
                           iterate_until_validation_max=flag)
        want = g[tag + name]
        assert us.shape == want.shape
        for i in range(us.shape[0]):
            assert np.array_equal(us[i] > 0, want[i] > 0), i
            assert np.allclose(us[i], want[i], rtol=1e-5, atol=1e-8), np.abs(us[i] - want[i]).max()


def test_fit_3d_reproduces_the_references_fit_on_device():
    """8 volumes 25^3, F = 16, 12-tree forest: the C3 workload family end to end (fit3d.npz)."""
    from lsml_b200.core import model as M
    from test_fit_host_logic_cpu import fit3d_model
    g = load("fit3d.npz")
    model, rs = fit3d_model(M, g, "cuda")
    assert rs.randint(1 << 30) == int(g["rng_after"])
    assert model.step == float(g["step"]) and model.iteration == int(g["iterations"])
    for i in range(1, model.iteration + 1):
        assert np.allclose(model.regression_model(i).predict(g["probe"]), g["probe_%d" % i],
                           rtol=1e-4, atol=1e-4), i
    for key in ("training", "validation", "testing"):
        assert np.allclose(getattr(model, key + "_scores"), g[key + "_scores"], rtol=0, atol=1e-9)
    us = model.segment(g["imgs"][8], verbose=False, iterate_until_validation_max=False)
    want = g["us"]
    assert us.shape == want.shape
    for i in range(us.shape[0]):
        assert np.array_equal(us[i] > 0, want[i] > 0), i
        assert np.allclose(us[i], want[i], rtol=1e-5, atol=1e-8), np.abs(us[i] - want[i]).max()


def test_fit_uses_the_models_scorer():
    """fit_job_handler.py:258-291 records ``model.scorer(u, seg)``: dice comes from the same exact
    device counts as jaccard, any other callable is evaluated on host copies."""
    import numpy as np
    from sklearn.linear_model import LinearRegression
    from lsml_b200 import LevelSetMachineLearning
    from lsml_b200.feature import get_basic_image_features
    from lsml_b200.initializer import BallInitializer
    from lsml_b200.score_functions import dice, jaccard
    rs = np.random.RandomState(0)
    g = np.indices((33, 33)).astype(float) - 16
    imgs = [((g ** 2).sum(0) < r * r).astype(float) + 0.1 * rs.randn(33, 33) for r in (8, 9, 10, 11, 7, 12)]
    segs = [((g ** 2).sum(0) < r * r) for r in (8, 9, 10, 11, 7, 12)]

    def custom(u, seg, threshold=0.0):
        return float(((u > threshold) == seg).mean())

    got = {}
    for name, scorer in (("jaccard", jaccard), ("dice", dice), ("custom", custom)):
        m = LevelSetMachineLearning(get_basic_image_features(ndim=2, sigmas=[0, 2]),
                                    BallInitializer(radius=5), scorer=scorer)
        m.fit(imgs=imgs, segs=segs, regression_model_class=LinearRegression, max_iters=3,
              datasets_split=([0, 1, 2, 3], [4], [5]), step=0.5, validation_history_len=10,
              random_state=np.random.RandomState(1))
        got[name] = np.asarray(m.training_scores)
    j = got["jaccard"]
    assert np.allclose(got["dice"], 2 * j / (j + 1), rtol=0, atol=1e-15)
    assert got["custom"].shape == j.shape and not np.allclose(got["custom"], j)
    assert (got["custom"] > 0.5).all()
