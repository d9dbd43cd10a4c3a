"""CPU-only: the HOST logic of ``lsml_b200``'s ``fit`` and ``segment`` against the reference's
OWN ``fit`` (tests/golden/fit.npz, made by oracle/gen_golden.py from the unmodified reference
with an in-memory h5py stand-in).

The device engine is replaced by tests/oracle_engine.py (oracle numerics), everything else is
the product's code: dataset split and RNG stream, ball initializer parameters, ground-truth
distances as regression targets, the automatic step, per-example target balancing, the order of
the training rows, the per-iteration scikit-learn fit, score bookkeeping, ``segment``'s
reconstruction of all iterates.  With identical numerics underneath, the result has to be the
reference's bit for bit.  tests/test_model_gpu.py repeats the comparison with the real engine."""
import numpy as np
import pytest

from golden_util import load


@pytest.fixture()
def patched_model(monkeypatch, oracle):
    import oracle_engine
    from lsml_b200.core import model as M
    monkeypatch.setattr(M, "BandEngine", oracle_engine.OracleEngine)
    monkeypatch.setattr(M, "_ball_u0", oracle_engine.ball_u0)
    monkeypatch.setattr(M, "overlap_scores", oracle_engine.overlap_scores)
    monkeypatch.setattr(M.LevelSetMachineLearning, "_device_model",
                        lambda self, i, device: self.regression_models[i])
    return M


def reference_features():
    from lsml_b200.feature import (DistanceToCenterOfMass, Moments, Size, get_basic_image_features)
    return (get_basic_image_features(ndim=2, sigmas=[0, 2]) +
            [DistanceToCenterOfMass(ndim=2), Moments(ndim=2), Size(ndim=2)])


def seed_ball_initializer():
    """The user initializer of the `seedball` golden case (oracle/gen_golden.py:gen_fit): runs on
    the host and USES the seed, which defaults to the centre of mass of the segmentation."""
    from lsml_b200.initializer import InitializerBase

    class SeedBall(InitializerBase):
        def initialize(self, img, dx, seed):
            ind = np.indices(img.shape, dtype=float)
            for a in range(img.ndim):
                ind[a] = ind[a] * dx[a] - seed[a]
            return np.sqrt((ind ** 2).sum(axis=0)) < 5.0

    return SeedBall()


@pytest.mark.parametrize("tag,balance,seed", [("bal", True, 77), ("nobal", False, 5),
                                              ("forest", True, 9), ("seedball", True, 21),
                                              ("aniso", True, 33)])
def test_fit_reproduces_the_references_fit(patched_model, tag, balance, seed):
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from lsml_b200.initializer import BallInitializer
    g = load("fit.npz")
    imgs, segs = g["imgs"], g["segs"]
    init = seed_ball_initializer() if tag == "seedball" else BallInitializer(radius=6)
    model = patched_model.LevelSetMachineLearning(reference_features(), init, band=3)
    rs = np.random.RandomState(seed)
    # "forest": ONE RandomState for fit() and for the forest, as the reference's examples do; the
    # reference fits in a forked child, so the forest never advances the caller's stream
    final = (RandomForestRegressor(n_estimators=10, max_samples=200, random_state=rs)
             if tag == "forest" else LinearRegression())
    model.fit(imgs=list(imgs[:14]), segs=list(segs[:14]), max_iters=7,
              random_state=rs, regression_model_class=Pipeline,
              regression_model_kwargs=dict(steps=[('s', StandardScaler()), ('r', final)]),
              dx=(np.tile(g["aniso_dx"], (14, 1)) if tag == "aniso" else None),
              balance_regression_targets=balance, validation_history_len=4,
              validation_history_tol=-1.0, device="cpu")
    assert rs.randint(1 << 30) == int(g[tag + "_rng_after"])
    for key in ("training", "validation", "testing"):
        assert list(model._split[key]) == list(g["%s_%s_idx" % (tag, key)])
    assert model.step == float(g[tag + "_step"])
    assert model.iteration == int(g[tag + "_iterations"])
    for i in range(1, model.iteration + 1):
        reg = model.regression_model(i)
        assert np.array_equal(reg.steps[0][1].mean_, g["%s_mean_%d" % (tag, i)]), i
        assert np.array_equal(reg.steps[0][1].scale_, g["%s_scale_%d" % (tag, i)]), i
        assert np.array_equal(reg.predict(g["probe"]), g["%s_probe_%d" % (tag, i)]), i
        if tag != "forest":
            assert np.array_equal(reg.steps[1][1].coef_, g["%s_coef_%d" % (tag, i)]), i
            assert np.array_equal(reg.steps[1][1].intercept_, g["%s_intercept_%d" % (tag, i)]), i
    assert np.array_equal(model.training_scores, g[tag + "_training_scores"])
    assert np.array_equal(model.validation_scores, g[tag + "_validation_scores"])
    assert np.array_equal(model.testing_scores, g[tag + "_testing_scores"])
    sdx = g["aniso_dx"] if tag == "aniso" else None
    us = model.segment(imgs[14], dx=sdx, verbose=False, iterate_until_validation_max=False, device="cpu")
    assert np.array_equal(us, g[tag + "_us"])
    us = model.segment(imgs[14], dx=sdx, verbose=False, device="cpu")
    assert np.array_equal(us, g[tag + "_us_valmax"])


def test_fit_input_validation_as_the_reference(patched_model):
    """datasets_handler.py:131-170: same exception types for bad `imgs` / `segs` / `dx`."""
    from sklearn.linear_model import LinearRegression
    from lsml_b200.initializer import BallInitializer
    M = patched_model
    img, seg = np.zeros((8, 9)), np.zeros((8, 9), dtype=bool)

    def fit(imgs, segs, **kw):
        M.LevelSetMachineLearning(reference_features(), BallInitializer(radius=2)).fit(
            imgs=imgs, segs=segs, regression_model_class=LinearRegression,
            random_state=np.random.RandomState(0), device="cpu", **kw)

    with pytest.raises(ValueError):
        fit([img, img], [seg])
    with pytest.raises(TypeError):
        fit([img.astype(np.float32)], [seg])
    with pytest.raises(TypeError):
        fit([img], [seg.astype(np.uint8)])
    with pytest.raises(ValueError):
        fit([img, np.zeros((8, 9, 2))], [seg, np.zeros((8, 9, 2), dtype=bool)])
    with pytest.raises(ValueError):
        fit([img], [np.zeros((8, 8), dtype=bool)])
    with pytest.raises(ValueError):
        fit([img, img], [seg, seg], dx=np.ones((3, 2)))
    with pytest.raises(NotImplementedError):      # the one deliberate restriction
        fit([img, np.zeros((9, 9))], [seg, np.zeros((9, 9), dtype=bool)])


def fit3d_model(M, g, device):
    """The 3-D golden case (oracle/gen_golden.py:gen_fit3d) through `M.LevelSetMachineLearning`."""
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from lsml_b200.feature import (DistanceToCenterOfMass, Moments, Size, get_basic_image_features)
    from lsml_b200.initializer import BallInitializer
    feats = (get_basic_image_features(ndim=3, sigmas=[0, 1.5]) +
             [DistanceToCenterOfMass(ndim=3), Moments(ndim=3, axes=(0, 1, 2), orders=(1, 2)),
              Size(ndim=3)])
    rs = np.random.RandomState(12)
    model = M.LevelSetMachineLearning(feats, BallInitializer(radius=4), band=3)
    model.fit(imgs=list(g["imgs"][:8]), segs=list(g["segs"][:8]), max_iters=5, random_state=rs,
              regression_model_class=Pipeline,
              regression_model_kwargs=dict(steps=[
                  ('s', StandardScaler()),
                  ('r', RandomForestRegressor(n_estimators=12, max_samples=300, random_state=rs))]),
              datasets_split=([0, 1, 2, 3, 4], [5, 6], [7]), validation_history_len=3,
              validation_history_tol=-1.0, device=device)
    return model, rs


def test_fit_3d_reproduces_the_references_fit(patched_model):
    g = load("fit3d.npz")
    model, rs = fit3d_model(patched_model, g, "cpu")
    assert rs.randint(1 << 30) == int(g["rng_after"])
    assert model.step == float(g["step"]) and model.iteration == int(g["iterations"])
    for i in range(1, model.iteration + 1):
        assert np.array_equal(model.regression_model(i).predict(g["probe"]), g["probe_%d" % i]), i
    for key in ("training", "validation", "testing"):
        assert np.array_equal(getattr(model, key + "_scores"), g[key + "_scores"])
    us = model.segment(g["imgs"][8], verbose=False, iterate_until_validation_max=False, device="cpu")
    assert np.array_equal(us, g["us"])
