"""CPU-only, and only where /root/reference exists: the HOST side of the drop-in (names, sizes,
constructor errors, initializers, scores, balance_mask) compared directly with the unmodified
reference imported through oracle/ref_shim.py."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle.ref_shim import import_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference is not present on this machine")
    import_reference(skfmm_distance=oracle.skfmm_distance)
    import lsml
    return lsml


def test_feature_names_sizes_and_constructor_errors(ref):
    import lsml.feature.provided.image as rimg
    import lsml.feature.provided.shape as rshp
    from lsml.feature.feature_map import FeatureMap as RFM
    import lsml_b200.feature.provided.image as mimg
    import lsml_b200.feature.provided.shape as mshp
    from lsml_b200.feature.feature_map import FeatureMap as MFM
    pairs = []
    for ndim in (2, 3):
        for s in (0, 1.5, 3):
            for cls in ("ImageSample", "ImageEdgeSample", "InteriorImageAverage", "InteriorImageVariation"):
                pairs.append((getattr(rimg, cls)(ndim=ndim, sigma=s), getattr(mimg, cls)(ndim=ndim, sigma=s)))
            pairs.append((rimg.COMRaySample(ndim=ndim, sigma=s, n_samples=7),
                          mimg.COMRaySample(ndim=ndim, sigma=s, n_samples=7)))
        for cls in ("Size", "BoundarySize", "IsoperimetricRatio", "DistanceToCenterOfMass"):
            pairs.append((getattr(rshp, cls)(ndim=ndim), getattr(mshp, cls)(ndim=ndim)))
        pairs.append((rshp.Moments(ndim=ndim, axes=list(range(ndim)), orders=[1, 2]),
                      mshp.Moments(ndim=ndim, axes=list(range(ndim)), orders=[1, 2])))
    for r, m in pairs:
        for attr in ("name", "size", "type", "locality"):
            assert getattr(r, attr) == getattr(m, attr), (type(r).__name__, attr)
    for ndim in (2, 3):
        rfeats = rimg.get_basic_image_features(ndim=ndim, sigmas=[0, 2, 3]) + rshp.get_basic_shape_features(ndim=ndim)
        mfeats = mimg.get_basic_image_features(ndim=ndim, sigmas=[0, 2, 3]) + mshp.get_basic_shape_features(ndim=ndim)
        assert [f.name for f in rfeats] == [f.name for f in mfeats]
        rf, mf = RFM(rfeats), MFM(mfeats)
        assert rf.n_features == mf.n_features and rf.feature_slices == mf.feature_slices

    def err(f):
        try:
            f()
        except Exception as e:
            return type(e).__name__
        return None

    cases = [lambda M: M[0].COMRaySample(ndim=1), lambda M: M[1].Moments(ndim=2, axes=(2,)),
             lambda M: M[1].BoundarySize(ndim=1), lambda M: M[1].IsoperimetricRatio(ndim=1),
             lambda M: M[1].Moments(ndim=2, orders=(0,)), lambda M: M[2]([M[1].Size(2), M[1].Size(2)]),
             lambda M: M[2]([1, 2]), lambda M: M[2](5)]
    for c in cases:
        a, b = err(lambda: c((rimg, rshp, RFM))), err(lambda: c((mimg, mshp, MFM)))
        assert a == b and a is not None


def test_initializers_scores_and_balance_mask(ref):
    from lsml.initializer.provided.ball import BallInitializer as RB, RandomBallInitializer as RRB
    from lsml.score_functions import dice as rd, jaccard as rj
    from lsml.util.balance_mask import balance_mask as rbal
    from lsml_b200.initializer import BallInitializer as MB, RandomBallInitializer as MRB
    from lsml_b200.score_functions import dice as md, jaccard as mj
    from lsml_b200.util.balance_mask import balance_mask as mbal
    rs = np.random.RandomState(0)
    for t in range(120):
        ndim = rs.choice([1, 2, 3])
        shape = tuple(rs.randint(8, 30, size=ndim))
        dx = rs.uniform(0.5, 2.0, size=ndim)
        img = rs.randn(*shape)
        seed = np.array(shape) / 2
        loc = None if t % 2 else list(rs.uniform(2, 6, size=ndim))
        r = rs.uniform(2, 6)
        assert np.array_equal(RB(radius=r, location=loc).initialize(img=img, dx=dx, seed=seed),
                              MB(radius=r, location=loc).initialize(img=img, dx=dx, seed=seed))
        if ndim >= 2:
            a = RRB(random_state=np.random.RandomState(t), randomize_center=bool(t % 3))
            b = MRB(random_state=np.random.RandomState(t), randomize_center=bool(t % 3))
            assert np.array_equal(a.initialize(img=img, dx=dx, seed=seed),
                                  b.initialize(img=img, dx=dx, seed=seed))
        u, seg = rs.randn(*shape), rs.rand(*shape) > 0.5
        assert rj(u, seg) == mj(u, seg) and rd(u, seg) == md(u, seg) and rj(u, seg, 0.3) == mj(u, seg, 0.3)
        n = rs.randint(1, 60)
        arr = rs.randn(n) * rs.choice([0, 1, 1, 1], size=n)
        if t % 7 == 0:
            arr = np.abs(arr)
        if t % 11 == 0:
            arr = -np.abs(arr)
        r1, r2 = np.random.RandomState(t), np.random.RandomState(t)
        assert np.array_equal(rbal(arr, r1), mbal(arr, r2))
        assert r1.randint(1 << 30) == r2.randint(1 << 30)            # same RNG consumption
    assert mj(np.zeros((4, 4)) - 1, np.zeros((4, 4), bool)) == 1.0


def test_masked_gradient_validation_errors(ref):
    """Same exception type AND message, in the reference's order of checks
    (masked_gradient.py:112-127, 202-217); raised before any kernel is reached."""
    from lsml.gradient import masked_gradient as rmg
    from lsml_b200.gradient import masked_gradient as mmg

    def err(f):
        try:
            f()
        except BaseException as e:
            return type(e).__name__, str(e)
        return None

    a4, a2 = np.zeros((2, 2, 2, 2)), np.zeros((4, 5))
    f32 = np.zeros((4, 5), dtype=np.float32)
    cases = [lambda m: m.gradient_centered(a4), lambda m: m.gradient_centered(f32),
             lambda m: m.gradient_centered(a2, mask=np.ones(4, bool)),
             lambda m: m.gradient_centered(a2, dx=[1.0]),
             lambda m: m.gradient_centered(f32, mask=np.ones(4, bool), dx=[1.0]),
             lambda m: m.gradient_magnitude_osher_sethian(a4, a4),
             lambda m: m.gradient_magnitude_osher_sethian(f32, f32),
             lambda m: m.gradient_magnitude_osher_sethian(a2, a2, mask=np.ones(4, bool)),
             lambda m: m.gradient_magnitude_osher_sethian(a2, a2, dx=[1.0, 2.0, 3.0])]
    for c in cases:
        want, got = err(lambda: c(rmg)), err(lambda: c(mmg))
        assert want is not None and want == got


def test_threshold_ball_initializer_host_logic(ref, oracle, monkeypatch):
    """ThresholdBallInitializer (ball.py:118-159): the reference with the oracle's marcher plugged
    in for skfmm vs ours with the same marcher plugged in for the device call -- everything
    else (smoothing, labelling, seed handling, the ball) is host code and must agree."""
    from lsml.initializer.provided.ball import ThresholdBallInitializer as RTB
    from lsml_b200.initializer.provided import ball as mball
    import lsml_b200.util.distance_transform as mdt

    def oracle_dt(arr, band, dx):
        d = oracle.skfmm_distance(arr, dx=dx)
        return np.asarray(d), np.ones(arr.shape, dtype=bool)

    monkeypatch.setattr(mdt, "distance_transform", oracle_dt)
    rs = np.random.RandomState(4)
    checked = 0
    for t in range(12):
        shape = (40, 48) if t % 2 else (18, 20, 22)
        g = np.indices(shape).astype(float)
        c = np.array(shape) / 2.0 + rs.uniform(-2, 2, size=len(shape))
        r = np.sqrt(sum((g[a] - c[a]) ** 2 for a in range(len(shape))))
        img = (r < 6).astype(float) * 2.0 + 0.3 * rs.randn(*shape)
        dx = np.ones(len(shape))
        seed = np.round(c)
        want = RTB(sigma=3.0).initialize(img=img, dx=dx, seed=seed)
        got = mball.ThresholdBallInitializer(sigma=3.0).initialize(img=img, dx=dx, seed=seed)
        assert got.dtype == bool and np.array_equal(want, got)
        checked += int(got.any())
    assert checked >= 8
    # the seed outside every component: the nearest component is used (the reference's own
    # fallback multiplies an integer array by dx in place and fails on current numpy)
    img = np.zeros((30, 30))
    img[18:25, 18:25] = 5.0
    got = mball.ThresholdBallInitializer(sigma=2.0).initialize(img=img, dx=np.ones(2), seed=np.array([3.0, 3.0]))
    assert got.any() and got[18:25, 18:25].any() and not got[:10, :10].any()


def test_threshold_initializer_reference_test():
    """initializer/provided/test/test_threshold.py:8-26 on the host part (`initialize`)."""
    from lsml_b200.initializer import ThresholdInitializer
    random_state = np.random.RandomState(1234)
    mask = np.pad(np.ones((5, 5), dtype=bool), 5, 'constant')
    image = np.zeros(mask.shape)
    image[mask] = 1 + 0.5 * random_state.randn(mask.sum())
    image[~mask] = -1 - 0.5 * random_state.rand((~mask).sum())
    assert (mask == ThresholdInitializer(sigma=0).initialize(image, np.ones(2), None)).all()
    blurred = ThresholdInitializer(sigma=1.0).initialize(image, np.ones(2), None)
    assert blurred.dtype == bool and blurred[7, 7] and not blurred[0, 0]
    with pytest.raises(ValueError):
        ThresholdInitializer(sigma=-1)


def test_base_feature_error_messages(ref):
    """Same exception type and text as the reference for every kind of bad call
    (feature/base_feature.py:67-139)."""
    from lsml.feature.provided.image import ImageSample as RI
    from lsml.feature.provided.shape import Size as RS
    from lsml_b200.feature import ImageSample as MI, Size as MS

    def err(f):
        try:
            f()
        except Exception as e:
            return type(e).__name__, str(e)
        return None

    u, img = np.ones((3, 2)), np.ones((3, 2))
    mask = np.ones((3, 2), dtype=bool)
    calls = [
        lambda I, S: I(ndim=2, sigma=0)(u='x'),
        lambda I, S: I(ndim=2, sigma=0)(u=u),
        lambda I, S: I(ndim=2, sigma=0)(u=u, img='x'),
        lambda I, S: I(ndim=2, sigma=0)(u=u, img=np.ones((3, 1))),
        lambda I, S: S(ndim=2)(u=u, dist='x'),
        lambda I, S: S(ndim=2)(u=u, dist=np.ones((3,))),
        lambda I, S: S(ndim=2)(u=u, dx=[1]),
        lambda I, S: S(ndim=2)(u=u, mask='x'),
        lambda I, S: S(ndim=2)(u=u, mask=np.ones((3, 9), dtype=bool)),
        lambda I, S: S(ndim=2)(u=u, mask=u),
        lambda I, S: I(ndim=2, sigma=0)(u=u, img=img, dist=np.ones((2, 2)), mask=mask, dx=[1, 2, 3]),
    ]
    for c in calls:
        want, got = err(lambda: c(RI, RS)), err(lambda: c(MI, MS))
        assert want is not None and want == got, (want, got)
