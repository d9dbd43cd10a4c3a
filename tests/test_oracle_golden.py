"""CPU-only: the oracle (oracle/lsml_oracle.c + oracle/oracle.py) against the golden fixtures
generated from the UNMODIFIED reference (oracle/gen_golden.py).  This is what pins the oracle:
its stencils, feature restatements and evolution loop must reproduce the reference's own
outputs bit for bit; the GPU tests then compare the CUDA path with the oracle and with the same
fixtures."""
import numpy as np

from golden_util import load, regressors_of, specs_of


def test_stencils_bit_exact(oracle):
    g = load("masked_gradient.npz")
    for ndim in (1, 2, 3):
        arr, nu, mask, dx = g["arr%d" % ndim], g["nu%d" % ndim], g["mask%d" % ndim], g["dx%d" % ndim]
        for use_ref in (False, True):
            if use_ref and oracle.reference_lib() is None:
                continue
            assert np.array_equal(oracle.gmag_os(arr, nu, mask, dx, use_reference=use_ref),
                                  g["gmag_os%d" % ndim])
            assert np.array_equal(oracle.gmag_os(arr, nu, use_reference=use_ref),
                                  g["gmag_os_full%d" % ndim])
            for norm in (0, 1):
                grads, gm = oracle.gradient_centered(arr, mask, dx, bool(norm), use_reference=use_ref)
                assert np.array_equal(np.stack(grads), g["gc%d_n%d_grads" % (ndim, norm)])
                assert np.array_equal(gm, g["gc%d_n%d_gmag" % (ndim, norm)])


def test_feature_restatements_bit_exact(oracle):
    f = load("features.npz")
    specs2, specs3 = specs_of(f, "specs2"), specs_of(f, "specs3")
    for b in range(3):
        u, mask, dx = f["u2_%d" % b], f["mask2_%d" % b], f["dx2_%d" % b]
        X = oracle.feature_matrix(specs2, u, f["imgs"][b], None, mask, dx)
        assert np.array_equal(X, f["X2_%d" % b])
        # the band mask itself (reference wrapper + plugged marcher) is reproduced too
        _, m = oracle.distance_transform(u, 3, dx)
        assert np.array_equal(m, mask)
    X = oracle.feature_matrix(specs3, f["u3"], f["img3"], None, f["mask3"], f["dx3"])
    assert np.array_equal(X, f["X3"])
    _, m = oracle.distance_transform(f["u3"], 3, f["dx3"])
    assert np.array_equal(m, f["mask3"])


def test_every_fixture_band_is_the_current_oracles(oracle):
    """The fixtures' band masks come from the reference's distance_transform wrapper with the
    oracle's marcher plugged in for the absent scikit-fmm: a fixture generated before a change of
    the marcher (e.g. the contributor rule) must be regenerated, or the GPU golden
    tests would compare against a stale band."""
    for name in ("com_ray.npz", "prev_iter.npz"):
        g = load(name)
        for c in range(int(g["n_cases"])):
            _, m = oracle.distance_transform(g["u%d" % c], 3, g["dx%d" % c])
            assert np.array_equal(m, g["mask%d" % c]), (name, c)


def test_initializers(oracle):
    ini = load("initializers.npz")
    u0 = oracle.ball_init((51, 51), 10, np.ones(2))
    assert np.array_equal(u0, ini["ball_u0"])
    _, m = oracle.distance_transform(u0, 3, np.ones(2))
    assert np.array_equal(m, ini["ball_mask"])
    dx = np.array([1.0, 0.8, 1.2])
    u0 = oracle.ball_init((20, 22, 18), 4, dx, location=[10, 12, 9])
    assert np.array_equal(u0, ini["ball3_u0"])
    _, m = oracle.distance_transform(u0, 2, dx)
    assert np.array_equal(m, ini["ball3_mask"])


def test_segment_loop_reproduces_reference_iterates(oracle):
    """oracle.segment == the reference's LevelSetMachineLearning.segment output `us`."""
    s = load("segment.npz")
    specs = specs_of(load("features.npz"), "specs2")
    for kind in ("linear", "forest"):
        regs = regressors_of(s, kind)
        img = s["%s_img" % kind]
        us, masks, _ = oracle.segment(img, oracle.ball_init(img.shape, 10, np.ones(2)), np.ones(2),
                                      specs, regs, float(s["step"]), float(s["band"]))
        want = s["%s_us" % kind]
        assert us.shape == want.shape
        if kind == "forest":
            assert np.array_equal(us, want)
        else:   # BLAS dot-product order inside sklearn vs numpy: last-bit differences
            assert np.allclose(us, want, rtol=1e-12, atol=1e-12)
        assert np.array_equal(us > 0, want > 0)
