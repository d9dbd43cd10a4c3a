"""CPU-only: the user-feature registry.

`smooth_u` is the reference example's PrevIterFeature (examples/gestalt2d/prev_iter_feature.py):
the oracle against golden vectors made by that class itself, and the decomposition the engine
relies on (a chain of scipy gaussian_filter1d passes == gaussian_filter, bit for bit)."""
import numpy as np
import pytest
from scipy.ndimage import gaussian_filter

from golden_util import load


def cases():
    g = load("prev_iter.npz")
    for c in range(int(g["n_cases"])):
        yield g["u%d" % c], g["mask%d" % c], g["dx%d" % c], g["sigmas%d" % c], g["X%d" % c]


def test_oracle_smooth_u_reproduces_reference_example(oracle):
    for u, mask, dx, sigmas, want in cases():
        specs = [("smooth_u", float(s)) for s in sigmas]
        got = oracle.feature_matrix(specs, u, np.zeros_like(u), None, mask, dx)
        assert np.array_equal(got, want)
        for j, s in enumerate(sigmas):
            # what lsml_b200.core.engine._smooth_chain runs on the device: one 1-D pass per axis
            chain = oracle.smooth(u, float(s), np.ones(u.ndim)) if s > 0 else u
            assert np.array_equal(chain[mask], want[:, j])
            assert np.array_equal(chain, gaussian_filter(u, sigma=float(s)))


def test_feature_classes_and_program():
    from lsml_b200.core.engine import FeatureProgram
    from lsml_b200.feature import (FeatureMap, ImageSample, PrecomputedImageFeature,
                                   PrevIterFeature, SmoothedLevelSet)

    class Corner(PrecomputedImageFeature):
        @property
        def name(self):
            return "corner-reponse-sigma={:.2f}".format(self.sigma)

        def compute_plane(self, img, dx):
            return gaussian_filter(img, sigma=self.sigma) ** 2

    assert PrevIterFeature is SmoothedLevelSet
    assert PrevIterFeature(sigma=2).name == "smooth-ls-sigma=2.00"       # the example's name
    feats = [ImageSample(2, 0), PrevIterFeature(2, 1.5), Corner(2, 1), Corner(2, 2),
             PrevIterFeature(2, 0)]
    fm = FeatureMap(feats)
    assert fm.n_features == 5
    prog = FeatureProgram(fm.specs(), 2)
    assert prog.planes == [("sample", 0.0), ("usmooth", 1.5), ("user", "corner-reponse-sigma=1.00"),
                           ("user", "corner-reponse-sigma=2.00"), ("usmooth", 0.0)]
    assert list(prog.a) == [0, 1, 2, 3, 4] and prog.needs_u_planes
    assert sorted(prog.user_planes) == [2, 3]
    with pytest.raises(ValueError):
        FeatureMap([Corner(2, 1), Corner(2, 1)])                        # duplicates, as the reference
    with pytest.raises(ValueError):
        FeatureProgram([("user_plane", "x", None)], 2)
    with pytest.raises(TypeError):
        PrecomputedImageFeature(2, 1)               # abstract: `name` and `compute_plane` are the user's
