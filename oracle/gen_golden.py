"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
(`python -m oracle.gen_golden com_ray` / `prev_iter` regenerate those fixtures alone.)

Every array stored here is an output of the reference's own code imported through
oracle/ref_shim.py (its compiled C kernels, its feature classes, its FeatureMap, its
initializers, its `LevelSetMachineLearning.segment` loop).  The two third-party functions that
are absent from this container are plugged with the oracle's restatements and that is recorded
per fixture in the `notes` entry:

  skfmm.distance            -> oracle.skfmm_distance   (PARITY UNPINNED, lsml_oracle.c header)
  skimage find_contours     -> not used: BoundarySize / IsoperimetricRatio are left out of the
                               reference-generated feature fixtures

The GPU box has no /root/reference: tests read these files instead.
"""
import os
import pickle
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def gen_com_ray():
    """tests/golden/com_ray.npz: COMRaySample of the unmodified reference (image.py:124-171)."""
    warnings.simplefilter("ignore")
    from oracle import oracle as O
    from oracle.ref_shim import import_reference
    O.build()
    import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.feature.provided.image import COMRaySample
    from lsml.util.distance_transform import distance_transform as ref_dt
    rs = np.random.RandomState(11)
    f = {}
    cases = [((41, 37), (1.0, 1.0), 3), ((41, 37), (0.5, 2.0), 10),
             ((18, 21, 19), (1.0, 1.0, 1.0), 2), ((18, 21, 19), (2.0, 0.5, 1.0), 4)]
    for c, (shape, dx, n_samples) in enumerate(cases):
        ndim, dx = len(shape), np.asarray(dx)
        gg = np.indices(shape).astype(float)
        centre = np.asarray(shape) / 2.6          # off-centre: outward samples leave the grid
        r = np.sqrt(sum(((gg[a] - centre[a]) * dx[a]) ** 2 for a in range(ndim)))
        u = min(s * d for s, d in zip(shape, dx)) / 3.7 - r + 0.3 * rs.randn(*shape)
        img = rs.randn(*shape)
        dist, mask = ref_dt(u, band=3, dx=dx)
        feat = COMRaySample(ndim=ndim, sigma=0, n_samples=n_samples)
        X = feat(u=u, img=img, dist=dist, mask=mask, dx=dx)[mask]
        f.update({"u%d" % c: u, "img%d" % c: img, "mask%d" % c: mask, "dx%d" % c: dx,
                  "n_samples%d" % c: n_samples, "X%d" % c: X})
    f["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "com_ray.npz"),
                        notes="reference COMRaySample (scipy RegularGridInterpolator); band masks "
                              "from the reference distance_transform wrapper with skfmm plugged "
                              "by the oracle; dx are powers of two so that the centre of mass "
                              "is exact in every summation order", **f)
    print("com_ray.npz", os.path.getsize(os.path.join(OUT, "com_ray.npz")))


def gen_prev_iter():
    """tests/golden/prev_iter.npz: PrevIterFeature of the reference's gestalt2d example
    (examples/gestalt2d/prev_iter_feature.py), loaded from where it lies."""
    import importlib.util
    warnings.simplefilter("ignore")
    from oracle import oracle as O
    from oracle.ref_shim import import_reference, REFERENCE_ROOT
    O.build()
    import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.util.distance_transform import distance_transform as ref_dt
    path = REFERENCE_ROOT / "examples" / "gestalt2d" / "prev_iter_feature.py"
    mod_spec = importlib.util.spec_from_file_location("ref_prev_iter_feature", str(path))
    mod = importlib.util.module_from_spec(mod_spec)
    mod_spec.loader.exec_module(mod)
    rs = np.random.RandomState(21)
    f = {}
    cases = [((45, 38), (1.0, 1.0), (0, 1, 2.5)), ((45, 38), (0.7, 1.4), (3,)),
             ((17, 20, 16), (1.0, 1.0, 1.0), (1.5,))]
    for c, (shape, dx, sigmas) in enumerate(cases):
        ndim, dx = len(shape), np.asarray(dx)
        gg = np.indices(shape).astype(float)
        r = np.sqrt(sum(((gg[a] - shape[a] / 2.2) * dx[a]) ** 2 for a in range(ndim)))
        u = min(s * d for s, d in zip(shape, dx)) / 3.4 - r + 0.3 * rs.randn(*shape)
        dist, mask = ref_dt(u, band=3, dx=dx)
        cols = [mod.PrevIterFeature(ndim=ndim, sigma=sg)(u=u, dist=dist, mask=mask, dx=dx)[mask]
                for sg in sigmas]
        f.update({"u%d" % c: u, "mask%d" % c: mask, "dx%d" % c: dx,
                  "sigmas%d" % c: np.asarray(sigmas, dtype=float), "X%d" % c: np.stack(cols, axis=1)})
    f["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "prev_iter.npz"),
                        notes="PrevIterFeature of the reference's examples/gestalt2d "
                              "(scipy.ndimage.gaussian_filter(u, sigma)); band masks via the "
                              "reference distance_transform wrapper with skfmm plugged by the oracle",
                        **f)
    print("prev_iter.npz", os.path.getsize(os.path.join(OUT, "prev_iter.npz")))


def gen_fit():
    """tests/golden/fit.npz: the UNMODIFIED reference's LevelSetMachineLearning.fit (model.py:89-249,
    fit_job_handler.py) run end to end on a small in-memory dataset -- h5py replaced by
    oracle/fake_h5py.py, skfmm by the oracle's marcher -- and its segment() on a held-out image."""
    import tempfile
    warnings.simplefilter("ignore")
    from oracle import oracle as O
    from oracle import fake_h5py
    from oracle.ref_shim import import_reference
    O.build()
    lsml = import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.feature import (ImageSample, ImageEdgeSample, InteriorImageAverage,
                              InteriorImageVariation, Size, Moments, DistanceToCenterOfMass)
    from lsml.initializer import BallInitializer
    from lsml.data.dim2 import hamburger
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    imgs, segs = hamburger.make_dataset(N=15, n=41, random_state=np.random.RandomState(3))
    imgs, segs = np.asarray(imgs, dtype=float), np.asarray(segs, dtype=bool)
    feats = ([cls(ndim=2, sigma=s) for cls in (ImageSample, ImageEdgeSample, InteriorImageAverage,
                                               InteriorImageVariation) for s in (0, 2)] +
             [DistanceToCenterOfMass(ndim=2), Moments(ndim=2), Size(ndim=2)])
    out = {"imgs": imgs, "segs": segs}
    cwd = os.getcwd()
    from sklearn.ensemble import RandomForestRegressor
    out["probe"] = np.random.RandomState(0).randn(64, 14) * np.r_[[1.0] * 8, 5, 20, 10, 20, 10, 100]
    from lsml.initializer.initializer_base import InitializerBase

    def seed_ball_initialize(self, img, dx, seed):
        ind = np.indices(img.shape, dtype=float)
        for a in range(img.ndim):
            ind[a] = ind[a] * dx[a] - seed[a]
        return np.sqrt((ind ** 2).sum(axis=0)) < 5.0

    # A user initializer that runs on the host and USES the seed (default seeds =
    # center_of_mass_seeder, model.py:107): a ball of radius 5 around it.  Module-level so that
    # the reference can pickle the model; the same class is defined on lsml_b200's
    # InitializerBase in tests/test_fit_host_logic_cpu.py.
    global SeedBall
    SeedBall = type("SeedBall", (InitializerBase,), {"initialize": seed_ball_initialize,
                                                     "__module__": __name__})

    out["aniso_dx"] = np.array([1.0, 0.5])
    for tag, balance, seed in (("bal", True, 77), ("nobal", False, 5), ("forest", True, 9),
                               ("seedball", True, 21), ("aniso", True, 33)):
        fake_h5py.reset()
        tmp = tempfile.mkdtemp()
        os.chdir(tmp)
        try:
            init = SeedBall() if tag == "seedball" else BallInitializer(radius=6)
            model = lsml.LevelSetMachineLearning(features=feats, initializer=init, band=3)
            rs = np.random.RandomState(seed)
            if tag == "forest":      # the SAME RandomState for fit() and the forest, as in
                final = RandomForestRegressor(n_estimators=10, max_samples=200,   # examples/*/main.py
                                              random_state=rs)
            else:
                final = LinearRegression()
            model.fit('dataset.h5', imgs=list(imgs[:14]), segs=list(segs[:14]), max_iters=7,
                      random_state=rs,
                      regression_model_class=Pipeline,
                      regression_model_kwargs=dict(steps=[('s', StandardScaler()), ('r', final)]),
                      dx=(np.tile(out["aniso_dx"], (14, 1)) if tag == "aniso" else None),
                      balance_regression_targets=balance, temp_data_dir=tmp,
                      save_filename=os.path.join(tmp, 'm.pkl'), redirect_stdout_to_logfile=False,
                      validation_history_len=4, validation_history_tol=-1.0)
            fj = model.fit_job_handler
            for key in ("training", "validation", "testing"):
                idx = [int(k.split("-")[-1]) for k in fj.datasets_handler.datasets[key]]
                out["%s_%s_idx" % (tag, key)] = np.asarray(idx, dtype=np.int64)
            out["%s_step" % tag] = float(model.step)
            out["%s_iterations" % tag] = int(fj.iteration)
            out["%s_training_scores" % tag] = np.asarray(model.training_scores)
            out["%s_validation_scores" % tag] = np.asarray(model.validation_scores)
            out["%s_testing_scores" % tag] = np.asarray(model.testing_scores)
            for i in range(1, fj.iteration + 1):
                reg = model.regression_model(i)
                out["%s_mean_%d" % (tag, i)] = reg.steps[0][1].mean_
                out["%s_scale_%d" % (tag, i)] = reg.steps[0][1].scale_
                out["%s_probe_%d" % (tag, i)] = reg.predict(out["probe"])
                if tag != "forest":
                    out["%s_coef_%d" % (tag, i)] = reg.steps[1][1].coef_
                    out["%s_intercept_%d" % (tag, i)] = np.asarray(reg.steps[1][1].intercept_)
            out["%s_rng_after" % tag] = rs.randint(1 << 30)          # the caller's stream afterwards
            sdx = out["aniso_dx"] if tag == "aniso" else None
            out["%s_us" % tag] = model.segment(imgs[14], dx=sdx, verbose=False,
                                               iterate_until_validation_max=False)
            out["%s_us_valmax" % tag] = model.segment(imgs[14], dx=sdx, verbose=False)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "fit.npz"),
                        notes="the reference's own fit() (in-memory h5py stand-in, skfmm plugged by "
                              "the oracle): split, auto step, scores, per-iteration StandardScaler+"
                              "LinearRegression parameters, segment() of a held-out image; "
                              "random_state seeds 77 (balanced targets) and 5 (not balanced), 9 (a "
                              "10-tree forest sharing fit()'s RandomState), "
                              "14 training images 41x41, BallInitializer(radius=6), band 3, "
                              "max_iters 7, validation_history_len 4, tol -1 (never exits early)",
                        **out)
    print("fit.npz", os.path.getsize(os.path.join(OUT, "fit.npz")))


def gen_fit3d():
    """tests/golden/fit3d.npz: the reference's own fit() + segment() on small 3-D volumes
    (lsml/data/dim3/hamburger.py), tree-ensemble velocity -- the C3 workload family end to end."""
    import tempfile
    warnings.simplefilter("ignore")
    from oracle import oracle as O
    from oracle import fake_h5py
    from oracle.ref_shim import import_reference
    O.build()
    lsml = import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.feature import (ImageSample, ImageEdgeSample, InteriorImageAverage,
                              InteriorImageVariation, Size, Moments, DistanceToCenterOfMass)
    from lsml.initializer import BallInitializer
    from lsml.data.dim3 import hamburger
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    imgs, segs = hamburger.make_dataset(N=9, n=25, rad=[6, 9], rs=np.random.RandomState(8),
                                        verbose=False)
    imgs, segs = np.asarray(imgs, dtype=float), np.asarray(segs, dtype=bool)
    feats = ([cls(ndim=3, sigma=s) for cls in (ImageSample, ImageEdgeSample, InteriorImageAverage,
                                               InteriorImageVariation) for s in (0, 1.5)] +
             [DistanceToCenterOfMass(ndim=3), Moments(ndim=3, axes=(0, 1, 2), orders=(1, 2)),
              Size(ndim=3)])
    out = {"imgs": imgs, "segs": segs,
           "probe": np.random.RandomState(0).randn(64, 16) * np.r_[[1.0] * 8, 4, [12, 6] * 3, 300]}
    cwd = os.getcwd()
    fake_h5py.reset()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        rs = np.random.RandomState(12)
        model = lsml.LevelSetMachineLearning(features=feats, initializer=BallInitializer(radius=4),
                                             band=3)
        model.fit('dataset.h5', imgs=list(imgs[:8]), segs=list(segs[:8]), max_iters=5,
                  random_state=rs, regression_model_class=Pipeline,
                  regression_model_kwargs=dict(steps=[
                      ('s', StandardScaler()),
                      ('r', RandomForestRegressor(n_estimators=12, max_samples=300, random_state=rs))]),
                  datasets_split=([0, 1, 2, 3, 4], [5, 6], [7]), temp_data_dir=tmp,
                  save_filename=os.path.join(tmp, 'm.pkl'), redirect_stdout_to_logfile=False,
                  validation_history_len=3, validation_history_tol=-1.0)
        fj = model.fit_job_handler
        out["step"] = float(model.step)
        out["iterations"] = int(fj.iteration)
        for key in ("training", "validation", "testing"):
            out[key + "_scores"] = np.asarray(getattr(model, key + "_scores"))
        for i in range(1, fj.iteration + 1):
            out["probe_%d" % i] = model.regression_model(i).predict(out["probe"])
        out["rng_after"] = rs.randint(1 << 30)
        out["us"] = model.segment(imgs[8], verbose=False, iterate_until_validation_max=False)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "fit3d.npz"),
                        notes="the reference's own fit() on 8 volumes 25^3 (explicit split 5/2/1), "
                              "F = 16, StandardScaler + 12-tree forest sharing fit()'s RandomState(12), "
                              "BallInitializer(radius=4), band 3, 5 iterations; segment() of a 9th "
                              "volume; h5py stand-in, skfmm plugged by the oracle", **out)
    print("fit3d.npz", os.path.getsize(os.path.join(OUT, "fit3d.npz")))


def main():
    if sys.argv[1:] == ["fit3d"]:
        return gen_fit3d()
    if sys.argv[1:] == ["fit"]:
        return gen_fit()
    if sys.argv[1:] == ["com_ray"]:       # regenerate that fixture only
        return gen_com_ray()
    if sys.argv[1:] == ["prev_iter"]:
        return gen_prev_iter()
    warnings.simplefilter("ignore")
    from oracle import oracle as O
    from oracle.ref_shim import import_reference
    O.build()
    lsml = import_reference(skfmm_distance=O.skfmm_distance)
    from lsml.gradient import masked_gradient as mg
    from lsml.feature import (ImageSample, ImageEdgeSample, InteriorImageAverage,
                              InteriorImageVariation, Size, Moments, DistanceToCenterOfMass)
    from lsml.feature.feature_map import FeatureMap
    from lsml.initializer import BallInitializer, RandomBallInitializer
    from lsml.util.distance_transform import distance_transform as ref_dt
    from lsml.data.dim2 import hamburger
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- a1 / a2: C kernels
    rs = np.random.RandomState(1234)
    g = {}
    for ndim in (1, 2, 3):
        dims = rs.randint(12, 28, size=ndim)
        arr, nu = rs.randn(*dims), rs.randn(*dims)
        mask = rs.rand(*dims) > 0.4
        dx = rs.rand(ndim) + 0.5
        g["arr%d" % ndim], g["nu%d" % ndim], g["mask%d" % ndim], g["dx%d" % ndim] = arr, nu, mask, dx
        g["gmag_os%d" % ndim] = mg.gradient_magnitude_osher_sethian(arr, nu, mask, dx)
        g["gmag_os_full%d" % ndim] = mg.gradient_magnitude_osher_sethian(arr, nu)
        for norm in (0, 1):
            grads, gm = mg.gradient_centered(arr, mask, dx, normalize=bool(norm))
            g["gc%d_n%d_grads" % (ndim, norm)] = np.stack(grads)
            g["gc%d_n%d_gmag" % (ndim, norm)] = gm
    np.savez_compressed(os.path.join(OUT, "masked_gradient.npz"),
                        notes="reference C (masked_gradient.c) via its own ctypes wrapper", **g)

    # ---------------------------------------------------------------- images: reference generator
    imgs, segs = hamburger.make_dataset(N=6, n=51, random_state=np.random.RandomState(1234))
    imgs, segs = np.asarray(imgs, dtype=float), np.asarray(segs, dtype=bool)

    # ---------------------------------------------------------------- a4-a10: features
    f = {"imgs": imgs[:3], "segs": segs[:3]}
    rs = np.random.RandomState(7)
    feats2 = ([cls(ndim=2, sigma=s) for cls in (ImageSample, ImageEdgeSample, InteriorImageAverage,
                                                InteriorImageVariation) for s in (0, 2)] +
              [DistanceToCenterOfMass(ndim=2), Moments(ndim=2), Size(ndim=2)])
    fm2 = FeatureMap(feats2)
    for b, dx in enumerate([np.ones(2), np.ones(2), np.array([0.8, 1.25])]):
        gg = np.indices((51, 51)).astype(float)
        u = 11.0 - np.sqrt((gg[0] - 25.2) ** 2 + (gg[1] - 26.1) ** 2) + 0.3 * rs.randn(51, 51)
        dist, mask = ref_dt(u, band=3, dx=dx)
        f["u2_%d" % b], f["mask2_%d" % b], f["dx2_%d" % b] = u, mask, dx
        f["X2_%d" % b] = fm2(u=u, img=imgs[b], dist=dist, mask=mask, dx=dx)[mask]
    feats3 = ([cls(ndim=3, sigma=s) for cls in (ImageSample, ImageEdgeSample, InteriorImageAverage,
                                                InteriorImageVariation) for s in (0, 1.5)] +
              [DistanceToCenterOfMass(ndim=3), Moments(ndim=3, axes=(0, 1, 2), orders=(1, 2)),
               Size(ndim=3)])
    fm3 = FeatureMap(feats3)
    n3 = 24
    gg = np.indices((n3, n3, n3)).astype(float)
    r3 = np.sqrt(((gg - 11.6) ** 2).sum(0))
    img3 = (r3 < 7).astype(float) + 0.2 * rs.randn(n3, n3, n3)
    u3 = 6.0 - r3 + 0.2 * rs.randn(n3, n3, n3)
    dx3 = np.array([1.0, 1.0, 1.0])
    dist3, mask3 = ref_dt(u3, band=3, dx=dx3)
    f.update(img3=img3, u3=u3, mask3=mask3, dx3=dx3,
             X3=fm3(u=u3, img=img3, dist=dist3, mask=mask3, dx=dx3)[mask3])
    f["specs2"] = repr([("image_sample", 0), ("image_sample", 2), ("image_edge", 0),
                        ("image_edge", 2), ("band_mean", 0), ("band_mean", 2), ("band_std", 0),
                        ("band_std", 2), ("dist_to_com",), ("moments", (0, 1), (1, 2)), ("size",)])
    f["specs3"] = repr([("image_sample", 0), ("image_sample", 1.5), ("image_edge", 0),
                        ("image_edge", 1.5), ("band_mean", 0), ("band_mean", 1.5),
                        ("band_std", 0), ("band_std", 1.5), ("dist_to_com",),
                        ("moments", (0, 1, 2), (1, 2)), ("size",)])
    np.savez_compressed(os.path.join(OUT, "features.npz"),
                        notes="reference FeatureMap / feature classes; band masks from the "
                              "reference distance_transform wrapper with skfmm plugged by the "
                              "oracle (parity unpinned for the mask only; features pinned)", **f)

    # ---------------------------------------------------------------- initializers
    ini = {}
    u0, d0, m0 = BallInitializer(radius=10)(imgs[0], band=3, dx=np.ones(2))
    ini.update(ball_u0=u0, ball_mask=m0)
    u0, d0, m0 = BallInitializer(radius=4, location=[10, 12, 9])(
        np.zeros((20, 22, 18)), band=2, dx=np.array([1.0, 0.8, 1.2]))
    ini.update(ball3_u0=u0, ball3_mask=m0)
    img_ = (imgs[1] - imgs[1].mean()) / imgs[1].std()
    rb = RandomBallInitializer(random_state=np.random.RandomState(99))
    u0, d0, m0 = rb(img_, band=3, dx=np.ones(2))
    ini.update(rand_img=img_, rand_u0=u0, rand_mask=m0)
    np.savez_compressed(os.path.join(OUT, "initializers.npz"),
                        notes="reference BallInitializer / RandomBallInitializer through "
                              "InitializerBase.__call__ (u0 pinned; band mask via plugged skfmm)",
                        **ini)

    # ---------------------------------------------------------------- a15: the segment loop
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    sys.path.insert(0, ROOT)
    from lsml_b200.core.regressors import pack_trees

    seg_fix = {}
    band, step, n_iters = 3, 0.3, 6
    dx = np.ones(2)
    for kind in ("linear", "forest"):
        model = lsml.LevelSetMachineLearning(features=feats2, initializer=BallInitializer(radius=10),
                                             band=band)
        # train with the reference's own step (features -> fit -> update), no HDF5 bookkeeping
        img, seg = imgs[4], segs[4]
        img_ = (img - img.mean()) / img.std()
        gt = O.skfmm_distance(2 * seg.astype(float) - 1, dx=1.0)
        u, dist, mask = model.initializer(img_, band, dx=dx)
        regs = []
        rs2 = np.random.RandomState(5)
        for it in range(n_iters):
            X = model.feature_map(u=u, img=img_, dist=dist, mask=mask, dx=dx)[mask]
            if kind == "linear":
                reg = Pipeline([("s", StandardScaler()), ("r", LinearRegression())])
            else:
                reg = Pipeline([("s", StandardScaler()),
                                ("r", RandomForestRegressor(n_estimators=30, max_samples=200,
                                                            random_state=rs2))])
            reg.fit(X, np.asarray(gt)[mask])
            regs.append(reg)
            vel = np.zeros_like(u)
            vel[mask] = reg.predict(X)
            gm = mg.gradient_magnitude_osher_sethian(u, vel, mask, dx)
            u = u.copy()
            u[mask] += step * vel[mask] * gm[mask]
            dist, mask = ref_dt(u, band=band, dx=dx)
        model._is_fitted = True
        model.fit_job_handler = types.SimpleNamespace(
            step=step, iteration=n_iters,
            _load_regression_model=lambda iteration, regs=regs: regs[iteration - 1])
        us = model.segment(img, dx=dx, verbose=False, iterate_until_validation_max=False)
        seg_fix["%s_img" % kind] = img
        seg_fix["%s_seg" % kind] = seg
        seg_fix["%s_us" % kind] = us
        for i, reg in enumerate(regs):
            sc = reg.steps[0][1]
            seg_fix["%s_mean_%d" % (kind, i)] = sc.mean_
            seg_fix["%s_scale_%d" % (kind, i)] = sc.scale_
            fin = reg.steps[1][1]
            if kind == "linear":
                seg_fix["linear_coef_%d" % i] = fin.coef_
                seg_fix["linear_intercept_%d" % i] = np.asarray(fin.intercept_)
            else:
                nodes, bases, root = pack_trees([e.tree_ for e in fin.estimators_], X.shape[1])
                seg_fix["forest_nodes_%d" % i] = nodes
                seg_fix["forest_bases_%d" % i] = bases
                seg_fix["forest_root_%d" % i] = root
    seg_fix.update(band=band, step=step, n_iters=n_iters)
    np.savez_compressed(os.path.join(OUT, "segment.npz"),
                        notes="iterates `us` returned by the reference's own "
                              "LevelSetMachineLearning.segment (model.py:283-411) with skfmm "
                              "plugged by the oracle; regressors stored as exported parameters",
                        **seg_fix)
    gen_com_ray()
    gen_prev_iter()
    gen_fit()
    gen_fit3d()
    for name in sorted(os.listdir(OUT)):
        print(name, os.path.getsize(os.path.join(OUT, name)))


if __name__ == "__main__":
    main()
