"""``LevelSetMachineLearning`` -- host-side mirror of lsml/core/model.py on the device engine.

Kept from the reference: the constructor (:28-29), ``segment`` (:283-411) with its arguments,
return value (all iterates ``us`` as a float64 array, optionally the scores) and error
behaviour, ``fit`` with the reference's keyword arguments (:89-111), ``save``/``load``, ``step``,
``regression_model(i)`` and the score accessors.

Changed on purpose (out of scope for the hot path):
no HDF5 dataset file, no ``tmp.h5`` state file, no per-iteration pickle of the regressor to the
current directory and no child process for the fit -- ``fit`` takes ``imgs``/``segs`` in
memory, keeps the level sets of ALL examples resident in HBM as one batch, trains the
scikit-learn regressor on the host each iteration and keeps the fitted regressors in memory.
``segment`` never touches the disk.
"""
import copy
import functools
import pickle

import numpy
import torch

from .. import _lib
from ..feature.feature_map import FeatureMap
from ..initializer.initializer_base import InitializerBase
from ..initializer.provided import BallInitializer
from ..initializer.seed import center_of_mass_seeder
from ..score_functions import jaccard
from .engine import BandEngine
from .regressors import export_regressor

DEFAULT_MODEL_FILENAME = 'LSML-model.pkl'


class ModelNotFit(Exception):
    """ Raised when a fitted-model attribute is used before ``fit`` (lsml/core/exception.py) """


class ModelAlreadyFit(Exception):
    """ Raised when ``fit`` is called twice (lsml/core/exception.py) """


def _ball_u0(eng, centres, radii):
    """u0 = 2*ball - 1 for every item, built in HBM (centres: (batch, ndim) physical units)."""
    import ctypes
    pad = 3 - eng.ndim
    c3 = numpy.zeros((eng.batch, 3))
    c3[:, pad:] = centres
    cd = torch.from_numpy(c3).to(eng.device)
    rd = torch.from_numpy(numpy.asarray(radii, dtype=numpy.float64)).to(eng.device)
    u0 = torch.empty((eng.batch,) + eng.shape, dtype=torch.float64, device=eng.device)
    with torch.cuda.device(eng.device):
        _lib.check(_lib.lib.lsml_ball_init(
            ctypes.c_int(eng.ndim), eng._shape_c, _lib.c_i64(eng.batch), eng._dx_c, _lib.ptr(u0),
            _lib.c_p(0), ctypes.c_double(0.0), _lib.ptr(cd), _lib.ptr(rd), _lib.current_stream()),
            "ball_init")
    return u0


def overlap_scores(eng, segs_dev, scorer=jaccard):
    """Score of {u>0} vs seg per item (fit_job_handler.py:258-291 calls ``model.scorer(u, seg)``
    per example).  ``jaccard`` and ``dice`` (score_functions.py:4-30) come from exact device
    counts of the intersection and the union -- no level set leaves the GPU; any other scorer
    is called on a host copy of each example's ``u``, as the reference does."""
    from ..score_functions import dice
    if scorer is not jaccard and scorer is not dice:
        u = eng.u.cpu().numpy()
        seg = segs_dev.cpu().numpy().astype(bool).reshape(u.shape)
        return numpy.array([scorer(u[b], seg[b]) for b in range(eng.batch)], dtype=float)
    counts = torch.zeros((eng.batch, 2), dtype=torch.int64, device=eng.device)
    with torch.cuda.device(eng.device):
        _lib.check(_lib.lib.lsml_overlap_counts(
            _lib.c_i64(eng.voxels), _lib.c_i64(eng.batch), _lib.ptr(eng.u), _lib.ptr(segs_dev),
            _lib.ptr(counts), _lib.current_stream()), "overlap_counts")
    c = counts.cpu().numpy().astype(numpy.float64)
    with numpy.errstate(invalid="ignore", divide="ignore"):
        j = numpy.where(c[:, 1] == 0, 1.0, c[:, 0] / c[:, 1])
    return j if scorer is jaccard else 2.0 * j / (j + 1.0)


def balance_mask(arr, random_state=None):
    """lsml/util/balance_mask.py:4-44: a mask under which ``arr`` has as many positive as
    negative entries.  The larger sign class is down-sampled at random (one ``choice`` call, so
    the RNG stream matches the reference); zeros always stay; an array that cannot be balanced
    (no positive or no negative entry) or is already balanced keeps everything."""
    rs = numpy.random.RandomState() if random_state is None else random_state
    pos, neg = arr > 0, arr < 0
    n_pos, n_neg = int(pos.sum()), int(neg.sum())
    if n_pos == 0 or n_neg == 0 or n_pos == n_neg:
        return numpy.ones(arr.shape, dtype=bool)
    major, n_keep = (pos, n_neg) if n_pos > n_neg else (neg, n_pos)
    keep = rs.choice(numpy.where(major)[0], replace=False, size=n_keep)
    mask = ~major
    mask[keep] = True
    return mask


def split_datasets(n, datasets_split, subset_size, random_state):
    """Indices of the training / validation / testing examples (datasets_handler.py:204-362):
    three index lists are used as given; three probabilities assign at random with the
    reference's RNG stream and quirk -- a permutation is drawn
    (``choice(keys, replace=False, size=subset_size)``) but only its LENGTH is used, so the
    examples that take part are the first ``subset_size`` ones, each placed by one multinomial
    draw."""
    if all(numpy.iterable(s) for s in datasets_split):
        return tuple(numpy.asarray(s, dtype=int) for s in datasets_split)
    if not all(isinstance(s, float) for s in datasets_split):
        raise ValueError("`training`, `validation`, and `testing` should be "
                         "all floats or all list of ints")
    if subset_size is not None and subset_size > n:
        raise ValueError("`subset_size` must be <= `len(keys)`")
    n_sub = n if subset_size is None else int(subset_size)
    random_state.choice(n, replace=False, size=n_sub)
    indicators = random_state.multinomial(n=1, pvals=datasets_split, size=n_sub)
    assign = indicators.argmax(axis=1)
    idx = numpy.arange(n_sub)
    return idx[assign == 0], idx[assign == 1], idx[assign == 2]


class LevelSetMachineLearning:

    def __init__(self, features, initializer, scorer=jaccard, band=3, normalize_imgs=True):
        """ Same parameters as the reference (model.py:28-66). """
        self.feature_map = FeatureMap(features=features)
        if not isinstance(initializer, InitializerBase):
            msg = ("`initializer` should be a class derived from "
                   "`level_set_machine_learning.initializer.InitializerBase`")
            raise ValueError(msg)
        self.initializer = initializer
        self.scorer = scorer
        self.band = band
        self.normalize_imgs = normalize_imgs
        self.training_data = None
        self.validation_data = None
        self.testing_data = None
        self.regression_models = []      # fitted host regressors, one per iteration
        self.scores = {}                 # dataset key -> (n_iters+1, n_examples) array
        self.iteration = 0
        self._step = None
        self._is_fitted = False
        self._device_models = {}

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_regressors(cls, features, initializer, regressors, step, band=3,
                        normalize_imgs=True, scorer=jaccard):
        """A fitted model from regressors trained elsewhere (one per iteration)."""
        model = cls(features, initializer, scorer=scorer, band=band,
                    normalize_imgs=normalize_imgs)
        model.regression_models = list(regressors)
        model.iteration = len(model.regression_models)
        model._step = float(step)
        model._is_fitted = True
        return model

    def _initial_level_sets(self, eng, imgs_norm_first_values=None, seeds=None):
        """u0 for every item of ``eng`` (InitializerBase.__call__, initializer_base.py:19-61)."""
        init = self.initializer
        if type(init) is BallInitializer:
            centre, radius = init.device_ball(eng.shape, eng.dx)
            return _ball_u0(eng, numpy.tile(centre, (eng.batch, 1)), [radius] * eng.batch)
        # generic initializers run their host `initialize` on the (normalised) image
        imgs = eng.img.cpu().numpy()
        u0 = numpy.empty(imgs.shape)
        for b in range(eng.batch):
            seed = (numpy.array(eng.shape) / 2 if seeds is None else numpy.array(seeds[b]))
            m = init.initialize(img=imgs[b], dx=eng.dx, seed=seed)
            if not isinstance(m, numpy.ndarray) or m.dtype != bool or m.shape != eng.shape:
                raise TypeError("initializer must return a bool array of the image shape")
            u0[b] = 2 * m.astype(float) - 1
        return u0

    # ------------------------------------------------------------------ fit
    def fit(self, data_filename=None, regression_model_class=None, regression_model_kwargs=None,
            balance_regression_targets=True, datasets_split=(0.6, 0.2, 0.2), dx=None, imgs=None,
            max_iters=100, random_state=None, save_filename=None, seeds=center_of_mass_seeder,
            segs=None,
            step=None, subset_size=None, temp_data_dir=None, validation_history_len=5,
            validation_history_tol=0.0, redirect_stdout_to_logfile=False, device="cuda",
            verbose=False):
        """ Fit the per-iteration velocity regressors (reference model.py:89-249,
        fit_job_handler.py:193-562).  ``imgs`` / ``segs`` must be given in memory and share one
        shape; see the module docstring for what is deliberately not reproduced. """
        if self._is_fitted:
            raise ModelAlreadyFit("This model has already been fit")
        if imgs is None or segs is None:
            raise NotImplementedError(
                "lsml_b200.fit works on in-memory `imgs`/`segs`; the HDF5 dataset file of the "
                "reference (datasets_handler.py) is outside the hot path and not reproduced")
        if regression_model_class is None:
            raise ValueError("`regression_model_class` is required")
        kwargs = dict(regression_model_kwargs or {})
        rs = random_state if random_state is not None else numpy.random.RandomState()
        # input validation of datasets_handler.py:131-170 (same exceptions and messages)
        if len(imgs) != len(segs):
            msg = "Mismatch in number of examples: imgs ({}), segs ({})"
            raise ValueError(msg.format(len(imgs), len(segs)))
        ndim_ = numpy.ndim(imgs[0])
        for i, (img, seg) in enumerate(zip(imgs, segs)):
            if img.dtype != float:
                raise TypeError("imgs[{}] (dtype {}) was not float".format(i, img.dtype))
            if seg.dtype != bool:
                raise TypeError("seg[{}] (dtype {}) was not bool".format(i, seg.dtype))
            if img.ndim != ndim_:
                msg = "imgs[{}] (ndim={}) did not have correct dimensions ({})"
                raise ValueError(msg.format(i, img.ndim, ndim_))
            if seg.ndim != ndim_:
                msg = "segs[{}] (ndim={}) did not have correct dimensions ({})"
                raise ValueError(msg.format(i, seg.ndim, ndim_))
            if img.shape != seg.shape:
                msg = "imgs[{}] shape {} does not match segs[{}] shape {}"
                raise ValueError(msg.format(i, img.shape, i, seg.shape))
        if dx is not None and numpy.ndim(dx) == 2 and numpy.shape(dx) != (len(imgs), ndim_):
            msg = "`dx` was shape {} but should be shape {}"
            raise ValueError(msg.format(numpy.shape(dx), (len(imgs), ndim_)))
        if len({numpy.shape(i) for i in imgs}) != 1:
            raise NotImplementedError(
                "lsml_b200.fit keeps all examples in one device batch: the images must share "
                "one shape (the reference stores them one by one in HDF5)")
        imgs = numpy.ascontiguousarray(numpy.stack([numpy.asarray(i, dtype=float) for i in imgs]))
        segs = numpy.ascontiguousarray(numpy.stack([numpy.asarray(s, dtype=bool) for s in segs]))
        n, shape = imgs.shape[0], imgs.shape[1:]
        dxv = numpy.ones(len(shape)) if dx is None else numpy.asarray(dx, dtype=float)
        if dxv.ndim == 2:
            if not (dxv == dxv[0]).all():
                raise NotImplementedError("per-example dx must be identical within a batch")
            dxv = dxv[0]
        tr, va, te = split_datasets(n, datasets_split, subset_size, rs)
        self._split = dict(training=tr, validation=va, testing=te)

        specs = self.feature_map.specs()
        eng = BandEngine(shape, n, dx=dxv, band=self.band, specs=specs, device=device)
        eng.load_images(imgs, normalize=self.normalize_imgs)
        segs_dev = torch.from_numpy(segs.view(numpy.uint8)).to(eng.device)
        # ground-truth signed distance skfmm.distance(2*seg-1, dx) (datasets_handler.py:194):
        # the band-rebuild solver with no narrow band
        gt = BandEngine(shape, n, dx=dxv, band=0.0, specs=[], device=device)
        gt.set_level_sets(2.0 * segs.astype(float) - 1.0)
        dist_gt = gt.distance().reshape(-1)
        del gt
        if callable(seeds):      # e.g. center_of_mass_seeder (the reference's default, model.py:107)
            import types
            seeds = [seeds(types.SimpleNamespace(img=imgs[i], seg=segs[i], dx=dxv, dist=None))
                     for i in range(n)]
        eng.set_level_sets(self._initial_level_sets(eng, seeds=seeds))

        def band_targets():
            return dist_gt[eng.band_idx[:eng.n_band]].cpu().numpy()

        def rows_of(items):
            off = eng.item_offsets.cpu().numpy()
            sel = [numpy.arange(off[i], off[i + 1]) for i in items]
            return numpy.concatenate(sel) if sel else numpy.zeros(0, dtype=int)

        if step is None:   # fit_job_handler.py:226-256
            t = numpy.abs(band_targets()[rows_of(numpy.r_[tr, va])])
            step = float(dxv.min() / t.max())
        self._step = float(step)

        def record_scores():
            s = overlap_scores(eng, segs_dev, self.scorer)
            for key, items in self._split.items():
                self.scores.setdefault(key, []).append(s[items])

        record_scores()
        for self.iteration in range(1, max_iters + 1):
            X = eng.compute_features().cpu().numpy()
            y = band_targets()
            rows = rows_of(tr)
            Xtr, ytr = X[rows], y[rows]
            if balance_regression_targets:
                # per training example, as fit_job_handler.py:324-375
                off = eng.item_offsets.cpu().numpy()
                keep = []
                for i in tr:
                    yi = y[off[i]:off[i + 1]]
                    keep.append(numpy.arange(off[i], off[i + 1])[balance_mask(yi, rs)])
                keep = numpy.concatenate(keep) if keep else numpy.zeros(0, dtype=int)
                Xtr, ytr = X[keep], y[keep]
            # The reference fits in a forked child process and pickles the result
            # (fit_job_handler.py:398-424): the objects in `regression_model_kwargs` (estimator
            # instances of a Pipeline, a shared RandomState ...) are COPIES there -- the parent's
            # are never fitted or advanced, and every iteration's model is its own snapshot.
            model = regression_model_class(**copy.deepcopy(kwargs))
            model.fit(Xtr, ytr)
            self.regression_models.append(model)
            eng.step(model, self._step)
            record_scores()
            if verbose:
                print("iteration {:03d}: mean validation score {:.5f}".format(
                    self.iteration, float(numpy.mean(self.scores["validation"][-1]))
                    if len(va) else float("nan")))
            if save_filename:
                self._is_fitted = True
                self.save(save_filename)
                self._is_fitted = False
            # early exit on the trend of the validation score over the last
            # `validation_history_len` scores, current one included (fit_job_handler.py:523-562:
            # possible from iteration validation_history_len - 1 on, the score at
            # initialisation counts)
            if self.iteration >= validation_history_len - 1 and len(va):
                hist = numpy.array([numpy.mean(s) for s in
                                    self.scores["validation"][-validation_history_len:]])
                xs = numpy.c_[numpy.arange(validation_history_len),
                              numpy.ones(validation_history_len)]
                slope = numpy.linalg.lstsq(xs, hist, rcond=None)[0][0]
                if slope < validation_history_tol:
                    break
        eng.check_converged()
        self.scores = {k: numpy.array(v) for k, v in self.scores.items()}
        self._is_fitted = True
        if save_filename:
            self.save(save_filename)

    def save(self, filename):
        with open(filename, 'wb') as f:
            pickle.dump(self, f)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_device_models"] = {}
        return state

    @staticmethod
    def load(filename):
        with open(filename, 'rb') as f:
            return pickle.load(f)

    # ------------------------------------------------------------------ fitted-model API
    def _requires_fit(method):
        @functools.wraps(method)
        def method_wrapped(self, *args, **kwargs):
            if not self._is_fitted:
                raise ModelNotFit("This model has not been fit yet")
            return method(self, *args, **kwargs)
        return method_wrapped

    @property
    @_requires_fit
    def step(self):
        return self._step

    @property
    @_requires_fit
    def training_scores(self):
        return self.scores["training"]

    @property
    @_requires_fit
    def validation_scores(self):
        return self.scores["validation"]

    @property
    @_requires_fit
    def testing_scores(self):
        return self.scores["testing"]

    @_requires_fit
    def regression_model(self, iteration):
        """ The fitted regressor of iteration ``iteration`` (1-based, reference model.py:441) """
        return self.regression_models[iteration - 1]

    def _device_model(self, i, device):
        key = (i, str(device))
        if key not in self._device_models:
            self._device_models[key] = export_regressor(self.regression_models[i], device)
        return self._device_models[key]

    def _n_iters(self, iterate_until_validation_max):
        if iterate_until_validation_max and len(self.scores.get("validation", [])):
            return int(numpy.asarray(self.validation_scores).mean(axis=1).argmax())
        return self.iteration

    @_requires_fit
    def segment(self, img, dx=None, verbose=True, on_iterate=None, return_scores=False,
                seg=None, iterate_until_validation_max=True, device="cuda"):
        """ Segment `img` (reference model.py:283-411).  Returns ``us`` of shape
        ``(n_iters+1,) + img.shape`` (float64) and, with ``return_scores``, the scores. """
        img = numpy.asarray(img) if not torch.is_tensor(img) else img
        ndim = img.ndim
        dx = numpy.ones(ndim) if dx is None else numpy.asarray(dx, dtype=float)
        if dx.shape[0] != ndim:
            raise ValueError("`dx` has incorrect number of elements.")
        if on_iterate:
            if not isinstance(on_iterate, list):
                on_iterate = [on_iterate]
            if not all([callable(func) for func in on_iterate]):
                raise TypeError("All on_iterate items must be callable")
        scores = None
        if return_scores:
            if seg is None:
                raise ValueError("`seg` must be provided to when `return_scores` is True")
            scores = []
        n_iters = self._n_iters(iterate_until_validation_max)
        eng = self.segment_batch(img[None], dx=dx, n_iters=0, device=device, _return_engine=True)
        u0 = eng.u[0].cpu().numpy()
        us = numpy.zeros((n_iters + 1,) + tuple(img.shape))
        us[0] = u0

        def after(i):
            if scores is not None:
                scores.append(self.scorer(us[i], seg))
            if on_iterate:
                for func in on_iterate:
                    func(i, us[i])

        print_string = "Iter: {:02d}"
        if verbose:
            print(print_string.format(0))
        after(0)
        for i in range(n_iters):
            nb = eng.n_band
            us[i + 1] = us[i]
            if nb:                # `if mask.any()` (model.py:381)
                eng.step(self._device_model(i, eng.device), self._step)
                # the sparse record of this iteration: band indices and their new values
                idx, val = eng.last_record(nb)
                us[i + 1].ravel()[idx.cpu().numpy()] = val.cpu().numpy()
            if verbose:
                print(print_string.format(i + 1))
            after(i + 1)
        eng.check_converged()
        if return_scores:
            return us, numpy.array(scores)
        return us

    @_requires_fit
    def segment_batch(self, imgs, dx=None, n_iters=None, iterate_until_validation_max=True,
                      device="cuda", _return_engine=False):
        """ Evolve a whole batch of same-shape images at once, state resident in HBM.
        Returns ``(u_final, mask_final)`` as CUDA tensors ``(batch, *shape)`` -- the
        segmentation is ``u_final > 0``. """
        shape = tuple(imgs.shape[1:])
        dx = numpy.ones(len(shape)) if dx is None else numpy.asarray(dx, dtype=float)
        eng = BandEngine(shape, imgs.shape[0], dx=dx, band=self.band,
                         specs=self.feature_map.specs(), device=device)
        eng.load_images(imgs, normalize=self.normalize_imgs)
        eng.set_level_sets(self._initial_level_sets(eng))
        if n_iters is None:
            n_iters = self._n_iters(iterate_until_validation_max)
        for i in range(n_iters):
            eng.step(self._device_model(i, eng.device), self._step)
        eng.check_converged()
        if _return_engine:
            return eng
        return eng.u, eng.mask.bool()
