"""Export host-trained scikit-learn regressors to flat device arrays.

The reference calls ``regression_model.predict(features[mask])`` with whatever class the user
passed to ``fit`` (lsml/core/model.py:388, fit_job_handler.py:412-428, 504); the shipped examples
use ``Pipeline(StandardScaler, LinearRegression)`` (examples/hamburger2d/main.py:35-41) and
``Pipeline(StandardScaler, RandomForestRegressor)`` (examples/gestalt2d/main.py:46-55).
Training stays on the host; only ``predict`` runs on the GPU, with scikit-learn's arithmetic
(see lsml_b200/csrc/regress.cu).  Unknown model classes raise ``TypeError``: there is no
host-callback fallback.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _lib


def _dev(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype)))
    return t.to(device)


class DeviceRegressor:
    """Base: optional StandardScaler parameters + ``predict_into``."""

    def __init__(self, n_features, scaler, device):
        self.n_features = int(n_features)
        self.device = device
        if scaler is not None:
            mean, scale = scaler
            self.mean = _dev(mean, device, np.float64)
            self.scale = _dev(scale, device, np.float64)
        else:
            self.mean = self.scale = None

    def predict_into(self, X, n_band, out):
        raise NotImplementedError

    def predict_band(self, src, n_band, out):
        """Velocity of every band voxel with the feature program ``src`` (a
        :class:`lsml_b200._lib.BandSource`) as the input: no materialised matrix."""
        raise NotImplementedError

    def predict(self, X):
        """Convenience: X CUDA tensor (B, F) -> CUDA tensor (B,)."""
        X = X.contiguous()
        n_band = torch.tensor([X.shape[0]], dtype=torch.int64, device=X.device)
        out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
        with torch.cuda.device(X.device):
            self.predict_into(X, n_band, out)
        return out


class LinearDeviceRegressor(DeviceRegressor):
    def __init__(self, coef, intercept, scaler, device):
        coef = np.asarray(coef, dtype=np.float64).ravel()
        super().__init__(coef.size, scaler, device)
        self.coef = _dev(coef, device)
        self.intercept = float(np.asarray(intercept).ravel()[0]) if np.ndim(intercept) else float(intercept)

    def predict_into(self, X, n_band, out):
        _lib.check(_lib.lib.lsml_linear_predict(
            _lib.ptr(X), _lib.ptr(n_band), ctypes.c_int(self.n_features), _lib.ptr(self.mean),
            _lib.ptr(self.scale), _lib.ptr(self.coef), ctypes.c_double(self.intercept),
            _lib.ptr(out), _lib.current_stream()), "linear_predict")

    def predict_band(self, src, n_band, out):
        _lib.check(_lib.lib.lsml_linear_predict_band(
            ctypes.byref(src), _lib.ptr(n_band), _lib.ptr(self.mean),
            _lib.ptr(self.scale), _lib.ptr(self.coef), ctypes.c_double(self.intercept),
            _lib.ptr(out), _lib.current_stream()), "linear_predict_band")


def _round_down_to_f32(thr):
    """Largest float32 <= thr, so that `x_f32 <= thr_f64` <=> `x_f32 <= thr_f32`."""
    t32 = thr.astype(np.float32)
    too_big = t32.astype(np.float64) > thr
    t32[too_big] = np.nextafter(t32[too_big], np.float32(-np.inf))
    return t32


def pack_trees(trees, n_features):
    """sklearn ``Tree`` objects -> (nodes uint32[n,2], tree_base int32[T+1], root_is_leaf uint8[T]).

    Nodes are renumbered breadth-first per tree so that the two children of an internal node are
    adjacent; see regress.cu for the 8-byte layout.
    """
    all_nodes, bases, root_leaf = [], [], []
    base = 0
    for tree in trees:
        left, right = tree.children_left, tree.children_right
        feat, thr = tree.feature, tree.threshold
        val = tree.value.reshape(tree.node_count, -1)[:, 0].astype(np.float64)
        n = tree.node_count
        if n >= (1 << 22):
            raise ValueError("tree with %d nodes exceeds the 22-bit child index" % n)
        new_id = np.full(n, -1, dtype=np.int64)
        order = [0]
        new_id[0] = 0
        head = 0
        while head < len(order):
            o = order[head]
            head += 1
            if left[o] != -1:
                new_id[left[o]] = len(order)
                order.append(left[o])
                new_id[right[o]] = len(order)
                order.append(right[o])
        order = np.asarray(order)
        is_leaf = left[order] == -1
        nodes = np.zeros((n, 2), dtype=np.uint32)
        # leaves: the float64 value as (lo, hi) words
        nodes[is_leaf] = val[order[is_leaf]].view(np.uint32).reshape(-1, 2)
        io = order[~is_leaf]
        if (feat[io] >= n_features).any() or (feat[io] > 254).any():
            raise ValueError("tree uses a feature index outside the feature map")
        t32 = _round_down_to_f32(thr[io])
        lnew = new_id[left[io]]
        meta = ((feat[io].astype(np.uint32) << np.uint32(24)) |
                ((left[left[io]] == -1).astype(np.uint32) << np.uint32(23)) |
                ((left[right[io]] == -1).astype(np.uint32) << np.uint32(22)) |
                lnew.astype(np.uint32))
        nodes[~is_leaf, 0] = t32.view(np.uint32)
        nodes[~is_leaf, 1] = meta
        all_nodes.append(nodes)
        bases.append(base)
        root_leaf.append(1 if left[0] == -1 else 0)
        base += n
    bases.append(base)            # sentinel: tree t has bases[t+1] - bases[t] nodes
    return (np.concatenate(all_nodes, axis=0), np.asarray(bases, dtype=np.int32),
            np.asarray(root_leaf, dtype=np.uint8))


def pack_supernodes(nodes, bases, root_is_leaf):
    """The single-node packing of :func:`pack_trees` -> the layout the kernels walk (regress.cu):

    ``snodes`` uint32[S, 8]: one 32-byte super node per internal node of EVEN depth, holding that
    node and its two children -- thresholds (node, left, right) as float32 bits, then
    ``feature0 | featureL << 8 | featureR << 16 | flags << 24`` (flags bit 0 / 1: the left / right
    child is a leaf; bits 2..5: grandchild LL, LR, RL, RR is a leaf), then four references: per
    grandchild the index of its super node or of its leaf value; for a leaf child ``ref[2*side]``
    is that child's leaf index.  ``leaves`` float64[L], ``root_ref`` uint32[T] (root super node, or
    the leaf index of a single-leaf tree).  Two tree levels per fetch; same decisions, same leaf,
    same float64 value as the single-node walk."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    bases = np.asarray(bases, dtype=np.int64)
    root_is_leaf = np.asarray(root_is_leaf, dtype=np.uint8)
    n, T = nodes.shape[0], len(root_is_leaf)
    meta = nodes[:, 1].astype(np.int64)
    tree_of = np.repeat(np.arange(T), np.diff(bases))
    left = bases[tree_of] + (meta & 0x3fffff)           # valid for internal nodes only
    lleaf, rleaf = (meta >> 23) & 1, (meta >> 22) & 1
    feat = (meta >> 24) & 0xff
    is_leaf = np.zeros(n, dtype=bool)
    depth = np.zeros(n, dtype=np.int64)
    roots = bases[:-1]
    is_leaf[roots] = root_is_leaf != 0
    frontier = roots[root_is_leaf == 0]
    while frontier.size:
        L = left[frontier]
        is_leaf[L], is_leaf[L + 1] = lleaf[frontier] != 0, rleaf[frontier] != 0
        depth[L] = depth[L + 1] = depth[frontier] + 1
        kids = np.concatenate([L, L + 1])
        frontier = kids[~is_leaf[kids]]
    sup = ~is_leaf & (depth % 2 == 0)
    sid = np.cumsum(sup) - 1
    lid = np.cumsum(is_leaf) - 1
    leaves = np.ascontiguousarray(nodes[is_leaf]).view(np.float64).ravel().copy()
    S = np.flatnonzero(sup)
    out = np.zeros((S.size, 8), dtype=np.uint32)
    out[:, 0] = nodes[S, 0]
    ff = feat[S].copy()
    flags = lleaf[S] | (rleaf[S] << 1)
    for side in (0, 1):
        child = left[S] + side
        cl = is_leaf[child]
        # leaf child: its leaf index in ref[2 * side]
        out[cl, 4 + 2 * side] = lid[child[cl]]
        ci = child[~cl]
        out[~cl, 1 + side] = nodes[ci, 0]
        ff[~cl] |= feat[ci] << (8 * (1 + side))
        for gs in (0, 1):
            g = left[ci] + gs
            gl = is_leaf[g]
            out[~cl, 4 + 2 * side + gs] = np.where(gl, lid[g], sid[g])
            flags[~cl] |= gl.astype(np.int64) << (2 + 2 * side + gs)
    out[:, 3] = (ff | (flags << 24)).astype(np.uint32)
    root_ref = np.where(root_is_leaf != 0, lid[roots], sid[roots]).astype(np.uint32)
    return out, leaves, root_ref, root_is_leaf.copy()


RANKED_MAX_TREE_WORDS = 8192      # a tree must fit the kernel's shared-memory tree buffer
RANKED_TREE_ALIGN = 4             # trees start on 16-byte boundaries (vector copies)


def pack_ranked(nodes, bases, root_is_leaf, n_features):
    """The single-node packing of :func:`pack_trees` -> the RANK-QUANTISED layout walked from
    shared memory (regress.cu: forest_ranked_kernel), or ``None`` when the forest does not fit it.

    A tree only ever asks ``x[f] <= thr``; with ``S_f`` the sorted distinct (float32) thresholds
    the forest uses on feature ``f`` and ``rank_f(x) = #{s in S_f : s < x}``, that is
    ``rank_f(x) <= j`` for the node whose threshold is ``S_f[j]`` -- exactly, for every float32
    ``x`` (NaN ranks above everything: ``!(x <= thr)`` sends it right, as the other kernels do).
    A voxel's features are therefore reduced ONCE to ranks, and an internal node shrinks to one
    32-bit word, fields from the top::

        feature (fbits) | rank j (rbits) | offset of the right child in words | bit 1: the right
        child is a leaf | bit 0: the left child is a leaf

    so that, with the voxel's rank stored as ``feature << (32 - fbits) | rank << (32 - fbits -
    rbits)``, "go right" is ONE unsigned comparison of two words (the fields below the rank
    cannot flip it), and ``word & offset mask`` IS the byte offset of the right child.

    Trees are laid out in preorder: the left child is the next word, a leaf is its float64
    value in two words.  Returns a dict: ``words`` uint32, ``tree_start`` int32[T+1] (word
    offsets, multiples of 4), ``thr_tab`` float32 (the S_f back to back), ``thr_off``
    int32[F+1], ``fbits``, ``rbits``, ``max_tree_words``.
    """
    nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    bases = np.asarray(bases, dtype=np.int64)
    root_is_leaf = np.asarray(root_is_leaf, dtype=np.uint8)
    n, T = nodes.shape[0], len(root_is_leaf)
    if T == 0 or n_features < 1:
        return None
    meta = nodes[:, 1].astype(np.int64)
    tree_of = np.repeat(np.arange(T), np.diff(bases))
    left = bases[tree_of] + (meta & 0x3fffff)           # valid for internal nodes only
    lleaf, rleaf = (meta >> 23) & 1, (meta >> 22) & 1
    # leaf flags and depth of every node, level by level over all trees at once
    is_leaf = np.zeros(n, dtype=bool)
    roots = bases[:-1]
    is_leaf[roots] = root_is_leaf != 0
    levels = []
    frontier = roots[root_is_leaf == 0]
    while frontier.size:
        levels.append(frontier)
        L = left[frontier]
        is_leaf[L], is_leaf[L + 1] = lleaf[frontier] != 0, rleaf[frontier] != 0
        kids = np.concatenate([L, L + 1])
        frontier = kids[~is_leaf[kids]]
    internal = ~is_leaf
    feat = np.where(internal, (meta >> 24) & 0xff, 0)
    if internal.any() and feat[internal].max() >= n_features:
        raise ValueError("tree uses a feature index outside the feature map")
    # per-feature threshold tables and the rank of every node's threshold
    thr = nodes[:, 0].view(np.float32)
    tabs, thr_off, rank = [], [0], np.zeros(n, dtype=np.int64)
    for f in range(n_features):
        sel = internal & (feat == f)
        s = np.unique(thr[sel])
        if s.size > 65535:
            return None
        rank[sel] = np.searchsorted(s, thr[sel])
        tabs.append(s)
        thr_off.append(thr_off[-1] + s.size)
    # subtree sizes in words (bottom-up), then positions (top-down)
    size = np.where(is_leaf, 2, 0).astype(np.int64)
    for lv in reversed(levels):
        size[lv] = 1 + size[left[lv]] + size[left[lv] + 1]
    tree_words = (size[roots] + RANKED_TREE_ALIGN - 1) // RANKED_TREE_ALIGN * RANKED_TREE_ALIGN
    tree_start = np.concatenate([[0], np.cumsum(tree_words)])
    if tree_words.max() > RANKED_MAX_TREE_WORDS or tree_start[-1] >= 2 ** 31:
        return None
    pos = np.zeros(n, dtype=np.int64)
    pos[roots] = tree_start[:-1]
    for lv in levels:
        pos[left[lv]] = pos[lv] + 1
        pos[left[lv] + 1] = pos[lv] + 1 + size[left[lv]]
    off = np.where(internal, 1 + size[np.where(internal, left, 0)], 0)
    bits = lambda v: max(1, int(v).bit_length())
    fbits = bits(n_features - 1)
    rbits = bits(max(s.size for s in tabs))       # a voxel's rank goes up to |S_f| inclusive
    if fbits + rbits + bits(off.max()) > 30:
        return None
    words = np.zeros(int(tree_start[-1]), dtype=np.uint32)
    ii = np.flatnonzero(internal)
    words[pos[ii]] = (lleaf[ii] | (rleaf[ii] << 1) | (off[ii] << 2) |
                      (rank[ii] << (32 - fbits - rbits)) | (feat[ii] << (32 - fbits))).astype(np.uint32)
    li = np.flatnonzero(is_leaf)
    words[pos[li]] = nodes[li, 0]
    words[pos[li] + 1] = nodes[li, 1]
    return {"words": words, "tree_start": tree_start.astype(np.int32),
            "thr_tab": (np.concatenate(tabs) if tabs else np.zeros(0)).astype(np.float32),
            "thr_off": np.asarray(thr_off, dtype=np.int32), "fbits": fbits, "rbits": rbits,
            "max_tree_words": int(tree_words.max())}


class RankedForest(ctypes.Structure):
    """``lsml_ranked_forest`` of include/lsml_b200.h."""
    _fields_ = [("words", ctypes.c_void_p), ("tree_start", ctypes.c_void_p),
                ("root_is_leaf", ctypes.c_void_p), ("thr_tab", ctypes.c_void_p),
                ("thr_off", ctypes.c_void_p), ("n_trees", ctypes.c_int), ("nfeat", ctypes.c_int),
                ("fbits", ctypes.c_int), ("rbits", ctypes.c_int),
                ("max_tree_words", ctypes.c_int), ("divisor", ctypes.c_double)]


class ForestDeviceRegressor(DeviceRegressor):
    """``nodes`` / ``tree_base`` / ``root_is_leaf`` are the canonical single-node export
    (:func:`pack_trees`, also what the golden fixtures store); the kernels walk layouts derived
    from it on first use: rank-quantised 4-byte nodes from shared memory (:func:`pack_ranked`)
    when the forest fits that format, else 32-byte super nodes (:func:`pack_supernodes`)."""

    def __init__(self, trees, n_features, divisor, scaler, device):
        super().__init__(n_features, scaler, device)
        nodes, bases, root_leaf = pack_trees(trees, n_features)
        self.n_trees = len(bases) - 1
        self.n_nodes = int(nodes.shape[0])
        self.nodes = _dev(nodes, device)
        self.tree_base = _dev(bases, device)
        self.root_is_leaf = _dev(root_leaf, device)
        self.divisor = float(divisor)
        self.max_nodes = int(np.diff(bases).max()) if len(bases) > 1 else 0
        self._super = None                  # derived layouts are built on first use
        self._ranked = None

    def _pack_super(self, nodes, bases, root_leaf):
        sn, leaves, root_ref, _ = pack_supernodes(nodes, bases, root_leaf)
        if sn.shape[0] == 0:
            sn = np.zeros((1, 8), dtype=np.uint32)
        # torch's allocator hands out 512-byte aligned blocks: the 32-byte alignment holds
        return (_dev(sn, self.device), _dev(leaves, self.device), _dev(root_ref, self.device))

    def _layout(self):
        if getattr(self, "_super", None) is None:      # built around exported arrays (fixtures)
            bases = self.tree_base.cpu().numpy()
            self.max_nodes = int(np.diff(bases).max()) if len(bases) > 1 else 0
            self._super = self._pack_super(self.nodes.cpu().numpy(), bases,
                                           self.root_is_leaf.cpu().numpy())
        return self._super

    def _ranked_layout(self):
        """The rank-quantised layout (``RankedForest`` struct + the tensors it points into), or
        None when the forest does not fit it / LSML_FOREST_VARIANT asks for another kernel."""
        if getattr(self, "_ranked", None) is None:
            self._ranked = False
            variant = os.environ.get("LSML_FOREST_VARIANT", "")
            pk = None
            if variant in ("", "ranked") and self.n_features <= 64:
                pk = pack_ranked(self.nodes.cpu().numpy(), self.tree_base.cpu().numpy(),
                                 self.root_is_leaf.cpu().numpy(), self.n_features)
            if pk is not None:
                keep = [_dev(pk[k], self.device) for k in ("words", "tree_start", "thr_tab",
                                                           "thr_off")]
                if keep[2].numel() == 0:               # (no internal node anywhere)
                    keep[2] = torch.zeros(1, dtype=torch.float32, device=self.device)
                rf = RankedForest(keep[0].data_ptr(), keep[1].data_ptr(),
                                  self.root_is_leaf.data_ptr(), keep[2].data_ptr(),
                                  keep[3].data_ptr(), self.n_trees, self.n_features, pk["fbits"],
                                  pk["rbits"], pk["max_tree_words"], self.divisor)
                self._ranked = (rf, keep)
        return self._ranked or None

    def predict_into(self, X, n_band, out):
        rk = self._ranked_layout()
        if rk is not None:
            _lib.check(_lib.lib.lsml_forest_ranked_predict(
                _lib.ptr(X), _lib.ptr(n_band), ctypes.c_int(self.n_features), _lib.ptr(self.mean),
                _lib.ptr(self.scale), ctypes.byref(rk[0]), _lib.ptr(out), _lib.current_stream()),
                "forest_ranked_predict")
            return
        sn, leaves, root_ref = self._layout()
        _lib.check(_lib.lib.lsml_forest_predict(
            _lib.ptr(X), _lib.ptr(n_band), ctypes.c_int(self.n_features), _lib.ptr(self.mean),
            _lib.ptr(self.scale), _lib.ptr(sn), _lib.ptr(leaves), _lib.ptr(root_ref),
            _lib.ptr(self.root_is_leaf), ctypes.c_int(self.n_trees),
            ctypes.c_double(self.divisor), _lib.ptr(self.nodes), _lib.ptr(self.tree_base),
            ctypes.c_int(self.max_nodes), _lib.ptr(out), _lib.current_stream()),
            "forest_predict")

    def predict_band(self, src, n_band, out):
        rk = self._ranked_layout()
        if rk is not None:
            _lib.check(_lib.lib.lsml_forest_ranked_predict_band(
                ctypes.byref(src), _lib.ptr(n_band), _lib.ptr(self.mean), _lib.ptr(self.scale),
                ctypes.byref(rk[0]), _lib.ptr(out), _lib.current_stream()),
                "forest_ranked_predict_band")
            return
        sn, leaves, root_ref = self._layout()
        _lib.check(_lib.lib.lsml_forest_predict_band(
            ctypes.byref(src), _lib.ptr(n_band), _lib.ptr(self.mean),
            _lib.ptr(self.scale), _lib.ptr(sn), _lib.ptr(leaves), _lib.ptr(root_ref),
            _lib.ptr(self.root_is_leaf), ctypes.c_int(self.n_trees),
            ctypes.c_double(self.divisor), _lib.ptr(self.nodes), _lib.ptr(self.tree_base),
            ctypes.c_int(self.max_nodes), _lib.ptr(out), _lib.current_stream()),
            "forest_predict_band")


_ACT = {"identity": 0, "relu": 1, "tanh": 2, "logistic": 3}


class MlpDeviceRegressor(DeviceRegressor):
    def __init__(self, coefs, intercepts, activation, scaler, device):
        super().__init__(coefs[0].shape[0], scaler, device)
        if coefs[-1].shape[1] != 1:
            raise ValueError("only single-output MLP regressors are supported")
        self.W = [_dev(w, device, np.float64) for w in coefs]
        self.b = [_dev(b, device, np.float64) for b in intercepts]
        self.sizes = [coefs[0].shape[0]] + [w.shape[1] for w in coefs]
        self.act = _ACT[activation]

    def _layers(self):
        n = len(self.W)
        Wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in self.W])
        bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in self.b])
        return n, Wp, bp

    def predict_into(self, X, n_band, out):
        n, Wp, bp = self._layers()
        _lib.check(_lib.lib.lsml_mlp_predict(
            _lib.ptr(X), _lib.ptr(n_band), _lib.ptr(self.mean), _lib.ptr(self.scale),
            ctypes.c_int(n), _lib.int_array(self.sizes), Wp, bp, ctypes.c_int(self.act),
            _lib.ptr(out), _lib.current_stream()), "mlp_predict")

    def predict_band(self, src, n_band, out):
        n, Wp, bp = self._layers()
        _lib.check(_lib.lib.lsml_mlp_predict_band(
            ctypes.byref(src), _lib.ptr(n_band), _lib.ptr(self.mean), _lib.ptr(self.scale),
            ctypes.c_int(n), _lib.int_array(self.sizes), Wp, bp, ctypes.c_int(self.act),
            _lib.ptr(out), _lib.current_stream()), "mlp_predict_band")


def export_regressor(model, device="cuda"):
    """Turn a fitted scikit-learn regressor (optionally inside a Pipeline with a StandardScaler)
    into a :class:`DeviceRegressor`."""
    if isinstance(model, DeviceRegressor):
        return model
    device = torch.device(device)
    scaler = None
    final = model
    cls = type(model).__name__
    if cls == "Pipeline":
        steps = [s for _, s in model.steps if s is not None and s != "passthrough"]
        final = steps[-1]
        pre = steps[:-1]
        if len(pre) > 1 or (pre and type(pre[0]).__name__ != "StandardScaler"):
            raise TypeError("only Pipeline([StandardScaler,] regressor) can be exported")
        if pre:
            sc = pre[0]
            nfeat = sc.n_features_in_
            mean = sc.mean_ if sc.with_mean and sc.mean_ is not None else np.zeros(nfeat)
            scale = sc.scale_ if sc.with_std and sc.scale_ is not None else np.ones(nfeat)
            scaler = (mean, scale)
    name = type(final).__name__
    if name in ("LinearRegression", "Ridge", "Lasso", "ElasticNet", "SGDRegressor",
                "HuberRegressor", "BayesianRidge"):
        return LinearDeviceRegressor(final.coef_, final.intercept_, scaler, device)
    if name in ("RandomForestRegressor", "ExtraTreesRegressor"):
        trees = [e.tree_ for e in final.estimators_]
        return ForestDeviceRegressor(trees, final.n_features_in_, len(trees), scaler, device)
    if name in ("DecisionTreeRegressor", "ExtraTreeRegressor"):
        return ForestDeviceRegressor([final.tree_], final.n_features_in_, 1.0, scaler, device)
    if name == "MLPRegressor":
        return MlpDeviceRegressor(final.coefs_, final.intercepts_, final.activation, scaler, device)
    raise TypeError("no device exporter for regressor class %r (supported: linear models, "
                    "RandomForest/ExtraTrees/DecisionTree regressors, MLPRegressor, optionally "
                    "behind a StandardScaler in a Pipeline); there is no CPU fallback" % name)
