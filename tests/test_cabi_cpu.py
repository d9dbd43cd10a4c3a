"""CPU-only: the C-ABI library loads and exports every symbol include/lsml_b200.h declares
(no compute call is made -- there is no GPU here), and the product fails loudly without one."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lsml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    skip = {"defined", "C"}
    return sorted({n for n in names if n not in skip and (n.startswith("lsml_") or
                                                        n.startswith("gmag_os") or
                                                        n.startswith("gradient_centered"))})


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "lsml_b200", "liblsml_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 30
    for dropin in ("gmag_os1d", "gmag_os2d", "gmag_os3d", "gradient_centered1d",
                   "gradient_centered2d", "gradient_centered3d"):
        assert dropin in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.lsml_version() >= 100


def test_sass_carries_the_blackwell_instructions_the_design_claims():
    """DESIGN.md names three hardware paths; the built library must contain them: TMA tensor loads
    (dense stencils), a bulk global->shared copy completing on an mbarrier (tree groups of the
    ranked forest kernel) and FP64 tensor-core MMAs (MLP layers).  Needs cuobjdump (CUDA toolkit)."""
    import shutil, subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from lsml_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper() or "EF_CUDA_SM100" in sass
    for mnemonic, where in (("UTMALDG", "gmag_os_tma_kernel"), ("UBLKCP", "forest_ranked_kernel"),
                            ("SYNCS", "mbarrier waits"), ("DMMA", "mlp_dmma_kernel")):
        assert mnemonic in sass, "%s (%s) missing from the SASS of %s" % (mnemonic, where, _lib.LIB_PATH)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lsml_b200 import _lib
    from lsml_b200.gradient import masked_gradient as mg
    with pytest.raises(_lib.LsmlError):
        mg.gradient_magnitude_osher_sethian(np.zeros((4, 4)), np.zeros((4, 4)))
    from lsml_b200.core.engine import BandEngine
    with pytest.raises(_lib.LsmlError):
        BandEngine((8, 8), 1)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lsml_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "import oracle" not in text and "from oracle" not in text, name
                assert "liblsml_oracle" not in text, name


def test_reference_validation_errors_of_masked_gradient():
    from lsml_b200.gradient import masked_gradient as mg
    with pytest.raises(AssertionError):
        mg.gradient_centered(np.zeros((2, 2, 2, 2)))
    with pytest.raises(ValueError):
        mg.gradient_centered(np.zeros((4, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        mg.gradient_magnitude_osher_sethian(np.zeros((4, 4)), np.zeros((4, 4)), dx=[1.0])
    with pytest.raises(ValueError):
        mg.gradient_magnitude_osher_sethian(np.zeros((4, 4)), np.zeros((4, 4)),
                                            mask=np.ones(4, dtype=bool))


def test_gaussian_weights_identical_to_scipy():
    from scipy.ndimage._filters import _gaussian_kernel1d
    from lsml_b200.core.engine import gaussian_kernel1d, gaussian_weights
    for sigma in (0.5, 1.0, 2.0, 2.0 / 0.7, 3.0, 7.5):
        for order in (0, 1, 2):
            lw = int(4.0 * sigma + 0.5)
            assert np.array_equal(gaussian_kernel1d(sigma, order, lw),
                                  _gaussian_kernel1d(sigma, order, lw))
        w, r = gaussian_weights(sigma, 1)
        assert len(w) == 2 * r + 1


def test_feature_program_and_api_surface():
    from lsml_b200.core.engine import FeatureProgram
    from lsml_b200.feature import (BoundarySize, COMRaySample, FeatureMap, Moments,
                                   get_basic_image_features, get_basic_shape_features)
    feats = get_basic_image_features() + get_basic_shape_features()
    fm = FeatureMap(feats)
    assert fm.n_features == 16 and fm.feature_slices[11] == slice(11, 15)
    assert [f.name for f in feats][:2] == ["Image sample (σ = 0.000)", "Image sample (σ = 2.000)"]
    prog = FeatureProgram(fm.specs(), 2)
    assert prog.nfeat == 16 and len(prog.planes) == 4 and prog.needs_boundary
    p3 = FeatureProgram([BoundarySize(3).spec()], 3)         # 3-D: marching-cubes area (mcubes.cu)
    assert p3.needs_boundary and p3.nfeat == 1
    hi = FeatureProgram([Moments(2, orders=(1, 3, 4)).spec()], 2)   # orders > 2: axis histograms
    assert hi.nfeat == 6 and hi.hi_moments == [(1, 0, 3), (2, 0, 4), (4, 1, 3), (5, 1, 4)]
    with pytest.raises(ValueError):
        FeatureProgram([("moments", (0,), (0,))], 2)
    ray = FeatureProgram([COMRaySample(2, n_samples=3).spec()], 2)      # 2N columns, raw image
    assert ray.nfeat == 6 and ray.planes == [("sample", 0.0)]
    assert [b & 0xffff for b in ray.b] == list(range(6)) and {b >> 16 for b in ray.b} == {3}
    with pytest.raises(ValueError):
        FeatureProgram([COMRaySample(2).spec()], 1)
    with pytest.raises(ValueError):
        Moments(2, axes=(2,))
    with pytest.raises(ValueError):
        BoundarySize(1)


def test_pack_trees_round_trip():
    """The 8-byte node packing reproduces sklearn's predictions (numpy traversal)."""
    from sklearn.ensemble import RandomForestRegressor
    from lsml_b200.core.regressors import pack_trees
    rs = np.random.RandomState(0)
    X = rs.randn(500, 6)
    y = X[:, 0] * 2 + np.sin(X[:, 1]) + 0.1 * rs.randn(500)
    rf = RandomForestRegressor(n_estimators=8, max_depth=7, random_state=0).fit(X, y)
    nodes, bases, root = pack_trees([e.tree_ for e in rf.estimators_], 6)
    Xt = rs.randn(200, 6)
    Xt[:50] = X[:50]
    X32 = Xt.astype(np.float32)
    acc = np.zeros(200)
    assert len(bases) == len(root) + 1 and bases[-1] == nodes.shape[0]
    for t in range(len(root)):
        tree = nodes[bases[t]:bases[t + 1]]
        for i in range(200):
            idx, leaf = 0, bool(root[t])
            while not leaf:
                thr = tree[idx, 0:1].view(np.float32)[0]
                meta = int(tree[idx, 1])
                right = not (X32[i, meta >> 24] <= thr)
                leaf = bool((meta >> 22) & 1) if right else bool((meta >> 23) & 1)
                idx = (meta & 0x3fffff) + int(right)
            acc[i] += tree[idx].view(np.float64)[0]
    assert np.array_equal(acc / len(root), rf.predict(Xt))


def test_pack_supernodes_round_trip():
    """The 32-byte super-node layout the kernels walk (two tree levels per fetch) reaches the same
    leaves as scikit-learn, including depth-1 / depth-2 / single-leaf trees."""
    from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
    from sklearn.tree import DecisionTreeRegressor
    from lsml_b200.core.regressors import pack_supernodes, pack_trees
    rs = np.random.RandomState(0)
    X = rs.randn(600, 6)
    y = X[:, 0] * 2 + np.sin(X[:, 1]) + 0.1 * rs.randn(600)

    def walk(sn, leaves, root_ref, root_leaf, X32):
        out = np.zeros(len(X32))
        for t in range(len(root_ref)):
            for i in range(len(X32)):
                ref, leaf = int(root_ref[t]), bool(root_leaf[t])
                while not leaf:
                    nd = sn[ref]
                    thr, ff = nd[:3].view(np.float32), int(nd[3])
                    flags = ff >> 24
                    r0 = not (X32[i, ff & 0xff] <= thr[0])
                    if (flags >> int(r0)) & 1:
                        ref = int(nd[4 + 2 * int(r0)])
                        break
                    f1 = (ff >> 16) & 0xff if r0 else (ff >> 8) & 0xff
                    r1 = not (X32[i, f1] <= (thr[2] if r0 else thr[1]))
                    g = 2 * int(r0) + int(r1)
                    ref, leaf = int(nd[4 + g]), bool((flags >> (2 + g)) & 1)
                out[i] += leaves[ref]
        return out

    models = [RandomForestRegressor(n_estimators=6, max_depth=7, random_state=0).fit(X, y),
              RandomForestRegressor(n_estimators=5, max_samples=50, random_state=1).fit(X, y),
              ExtraTreesRegressor(n_estimators=3, max_depth=9, random_state=2).fit(X, y),
              DecisionTreeRegressor(max_depth=1).fit(X, y),
              DecisionTreeRegressor(max_depth=2).fit(X, y),
              DecisionTreeRegressor().fit(X, np.ones(600))]          # a single leaf
    Xt = rs.randn(120, 6)
    Xt[:40] = X[:40]
    for m in models:
        trees = [e.tree_ for e in m.estimators_] if hasattr(m, "estimators_") else [m.tree_]
        nodes, bases, root = pack_trees(trees, 6)
        sn, leaves, root_ref, root_leaf = pack_supernodes(nodes, bases, root)
        assert sn.dtype == np.uint32 and sn.shape[1] == 8 and leaves.dtype == np.float64
        got = walk(sn, leaves, root_ref, root_leaf, Xt.astype(np.float32)) / len(trees)
        assert np.array_equal(got, m.predict(Xt)), type(m).__name__


def _walk_ranked(pk, root_leaf, X32):
    """numpy restatement of regress.cu: forest_ranked_kernel (ranks once, 4-byte nodes)."""
    words, start = pk["words"], pk["tree_start"]
    tab, off, fb, rb = pk["thr_tab"], pk["thr_off"], pk["fbits"], pk["rbits"]
    F = len(off) - 1
    ranks = np.empty((len(X32), F), dtype=np.int64)
    for f in range(F):
        s = tab[off[f]:off[f + 1]]
        ranks[:, f] = np.where(np.isnan(X32[:, f]), len(s), np.searchsorted(s, X32[:, f], "left"))
    out = np.zeros(len(X32))
    for t in range(len(root_leaf)):
        for i in range(len(X32)):
            idx, leaf = int(start[t]), bool(root_leaf[t])
            while not leaf:
                n = int(words[idx])
                f = n >> (32 - fb)
                # the kernel's comparison: the voxel's (feature | rank) word against the node word
                right = ((f << (32 - fb)) | (int(ranks[i, f]) << (32 - fb - rb))) > n
                assert right == (ranks[i, f] > ((n >> (32 - fb - rb)) & ((1 << rb) - 1)))
                leaf = bool(n & (2 if right else 1))
                idx += ((n & ((1 << (32 - fb - rb)) - 4)) >> 2) if right else 1
            out[i] += words[idx:idx + 2].view(np.float64)[0]
    return out


def test_pack_ranked_round_trip():
    """The rank-quantised 4-byte-node layout (shared-memory forest kernel) reaches the same leaves
    as scikit-learn: thresholds replaced by their rank among the forest's thresholds of that
    feature, features by their rank -- including values EQUAL to a threshold, NaN, depth-1 /
    single-leaf trees and the bench's forest recipe (max_samples=300)."""
    from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
    from sklearn.tree import DecisionTreeRegressor
    from lsml_b200.core.regressors import pack_ranked, pack_trees
    rs = np.random.RandomState(3)
    X = rs.randn(600, 6)
    X[:, 5] = np.round(X[:, 5])                        # few distinct values: shared thresholds
    y = X[:, 0] * 2 + np.sin(X[:, 1]) + X[:, 5] + 0.1 * rs.randn(600)
    models = [RandomForestRegressor(n_estimators=6, max_depth=7, random_state=0).fit(X, y),
              RandomForestRegressor(n_estimators=9, max_samples=300, random_state=1).fit(X, y),
              ExtraTreesRegressor(n_estimators=3, max_depth=9, random_state=2).fit(X, y),
              DecisionTreeRegressor(max_depth=1).fit(X, y),
              DecisionTreeRegressor(max_depth=2).fit(X, y),
              DecisionTreeRegressor().fit(X, np.ones(600))]          # a single leaf
    Xt = rs.randn(150, 6)
    Xt[:40] = X[:40]
    Xt[:, 5] = np.round(Xt[:, 5])
    for m in models:
        trees = [e.tree_ for e in m.estimators_] if hasattr(m, "estimators_") else [m.tree_]
        nodes, bases, root = pack_trees(trees, 6)
        pk = pack_ranked(nodes, bases, root, 6)
        assert pk is not None and pk["words"].dtype == np.uint32
        assert (pk["tree_start"] % 4 == 0).all() and pk["tree_start"][-1] == len(pk["words"])
        assert pk["fbits"] + pk["rbits"] <= 28 and int(np.diff(pk["thr_off"]).max(initial=0)) < (1 << pk["rbits"])
        X32 = Xt.astype(np.float32)
        # values exactly ON thresholds (x <= thr goes left) and just above them
        tab = pk["thr_tab"]
        if len(tab):
            f0 = int(np.argmax(np.diff(pk["thr_off"])))
            s = tab[pk["thr_off"][f0]:pk["thr_off"][f0 + 1]]
            k = min(len(s), 20)
            X32[40:40 + k, f0] = s[:k]
            X32[60:60 + k, f0] = np.nextafter(s[:k], np.float32(np.inf))
        got = _walk_ranked(pk, root, X32) / len(trees)
        assert np.array_equal(got, m.predict(X32.astype(np.float64))), type(m).__name__
    # NaN features go right everywhere, as `!(x <= thr)` does in the other kernels
    m = models[1]
    nodes, bases, root = pack_trees([e.tree_ for e in m.estimators_], 6)
    pk = pack_ranked(nodes, bases, root, 6)
    Xn = np.full((1, 6), np.nan, dtype=np.float32)
    want = 0.0
    for e in m.estimators_:
        t, i = e.tree_, 0
        while t.children_left[i] != -1:
            i = t.children_right[i]
        want += t.value[i, 0, 0]
    assert _walk_ranked(pk, root, Xn)[0] == want
    # a forest that does not fit the format is refused, not mangled
    big = DecisionTreeRegressor().fit(rs.randn(9000, 6), rs.randn(9000))
    nodes, bases, root = pack_trees([big.tree_], 6)
    assert pack_ranked(nodes, bases, root, 6) is None


def test_fit_host_logic_matches_reference_statements():
    """The host-side pieces of `fit` that decide WHICH data is used, against literal
    transcriptions of the reference's statements (datasets_handler.py:312-362; the reference's
    own function needs h5py): same split for the same RandomState, same RNG consumption."""
    from lsml_b200.core.model import split_datasets

    def reference_split(n, probs, subset_size, rs):
        keys = ["%05d" % i for i in range(n)]
        if subset_size is None:
            subset_size = len(keys)
        sub_keys = rs.choice(keys, replace=False, size=subset_size)
        indicators = rs.multinomial(n=1, pvals=probs, size=len(sub_keys))
        return [[i for i, ind in enumerate(indicators) if list(ind).index(1) == d] for d in range(3)]

    for seed in range(40):
        n = 5 + seed
        subset = None if seed % 3 else max(1, n // 2)
        want = reference_split(n, (0.6, 0.2, 0.2), subset, np.random.RandomState(seed))
        rs = np.random.RandomState(seed)
        got = split_datasets(n, (0.6, 0.2, 0.2), subset, rs)
        assert [list(g) for g in got] == want
        ref_rs = np.random.RandomState(seed)
        reference_split(n, (0.6, 0.2, 0.2), subset, ref_rs)
        assert rs.randint(1 << 30) == ref_rs.randint(1 << 30)
    tr, va, te = split_datasets(10, ([0, 1, 2], [3], [4, 5]), None, np.random.RandomState(0))
    assert list(tr) == [0, 1, 2] and list(va) == [3] and list(te) == [4, 5]
    with pytest.raises(ValueError):
        split_datasets(4, (0.6, 0.2, 0.2), 9, np.random.RandomState(0))
    with pytest.raises(ValueError):
        split_datasets(4, (1, 0, 0), None, np.random.RandomState(0))
