"""How far is the product's band from upstream scikit-fmm's?  

The band rebuild the product reproduces bit for bit is oracle/lsml_oracle.c:orc_fmm_distance, a
marcher that departs from scikit-fmm on purpose (float32 ordering keys + flat-index ties, causality
test with contributor eviction, no break at the band edge, value2 reset).  scikit-fmm itself is
absent here, so the gap cannot be measured against the real thing; this script measures it against
oracle/skfmm_literal.c, the upstream-literal restatement (float64 keys, upstream's heap
mechanics, no causality test, break at `narrow`, value2 kept).

For each BASELINE-like recipe (C1 2-D 51^2 linear, C2 2-D 101^2 forest, C3 3-D volume forest,
C4 64^3 crops forest) level sets are evolved on the CPU with the oracle (product marcher).  Every
rebuild input u is ALSO given to the literal marcher and the two outputs are compared:

    mask_mismatch_voxels      band voxels in one mask and not the other
    rebuilds_with_mismatch    rebuild inputs with at least one such voxel
    dist_max_abs_diff         over voxels in both bands
    dist_bitwise_equal_frac   fraction of common band voxels with identical float64 distances

A second track evolves the same items under the LITERAL marcher from the start and compares the
final segmentations {u > 0} of the two tracks (what a user of the model sees).

    python tests/tools/band_gap_report.py [--quick] > profiles/r2_band_gap_vs_upstream_literal.json

TEST INFRASTRUCTURE (uses oracle/; lives under tests/ for that reason).
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def fit_reg(kind, X, y, seed):
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.linear_model import LinearRegression
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    rs = np.random.RandomState(seed)
    if X.shape[0] > 20000:
        sub = rs.choice(X.shape[0], size=20000, replace=False)
        X, y = X[sub], y[sub]
    est = (LinearRegression() if kind == "linear" else
           RandomForestRegressor(n_estimators=100, max_samples=min(300, X.shape[0]), random_state=rs))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return Pipeline([("standardscaler", StandardScaler()), ("regressor", est)]).fit(X, y)


def run_recipe(O, name, items, shape, specs, reg_kind, iters, step, r_init, band=3.0, log=None,
               collect=None):
    nd = len(shape)
    dx = np.ones(nd)
    g = np.indices(shape).astype(np.float64)
    rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(nd)))
    u0 = np.where(rc < r_init, 1.0, -1.0)
    st = dict(recipe=name, items=len(items), shape=list(shape), iterations=iters, rebuilds=0,
              band_voxels=0, mask_mismatch_voxels=0, rebuilds_with_mismatch=0,
              dist_max_abs_diff=0.0, dist_common=0, dist_bitwise_equal=0,
              final_seg_voxels=0, final_seg_mismatch_voxels=0, items_with_final_seg_mismatch=0,
              final_band_mismatch_voxels=0)
    reg = None
    collect = collect if collect is not None else []
    for b, (img, truth) in enumerate(items):
        img = O.normalize_image(img)
        tracks = {}
        for literal in (False, True):
            u = u0.copy()
            dist, mask = O.distance_transform(u, band, dx, literal=literal)
            for it in range(iters):
                if not mask.any():
                    break
                if reg is None:       # one regressor for the recipe: target = true signed distance
                    X = O.feature_matrix(specs, u, img, dist, mask, dx)
                    reg = fit_reg(reg_kind, X, truth[mask], 5)
                u, _, _, _ = O.evolve_step(u, img, dist, mask, dx, specs, reg, step, band,
                                           rebuild=False)
                dist, mask = O.distance_transform(u, band, dx, literal=literal)
                if not literal:       # the same rebuild input through the literal marcher
                    if len(collect) < 120 and it % 3 == 0:
                        collect.append(u.copy())
                    ld, lm = O.distance_transform(u, band, dx, literal=True)
                    st["rebuilds"] += 1
                    st["band_voxels"] += int(mask.sum())
                    mm = int((mask != lm).sum())
                    st["mask_mismatch_voxels"] += mm
                    st["rebuilds_with_mismatch"] += 1 if mm else 0
                    both = mask & lm
                    if both.any():
                        st["dist_max_abs_diff"] = max(st["dist_max_abs_diff"],
                                                      float(np.abs(dist - ld)[both].max()))
                        st["dist_common"] += int(both.sum())
                        st["dist_bitwise_equal"] += int((dist[both] == ld[both]).sum())
            tracks[literal] = (u, mask)
        seg_p, seg_l = tracks[False][0] > 0, tracks[True][0] > 0
        d = int((seg_p != seg_l).sum())
        st["final_seg_voxels"] += int(seg_p.sum())
        st["final_seg_mismatch_voxels"] += d
        st["items_with_final_seg_mismatch"] += 1 if d else 0
        st["final_band_mismatch_voxels"] += int((tracks[False][1] != tracks[True][1]).sum())
        if log:
            log("%s item %d/%d: rebuilds %d, mask mismatch voxels so far %d, final seg diff %d"
                % (name, b + 1, len(items), st["rebuilds"], st["mask_mismatch_voxels"], d))
    st["dist_bitwise_equal_frac"] = st["dist_bitwise_equal"] / max(1, st["dist_common"])
    st["mask_mismatch_per_million_band_voxels"] = 1e6 * st["mask_mismatch_voxels"] / max(1, st["band_voxels"])
    return st


def ablation(O, inputs, band=3.0):
    """Attribute the gap: the oracle marcher with ONE departure from upstream switched back at a
    time (oracle/lsml_oracle.c: orc_fmm_set_ablation) against the literal marcher, on the rebuild
    inputs collected from the product track."""
    names = {0: "product marcher (all four departures)", 1: "float64 keys (flat-index ties kept)",
             2: "no causality test / eviction", 4: "value2 kept (upstream quirk)",
             8: "break at the band edge", 15: "all four switched back (only heap tie mechanics "
                                              "left)"}
    rows = []
    for flags, label in names.items():
        O._lib().orc_fmm_set_ablation(flags)
        mm = nb = reb = 0
        worst = 0.0
        for u in inputs:
            dx = np.ones(u.ndim)
            d, m = O.distance_transform(u, band, dx)
            ld, lm = O.distance_transform(u, band, dx, literal=True)
            k = int((m != lm).sum())
            mm += k; nb += int(m.sum()); reb += 1 if k else 0
            both = m & lm
            if both.any():
                worst = max(worst, float(np.abs(d - ld)[both].max()))
        rows.append(dict(switched_back=label, flags=flags, mask_mismatch_voxels=mm, band_voxels=nb,
                         rebuilds_with_mismatch=reb, rebuilds=len(inputs), dist_max_abs_diff=worst))
    O._lib().orc_fmm_set_ablation(0)
    return rows


def blob_items(count, shape, seed, r_lo, r_hi, jitter, noise=0.2):
    rs = np.random.RandomState(seed)
    nd = len(shape)
    g = np.indices(shape).astype(np.float64)
    out = []
    for _ in range(count):
        c = np.array(shape) / 2.0 + rs.uniform(-jitter, jitter, size=nd)
        rad = rs.uniform(r_lo, r_hi)
        r = np.sqrt(sum((g[a] - c[a]) ** 2 for a in range(nd)))
        out.append(((r < rad).astype(np.float64) + noise * rs.standard_normal(shape), rad - r))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="a fraction of the items (CI-sized)")
    ap.add_argument("--no-ablation", action="store_true")
    args = ap.parse_args()
    import bench
    from oracle import oracle as O
    O.build()
    t0 = time.time()
    log = lambda m: print("[%6.0f s] %s" % (time.time() - t0, m), file=sys.stderr, flush=True)
    q = 4 if args.quick else 1
    img2 = [("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
            ("band_mean", 0.0), ("band_mean", 2.0), ("band_std", 0.0), ("band_std", 2.0)]
    shape2 = [("size",), ("boundary_size",), ("isoperimetric",), ("moments", (0, 1), (1, 2)),
              ("dist_to_com",)]
    out = {"what": "product marcher (oracle/lsml_oracle.c:orc_fmm_distance, which fmm.cu matches "
                   "bit for bit) vs upstream-literal restatement of scikit-fmm "
                   "(oracle/skfmm_literal.c) on the same rebuild inputs; band = 3, dx = 1",
           "recipes": []}
    inputs2, inputs3 = [], []
    # C1: hamburger2d-like 51^2 (a disc with a bar removed), linear velocity, F = 16
    rs = np.random.RandomState(1234)
    items = []
    g = np.indices((51, 51)).astype(np.float64)
    for _ in range(24 // q):
        c = 25.0 + rs.uniform(-3, 3, size=2)
        rad = rs.uniform(12, 20)
        r = np.sqrt((g[0] - c[0]) ** 2 + (g[1] - c[1]) ** 2)
        th = rs.uniform(0, np.pi)
        bar = np.abs(np.cos(th) * (g[0] - c[0]) + np.sin(th) * (g[1] - c[1])) < 2.0
        img = ((r < rad) & ~bar).astype(np.float64) + 0.25 * rs.standard_normal((51, 51))
        items.append((img, rad - r))
    out["recipes"].append(run_recipe(O, "C1 (51x51, linear, 40 iterations)", items, (51, 51),
                                     img2 + shape2, "linear", 40, 0.3, 8.0, log=log,
                                     collect=inputs2))
    # C2: gestalt2d-size 101^2, forest, 10 iterations (bench.py --workload c2 recipe)
    out["recipes"].append(run_recipe(O, "C2 (101x101, RF-100, 10 iterations)",
                                     blob_items(48 // q, (101, 101), 20000, 15, 35, 10), (101, 101),
                                     img2 + shape2, "forest", 10, 0.4, 10.0, log=log))
    # C4: 64^3 crops, forest, 30 iterations (bench.py --workload c4 recipe)
    out["recipes"].append(run_recipe(O, "C4 (64^3 crops, RF-100, 30 iterations)",
                                     blob_items(12 // q, (64, 64, 64), 10000, 6, 18, 8), (64, 64, 64),
                                     bench.SPECS, "forest", 30, 0.4, 5.0, log=log, collect=inputs3))
    # C3: one hamburger volume at 96^3 (bench.make_volume scaled), forest, 25 iterations
    n = 96
    img, radius, c = bench.make_volume(n, 1234)
    g3 = np.indices((n, n, n)).astype(np.float64)
    truth = radius - np.sqrt(sum((g3[a] - c[a]) ** 2 for a in range(3)))
    out["recipes"].append(run_recipe(O, "C3 (96^3 hamburger volume, RF-100, 25 iterations)",
                                     [(img, truth)], (n, n, n), bench.SPECS, "forest", 25,
                                     bench.STEP, bench.INIT_RADIUS * n / 256.0, log=log))
    if not args.no_ablation:
        out["ablation_C1_inputs"] = ablation(O, inputs2)
        out["ablation_C4_inputs"] = ablation(O, inputs3[:30 if args.quick else 60])
        log("ablation done")
    tot = {k: sum(r[k] for r in out["recipes"]) for k in
           ("rebuilds", "band_voxels", "mask_mismatch_voxels", "rebuilds_with_mismatch",
            "final_seg_voxels", "final_seg_mismatch_voxels")}
    tot["dist_max_abs_diff"] = max(r["dist_max_abs_diff"] for r in out["recipes"])
    out["total"] = tot
    out["seconds"] = time.time() - t0
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
