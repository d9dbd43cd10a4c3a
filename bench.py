#!/usr/bin/env python
"""bench.py -- narrow-band voxel-iterations / s of the lsml level-set evolution hot path.

Workload (BASELINE.json configs[2], the configuration the north_star target is quoted on;
SURVEY.md section 8d "C3"): one synthetic 256^3 volume per GPU (a noisy "hamburger" sphere with a
slab cut), +-1 ball initialisation r=40, band=3, dx=1, F=16 features (8 image features at
sigma in {0,2}, distance to centre of mass, 6 moments, volume), tree-ensemble velocity
(StandardScaler + RandomForestRegressor(100, max_samples=300) per iteration, trained on the host
in the untimed setup), 50 evolution iterations.

A "step" is ONE segmentation pass = (re-)initialise u0 on the device + 50 iterations of
features -> forest velocity -> upwind update -> band rebuild (fast marching) -> compaction,
i.e. the full body of lsml/core/model.py:378-398.  The metric counts the band voxels that enter
every iteration.  `value` is measured with the normalised image planes already resident in
HBM; `e2e` goes through the public API (LevelSetMachineLearning.segment_batch) with the raw
image in pinned HOST memory (H2D copy, normalisation, Gaussian/edge planes, evolution, D2H of
the final mask inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 256]
N > 1: torchrun, one volume per rank (independent items, no data-path collective; weak scaling).

Other BASELINE.json configurations (parity-test cases; measured for DESIGN.md / profiles/, not
the bench line the driver reads) run with --workload:
  c4   64^3 crops, --items per GPU (default 512) in one engine (configs[3]; weak scaling,
       no collective)
  c2   101^2 images, 1024 in total sharded over the GPUs (configs[1]; strong scaling)
  c5   ONE --size^3 volume slab-decomposed over the GPUs (configs[4]; NCCL halo exchange +
       all-reduce, strong scaling; default size 512, 1024 for the full configuration)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ITERS = 50
_DIAG = {}          # objects the watchdog inspects when a run hangs
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu captures
NCU_TRAFFIC_MARCH_BYTES = 16.951296e6 + 5.875200e6
NCU_TRAFFIC_STENCIL_BYTES = 354.348800e6 + 122.087424e6
BAND = 3.0
STEP = 0.35
INIT_RADIUS = 40.0
SPECS = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
          ("band_mean", 0.0), ("band_mean", 2.0), ("band_std", 0.0), ("band_std", 2.0),
          ("dist_to_com",), ("moments", (0, 1, 2), (1, 2)), ("size",)])


def make_features():
    from lsml_b200.feature import (DistanceToCenterOfMass, Moments, Size,
                                   get_basic_image_features)
    return (get_basic_image_features(ndim=3, sigmas=[0, 2]) +
            [DistanceToCenterOfMass(ndim=3), Moments(ndim=3, axes=(0, 1, 2), orders=(1, 2)),
             Size(ndim=3)])


def make_volume(n, seed):
    """Sphere with a slab removed, noise added (after lsml/data/dim3/hamburger.py:7-79, written
    independently; float64).  Returns (img, radius, centre)."""
    rs = np.random.RandomState(seed)
    scale = n / 256.0
    radius = rs.uniform(60, 90) * scale
    c = n / 2.0 + rs.uniform(-4, 4, size=3) * scale
    ax = [np.arange(n, dtype=np.float64) - c[a] for a in range(3)]
    r2 = ax[0][:, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[2][None, None, :] ** 2
    nrm = rs.randn(3)
    nrm /= np.linalg.norm(nrm)
    plane = np.abs(nrm[0] * ax[0][:, None, None] + nrm[1] * ax[1][None, :, None] +
                   nrm[2] * ax[2][None, None, :])
    img = ((r2 <= radius * radius) & (plane > 5 * scale)).astype(np.float64)
    img += 0.25 * rs.standard_normal(img.shape)
    return img, radius, c


def target_distance_on_band(band_idx, n, radius, c):
    """Ground-truth signed distance to the enclosing sphere at the band voxels (the regression
    target, fit_job_handler.py:371-375)."""
    k = band_idx % n
    j = (band_idx // n) % n
    i = band_idx // (n * n)
    return radius - np.sqrt((i - c[0]) ** 2 + (j - c[1]) ** 2 + (k - c[2]) ** 2)


def train_regressors(eng, u0, n, radius, c, seed, n_iters=N_ITERS, log=None):
    """Untimed setup: one RandomForest per iteration, trained on the host from the device
    feature matrix (the reference's fit loop, fit_job_handler.py:381-428, without its files)."""
    import warnings
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    rs = np.random.RandomState(seed)
    eng.set_level_sets(u0)
    regs = []
    for it in range(n_iters):
        nb = eng.n_band
        X = eng.compute_features().cpu().numpy()
        idx = eng.band_idx[:nb].cpu().numpy()
        y = target_distance_on_band(idx, n, radius, c)
        sub = rs.choice(nb, size=min(nb, 20000), replace=False)
        reg = Pipeline([("standardscaler", StandardScaler()),
                        ("regressor", RandomForestRegressor(n_estimators=100, max_samples=300,
                                                            random_state=rs))])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            reg.fit(X[sub], y[sub])
        regs.append(reg)
        eng.step(reg, STEP)
        if log and it % 10 == 0:
            log("  trained iteration %d (band %d)" % (it, nb))
    return regs


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag = index, [], False
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for q, nm in enumerate(names)
                   if any(len(s) > 2 + q and s[2 + q].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, timed on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_run(n, seed, steps, warmup, regs=None, budget_s=150.0, log=None):
    """Times the reference's CPU path (oracle port: numpy/scipy feature work with the
    reference's work pattern + sklearn predict + the reference's own compiled C for the upwind
    gradient when oracle/_ref exists + the sequential fast-marching restatement) for
    `warmup + steps` evolution iterations.  Each step is ONE iteration on a centred crop of the
    volume whose size is chosen so that the whole run fits `budget_s`."""
    import warnings
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    # crop size: the reference's iteration cost is ~1 us per grid voxel per feature pass
    est_per_voxel = 1.0e-6
    total_steps = max(1, steps + warmup)
    crop = n
    while crop > 64 and est_per_voxel * crop ** 3 * total_steps > budget_s:
        crop //= 2
    img, radius, c = make_volume(n, seed)
    lo = (n - crop) // 2
    img = np.ascontiguousarray(img[lo:lo + crop, lo:lo + crop, lo:lo + crop])
    c = c - lo
    dx = np.ones(3)
    img_ = O.normalize_image(img)
    r_init = min(INIT_RADIUS * n / 256.0, 0.3 * crop)
    u = O.ball_init(img.shape, r_init, dx)
    dist, mask = O.distance_transform(u, BAND, dx)
    if regs is None:
        from sklearn.ensemble import RandomForestRegressor
        from sklearn.pipeline import Pipeline
        from sklearn.preprocessing import StandardScaler
        X = O.feature_matrix(SPECS, u, img_, dist, mask, dx)
        idx = np.flatnonzero(mask.ravel())
        y = target_distance_on_band(idx, crop, radius, c)
        rs = np.random.RandomState(seed)
        reg = Pipeline([("standardscaler", StandardScaler()),
                        ("regressor", RandomForestRegressor(n_estimators=100, max_samples=300,
                                                            random_state=rs))])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            reg.fit(X, y)
        regs = [reg]
    use_ref = O.reference_lib() is not None
    times, counts = [], []
    for it in range(total_steps):
        nb = int(mask.sum())
        t1 = time.perf_counter()
        u, dist, mask, _ = O.evolve_step(u, img_, dist, mask, dx, SPECS, regs[it % len(regs)],
                                         STEP, BAND, use_reference=use_ref)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append(t2 - t1)
            counts.append(nb)
        if log:
            log("  cpu iteration %d: band %d, %.2f s" % (it, nb, t2 - t1))
    return {"seconds": float(sum(times)), "band_voxel_iterations": int(sum(counts)),
            "steps": len(times), "crop": crop, "kind": "port",
            "gradient_kernel": "oracle/_ref (reference C)" if use_ref else "oracle C port",
            "setup_s": time.perf_counter() - t0 - float(sum(times))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = cpu_reference_run(args.size, 1234, args.steps, args.warmup, budget_s=150.0)
    value = res["band_voxel_iterations"] / res["seconds"]
    cores = 1
    sample = ("%d evolution iterations (features+RF-100 velocity+update+band rebuild) on the "
              "centred %d^3 crop of the %d^3 volume, single thread (the reference is "
              "single-threaded)" % (res["steps"], res["crop"], args.size))
    line = {
        "impl": "reference", "metric": "narrow-band voxel-iterations/sec (feat+velocity+update+band rebuild)",
        "value": value, "unit": "band-voxel-iterations/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(1, res["steps"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3: 256^3 volume, RF-100 velocity, band=3, F=16 (CPU arm: one "
                               "iteration per step on a crop)", "crop": res["crop"]},
        "cpu_baseline": {"value": value, "unit": "band-voxel-iterations/s", "cores": cores,
                         "kind": res["kind"], "sample": sample,
                         "gradient_kernel": res["gradient_kernel"]},
        "e2e": {"value": value, "unit": "band-voxel-iterations/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# other configurations (--workload c2 | c4 | c5): same metric, same JSON line, shorter evidence
# --------------------------------------------------------------------------------------------
def _blob_items(count, shape, seed, r_lo, r_hi, jitter):
    """`count` items of `shape`: a noisy ball (radius in [r_lo, r_hi] voxels, jittered centre)."""
    rs = np.random.RandomState(seed)
    nd = len(shape)
    ax = [np.arange(s, dtype=np.float64) for s in shape]
    imgs = np.empty((count,) + tuple(shape))
    for b in range(count):
        c = np.array(shape) / 2.0 + rs.uniform(-jitter, jitter, size=nd)
        rad = rs.uniform(r_lo, r_hi)
        r2 = 0.0
        for a in range(nd):
            sl = [None] * nd
            sl[a] = slice(None)
            r2 = r2 + (ax[a][tuple(sl)] - c[a]) ** 2
        imgs[b] = (r2 < rad * rad).astype(np.float64) + 0.2 * rs.standard_normal(shape)
    return imgs


def harris_response(img, sigma):
    """The user-side Python of examples/gestalt2d/corner_feature.py:
    corner_harris(gaussian_filter(img, sigma)) with skimage's defaults (method 'k', k = 0.05,
    structure-tensor sigma 1, Sobel derivatives, constant boundary) written with scipy because
    scikit-image is not installed here.  Runs ONCE per image (u-independent)."""
    from scipy import ndimage as ndi
    sm = ndi.gaussian_filter(img, sigma) if sigma > 0 else img
    dr = ndi.sobel(sm, axis=0, mode="constant", cval=0)
    dc = ndi.sobel(sm, axis=1, mode="constant", cval=0)
    arr = ndi.gaussian_filter(dr * dr, 1.0, mode="constant", cval=0)
    arc = ndi.gaussian_filter(dr * dc, 1.0, mode="constant", cval=0)
    acc = ndi.gaussian_filter(dc * dc, 1.0, mode="constant", cval=0)
    return arr * acc - arc ** 2 - 0.05 * (arr + acc) ** 2


def gestalt32_specs():
    """The F = 32 feature list of examples/gestalt2d/main.py:27-36: basic image features at
    sigma 0..3 (16), basic shape features (8), CornerFeature x4, PrevIterFeature x4."""
    import functools
    sig = (0.0, 1.0, 2.0, 3.0)
    specs = [(k, s) for k in ("image_sample", "image_edge", "band_mean", "band_std") for s in sig]
    specs += [("size",), ("boundary_size",), ("isoperimetric",), ("moments", (0, 1), (1, 2)),
              ("dist_to_com",)]
    specs += [("user_plane", "corner-reponse-sigma=%.2f" % s,
               functools.partial(lambda img, dx, s: harris_response(img, s), s=s)) for s in sig]
    specs += [("smooth_u", s) for s in sig]
    return specs


def _fit_forest(X, y, seed):
    import warnings
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    rs = np.random.RandomState(seed)
    sub = rs.choice(X.shape[0], size=min(X.shape[0], 20000), replace=False)
    reg = Pipeline([("standardscaler", StandardScaler()),
                    ("regressor", RandomForestRegressor(n_estimators=100, max_samples=300,
                                                        random_state=rs))])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        reg.fit(X[sub], y[sub])
    return reg


def run_other_workload(args):
    """c2 / c4: a batch of independent items per rank in ONE engine; c5: one volume over all
    ranks.  One step = `iters` evolution iterations from the initial level sets."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1 or args.workload == "c5":
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    from lsml_b200 import _lib
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.regressors import export_regressor
    from lsml_b200.core.sharding import shard_range
    wl = args.workload
    t_start = time.perf_counter()
    if wl == "c5":
        from sklearn.linear_model import LinearRegression
        from lsml_b200.core.slab import SlabVolume, TorchCollective
        n = args.size or 512
        iters = 20
        shape = (n, n, n)
        specs = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0),
                  ("image_edge", 2.0), ("band_mean", 2.0), ("band_std", 2.0), ("dist_to_com",),
                  ("moments", (0, 1, 2), (1, 2)), ("size",)])
        ax = np.arange(n, dtype=np.float64) - n / 2.0
        vol = SlabVolume(shape, specs=specs, device=device, collective=TorchCollective())
        # every rank only materialises the planes it holds (owned + halo); the noise of a plane
        # depends on the plane's index only, so that halo planes equal the neighbour's planes
        sl = vol.ranks[0].local_slice()
        a0 = ax[sl]
        r = np.sqrt((a0[:, None, None] / 0.30) ** 2 + (ax[None, :, None] / 0.26) ** 2 +
                    (ax[None, None, :] / 0.22) ** 2) / n          # ellipsoid 300/260/220 at 1024
        part = (r < 1.0).astype(np.float64)
        for q, plane in enumerate(range(sl.start, sl.stop)):
            part[q] += 0.25 * np.random.RandomState(7000 + plane).standard_normal((n, n))
        u0_part = np.where(np.sqrt(a0[:, None, None] ** 2 + ax[None, :, None] ** 2 +
                                   ax[None, None, :] ** 2) < 0.146 * n, 1.0, -1.0)
        del r
        vol.load_image_parts([part], normalize=True)
        X = np.random.RandomState(0).randn(64, 14)
        w = np.zeros(14); w[1] = 0.8
        reg = export_regressor(LinearRegression().fit(X, X @ w + 0.15), device)
        u0_t = torch.from_numpy(u0_part).to(device)

        def one_pass():
            vol.set_level_sets_parts([u0_t])
            tot = 0
            for _ in range(iters):
                tot += vol.step(reg, 0.4)
            return tot
        units_are_global = True
        desc = ("C5: ONE synthetic %d^3 ellipsoid volume slab-decomposed over %d GPU(s) along axis "
                "0, linear velocity, F=14, band=3, %d iterations per step; NCCL: 2 P2P halo "
                "exchanges per round, ~3 rounds per rebuild, 3 small all-reduces per iteration"
                % (n, world, iters))
        scaling = "strong"
    else:
        if wl == "c4":
            # (verified with 512 crops per GPU; a 2048-per-GPU run on 4 GPUs did not finish within
            # its 25-minute limit in round 1 and is still to be diagnosed -- DESIGN.md 8)
            shape, per_rank, iters = (64, 64, 64), args.items or 512, 30
            lo, hi = rank * per_rank, (rank + 1) * per_rank
            imgs = _blob_items(per_rank, shape, 10000 + rank, 6, 18, 8)
            r_init, scaling = 5.0, "weak"
            specs = SPECS
            total_items = per_rank * world
        else:
            shape, total_items, iters = (101, 101), args.items or 1024, 10
            lo, hi = shard_range(total_items, rank, world)
            imgs = _blob_items(hi - lo, shape, 20000 + rank, 15, 35, 10)
            r_init, scaling = 10.0, "strong"
            specs = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0),
                      ("image_edge", 2.0), ("band_mean", 0.0), ("band_mean", 2.0),
                      ("band_std", 0.0), ("band_std", 2.0), ("dist_to_com",),
                      ("moments", (0, 1), (1, 2)), ("size",), ("boundary_size",),
                      ("isoperimetric",)])
            if args.features == "gestalt32":
                specs = gestalt32_specs()
        nd = len(shape)
        vlog = (lambda m: print("[rank %d %7.1f s] %s" % (rank, time.perf_counter() - t_start, m),
                                file=sys.stderr, flush=True)) if args.verbose else (lambda m: None)
        vlog("items generated")
        eng = BandEngine(shape, hi - lo, dx=np.ones(nd), band=BAND, specs=specs, device=device)
        _DIAG["engine"] = eng
        eng.load_images(imgs, normalize=True)
        torch.cuda.synchronize(); vlog("images loaded, planes computed")
        g = np.indices(shape).astype(np.float64)
        rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(nd)))
        u0 = torch.from_numpy(np.where(rc < r_init, 1.0, -1.0)).to(device).expand(
            (hi - lo,) + tuple(shape)).contiguous()
        eng.set_level_sets(u0)
        torch.cuda.synchronize(); vlog("initial band: %d voxels, counters %r" % (eng.n_band, eng.fmm_counters()))
        X = eng.compute_features().cpu().numpy()
        y = X[:, 1] + 0.2 * np.random.RandomState(1).standard_normal(X.shape[0])
        reg = export_regressor(_fit_forest(X, y, 5), device)     # one forest reused every iteration
        vlog("forest trained")

        def one_pass():
            eng.set_level_sets(u0)
            tot = 0
            for it in range(iters):
                tot += eng.step(reg, 0.4)
                if args.verbose:
                    torch.cuda.synchronize()
                    vlog("iteration %d: band %d, %r" % (it, eng.n_band, eng.fmm_counters()))
            eng.check_converged()
            return tot
        units_are_global = False
        desc = ("%s: %d synthetic %s items (%d on this rank, one engine per GPU), RF-100 velocity, "
                "F=%d, band=3, %d iterations per step; no data-path collective"
                % (wl.upper(), total_items, "x".join(map(str, shape)), hi - lo, len(X[0]), iters))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(1, args.warmup)):
        one_pass()
    barrier()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    units = 0
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            units += one_pass()
        ev1.record()
        barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    cnt = torch.tensor([units, _lib.launch_count() - l0], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if units_are_global:
            dist.all_reduce(cnt[1:], op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({
            "metric": "narrow-band voxel-iterations/sec (feat+velocity+update+band rebuild)",
            "value": int(cnt[0]) / (ms * 1e-3), "unit": "band-voxel-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(1, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc,
                       "band_voxel_iterations_per_step": int(cnt[0]) // max(1, args.steps)},
            "gpu_launches": int(cnt[1]), "clocks": clocks.summary()}))
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--size", type=int, default=None, help="volume edge (c3: 256, c5: 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--items", type=int, default=0, help="c2: total images, c4: crops per GPU")
    ap.add_argument("--features", default="basic", choices=["basic", "gestalt32"],
                    help="c2: `gestalt32` = the full F = 32 list of examples/gestalt2d (corner and "
                         "previous-iterate features through the user-feature registry)")
    ap.add_argument("--watchdog", type=float, default=1500.0,
                    help="seconds after which the process kills itself (a hung kernel cannot be "
                         "interrupted from Python; exiting releases the GPU box instead of "
                         "holding it until an outer limit); 0 disables")
    args = ap.parse_args()
    if args.watchdog > 0:
        import threading

        def _abort():
            print("bench.py: watchdog fired after %.0f s -- aborting (exit 124)" % args.watchdog,
                  file=sys.stderr, flush=True)
            try:      # where is the host, and (from a second stream) where is the band rebuild?
                import faulthandler
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                eng = _DIAG.get("engine")
                if eng is not None and getattr(eng, "_ws_fmm", None) is not None:
                    import torch
                    with torch.cuda.stream(torch.cuda.Stream(device=eng.device)):
                        st = eng._fmm_status.to("cpu", non_blocking=False).tolist()
                    print("fmm status {failed, barriers, live sweep, rix, begin, end}: %r" % (st,),
                          file=sys.stderr, flush=True)
            except Exception as exc:            # diagnostics only
                print("watchdog diagnostics failed: %r" % (exc,), file=sys.stderr, flush=True)
            os._exit(124)
        timer = threading.Timer(args.watchdog, _abort)
        timer.daemon = True
        timer.start()
    if args.workload != "c3" and args.impl != "reference":
        return run_other_workload(args)
    args.size = args.size or 256
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    log = (lambda m: print("[rank %d] %s" % (rank, m), file=sys.stderr, flush=True)) \
        if args.verbose else None

    from lsml_b200 import LevelSetMachineLearning, _lib
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.model import _ball_u0
    from lsml_b200.initializer import BallInitializer

    n = args.size
    seed = 1234 + rank                      # every rank evolves its own volume
    img, radius, c = make_volume(n, seed)
    dx = np.ones(3)
    r_init = INIT_RADIUS * n / 256.0
    eng = BandEngine((n, n, n), 1, dx=dx, band=BAND, specs=SPECS, device=device)
    eng.load_images(img[None], normalize=True)
    centre = 0.5 * np.array([n, n, n], dtype=np.float64)
    u0 = _ball_u0(eng, centre[None], [r_init]).clone()
    regs = train_regressors(eng, u0, n, radius, c, seed, log=log)
    model = LevelSetMachineLearning.from_regressors(make_features(), BallInitializer(radius=r_init),
                                                    regs, step=STEP, band=BAND)
    dev_regs = [model._device_model(i, device) for i in range(N_ITERS)]

    def one_pass():
        eng.set_level_sets(u0)
        total = 0
        for i in range(N_ITERS):
            total += eng.step(dev_regs[i], STEP)
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        one_pass()
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bvi = 0
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            bvi += one_pass()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0

    # ---- end to end through the public API, host buffers --------------------------------------
    img_pinned = torch.from_numpy(img[None]).pin_memory()
    out_host = torch.empty((1, n, n, n), dtype=torch.bool).pin_memory()

    def e2e_pass():
        u_fin, m_fin = model.segment_batch(img_pinned, dx=dx, n_iters=N_ITERS, device=device)
        out_host.copy_(u_fin > 0, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_pass()
    barrier()
    e2e_s = time.perf_counter() - t0
    seg_voxels = int(out_host.sum())

    # ---- per-phase evidence: one more pass with CUDA-event brackets around every phase ---------
    # (events on the launching stream; the brackets add a few microseconds per phase, which is
    # why this is a separate pass and not the timed region)
    eng.timers = {}
    eng.set_level_sets(u0)
    cand_sum, sweeps_sum, replay_sum, not_converged = 0, 0, 0, 0
    for i in range(N_ITERS):
        eng.step(dev_regs[i], STEP)
        if i % 5 == 4:                      # counters of a sample of the rebuilds (host sync)
            ctr = eng.fmm_counters()
            cand_sum += ctr["n_cand"]; sweeps_sum += ctr["sweeps"]; replay_sum += ctr["replays"]
            not_converged += 0 if ctr["converged"] else 1
    phases = eng.phase_times_ms()
    eng.timers = None
    n_sampled = N_ITERS // 5
    phase_total = sum(v[1] for v in phases.values())
    rebuild_n, rebuild_ms = phases.get("rebuild", (1, 0.0))
    rebuild_avg_ms = rebuild_ms / max(1, rebuild_n)
    avg_cand = cand_sum / max(1, n_sampled)
    peak, peak_src = measured_peak_gbs()
    # DOMINANT kernel of the step: the band rebuild (fmm_march_kernel + its 4 small helper
    # launches).  Algorithmic bytes per candidate voxel: read u (8) + write the
    # distance (8) + write the mask byte (1) = 17 B.  The kernel is bound by the dependency depth
    # of the fast-marching order (latency), not by HBM -- the fraction below says exactly that.
    rebuild_bytes = 17.0 * avg_cand
    rebuild_gbs = rebuild_bytes / (rebuild_avg_ms * 1e-3) / 1e9 if rebuild_avg_ms > 0 else 0.0

    # ---- the HBM-bound kernel named by north_star: dense stencil at the same grid ---------------
    from lsml_b200.gradient import masked_gradient as mg
    A = eng.u[0]
    nu = torch.randn_like(A)
    full = torch.ones_like(A, dtype=torch.uint8)
    out = torch.zeros_like(A)
    shape_c = _lib.int_array((n, n, n))
    dx_c = _lib.double_array(dx)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > L2 (126 MB)

    def stencil_call():
        _lib.check(_lib.lib.lsml_gmag_os_f64(3, shape_c, _lib.c_i64(1), _lib.ptr(A), _lib.ptr(full),
                                             _lib.ptr(nu), _lib.ptr(out), dx_c,
                                             _lib.current_stream()), "gmag_os")
    for _ in range(3):
        stencil_call()
    st_ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stencil_call()
        b.record()
        torch.cuda.synchronize()
        st_ms.append(a.elapsed_time(b))
    st_ms = float(np.median(st_ms))
    stencil_bytes = 25.0 * n ** 3          # 8 (A) + 1 (mask) + 8 (nu) + 8 (out) per voxel
    stencil_gbs = stencil_bytes / (st_ms * 1e-3) / 1e9

    # ---- aggregate over ranks ------------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=device)
    cnt = torch.tensor([bvi, launches], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    bvi_all, launches_all = int(cnt[0]), int(cnt[1])
    value = bvi_all / (ms_max * 1e-3)
    e2e_value = (bvi_all / args.steps * e2e_steps) / (e2e_ms_max * 1e-3)

    if rank == 0:
        line = {
            "metric": "narrow-band voxel-iterations/sec (feat+velocity+update+band rebuild)",
            "value": value, "unit": "band-voxel-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C3: one synthetic %d^3 volume per GPU, RF-100 tree-ensemble "
                                   "velocity, band=3, dx=1, F=16, %d iterations per step "
                                   "(one step = one full segmentation pass)" % (n, N_ITERS),
                       "band_voxel_iterations_per_step": bvi_all // max(1, args.steps),
                       "l2": "working set > L2: 5 x %d MB planes + u; stencil probe flushes L2 "
                             "between launches" % (8 * n ** 3 // 2 ** 20),
                       "final_segmentation_voxels": seg_voxels,
                       "rebuilds_not_converged": "%d of %d sampled" % (not_converged, n_sampled)},
            "e2e": {"value": e2e_value, "unit": "band-voxel-iterations/s",
                    "h2d_bytes_per_step": int(img_pinned.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel()),
                    "api": "LevelSetMachineLearning.segment_batch(host image) -> host mask",
                    "ms_per_step": e2e_ms_max / e2e_steps},
            "gpu_launches": launches_all,
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm",
                         "kernel": "band rebuild: fmm_march_kernel<3> (+ clear/prepare/init/seed), "
                                   "the dominant kernel of the step",
                         "achieved": rebuild_gbs, "peak": peak, "unit": "GB/s",
                         "frac": rebuild_gbs / peak, "traffic": NCU_TRAFFIC_MARCH_BYTES,
                         "traffic_source": "profiles/r1_ncu_fmm_march_r1.txt (one ncu --set full "
                                           "capture of iteration 12 of tools/iter_probe.py)",
                         "peak_source": peak_src, "ms": rebuild_avg_ms, "share_of_step": rebuild_ms / max(phase_total, 1e-9),
                         "algorithmic_bytes_per_launch": rebuild_bytes,
                         "candidates_per_launch": avg_cand,
                         "sweeps_per_launch": sweeps_sum / max(1, n_sampled),
                         "replays_per_launch": replay_sum / max(1, n_sampled),
                         "note": "latency-bound (dependency depth of the fast-marching order x "
                                 "replay latency), not HBM-bound: DESIGN.md 4.6"},
            "roofline_stencil": {"bound": "hbm",
                                 "kernel": "gmag_os_marchv_kernel<double> (dense %d^3, full mask, "
                                           "L2 flushed between launches)" % n,
                                 "achieved": stencil_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": stencil_gbs / peak, "traffic": NCU_TRAFFIC_STENCIL_BYTES,
                                 "traffic_source": "profiles/r1_ncu_stencil_r1.txt",
                                 "algorithmic_bytes_per_launch": stencil_bytes,
                                 "peak_source": peak_src, "ms": st_ms},
            "phases_ms_per_step": {k: v[1] for k, v in phases.items()},
        }
        # SURVEY.md 8(d) region (A): features + velocity + update with the band given -- the part
        # BASELINE.json's metric names; `value` above is region (B), the full step with band
        # rebuild and compaction.  From the per-phase device times of the evidence pass (one
        # step on rank 0's volume).
        ms_a = sum(phases.get(k, (0, 0.0))[1] for k in ("features", "predict", "update"))
        if ms_a > 0:
            line["region_a"] = {"value": (bvi_all / world / max(1, args.steps)) / (ms_a * 1e-3),
                                "unit": "band-voxel-iterations/s", "ms_per_step": ms_a,
                                "what": "features + velocity + update, band given, one GPU"}
        if not args.no_cpu_baseline:
            res = cpu_reference_run(n, 1234, 1, 0, regs=None, budget_s=25.0, log=log)
            v = res["band_voxel_iterations"] / res["seconds"]
            line["cpu_baseline"] = {
                "value": v, "unit": "band-voxel-iterations/s", "cores": 1, "kind": res["kind"],
                "sample": "1 evolution iteration (features+RF-100+update+band rebuild) on the "
                          "centred %d^3 crop, single thread, %.1f s" % (res["crop"], res["seconds"]),
                "gradient_kernel": res["gradient_kernel"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
