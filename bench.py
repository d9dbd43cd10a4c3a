#!/usr/bin/env python
"""bench.py -- narrow-band voxel-iterations / s of the lsml level-set evolution hot path.

Headline workload (BASELINE.json configs[2], the configuration the north_star target is quoted
on; SURVEY.md section 8d "C3"): one synthetic 256^3 volume (a noisy "hamburger" sphere with a slab
cut), +-1 ball initialisation r=40, band=3, dx=1, F=16 features (8 image features at sigma in
{0,2}, distance to centre of mass, 6 moments, volume), tree-ensemble velocity (StandardScaler +
RandomForestRegressor(100, max_samples=300) per iteration, trained on the host in the untimed
setup), 50 evolution iterations.

A "step" is ONE segmentation pass = (re-)initialise u0 on the device + 50 iterations of
features -> forest velocity -> upwind update -> band rebuild (fast marching) -> compaction,
i.e. the full body of lsml/core/model.py:378-398.  The metric counts the band voxels that enter
every iteration.  `value` is measured with the normalised image planes already resident in
HBM; `e2e` goes through the public API (LevelSetMachineLearning.segment_batch) with the raw
image in pinned HOST memory (H2D copy, normalisation, Gaussian/edge planes, evolution, D2H of
the final mask inside the timed region); `e2e_segment` is the reference-signature
`segment(img)` that materialises all 51 iterates on the host.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun): the headline line is the same volume evolved by every rank (replicas; weak
scaling).  The SHARDED workloads north_star names are timed in the same run, at every N, and
reported as sub-objects of the JSON line:
  "c4"  BASELINE configs[3]: 16 384 crops of 64^3, strong scaling -- the fixed total is split
        into contiguous index ranges (lsml_b200/core/sharding.py), every rank evolves its range in
        engine batches of 2048 crops, no data-path collective;
  "c5"  BASELINE configs[4]: ONE 1024^3 volume slab-decomposed along axis 0 over the N GPUs
        (lsml_b200/core/slab.py: NCCL halo exchange + all-reduce of the shape statistics), strong
        scaling with the 1-GPU run of the same volume as denominator.
`--sections none` skips them; `--workload c2|c4|c5` runs one of them alone (manual runs).
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
BAND = 3.0
STEP = 0.35
INIT_RADIUS = 40.0
SPECS = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
          ("band_mean", 0.0), ("band_mean", 2.0), ("band_std", 0.0), ("band_std", 2.0),
          ("dist_to_com",), ("moments", (0, 1, 2), (1, 2)), ("size",)])
METRIC = "narrow-band voxel-iterations/sec (feat+velocity+update+band rebuild)"
UNIT = "band-voxel-iterations/s"
C4_TOTAL, C4_CHUNK, C4_BLOCK, C4_ITERS = 16384, 2048, 256, 30
C5_SIZE, C5_ITERS = 1024, 20


def workload_config(n):
    """`config` of BOTH arms (the CPU arm times the same workload on the host cores)."""
    return {"workload": "C3: one synthetic %d^3 volume per GPU, RF-100 tree-ensemble velocity, "
                        "band=3, dx=1, F=16, %d iterations per segmentation pass" % (n, N_ITERS),
            "grid": [n, n, n], "iterations": N_ITERS, "features": 16, "regressor": "RF-100",
            "l2": "working set > L2: 5 x %d MB planes + u; the stencil probe flushes L2 between "
                  "launches" % (8 * n ** 3 // 2 ** 20)}


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


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed
    summary of the ncu --set full captures of THIS round (profiles/r2_ncu_traffic.json, written
    by tools/ncu_summary.py from the .ncu-rep files); null when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
            d = json.load(f)
        e = d[kernel_key]
        return float(e["dram_bytes_per_launch"]), "profiles/r2_ncu_traffic.json (%s)" % e.get("source", "")
    except Exception:
        return None, "no ncu capture committed for this kernel"


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, timed on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_run(n, seed, steps, warmup, budget_s=170.0, log=None, literal_gap=False):
    """Times the reference's CPU path (oracle port: numpy/scipy feature work with the
    reference's work pattern + sklearn predict + the reference's own compiled C for the upwind
    gradient when oracle/_ref exists + the sequential fast-marching restatement) on the SAME
    n^3 volume as the GPU arm.  Each step is ONE evolution iteration of that volume (the CPU
    needs ~8 s for it at 256^3); the timed steps stop early when `budget_s` is used up, and the
    line reports how many ran."""
    import warnings
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    img, radius, c = make_volume(n, seed)
    dx = np.ones(3)
    img_ = O.normalize_image(img)
    u = O.ball_init(img.shape, INIT_RADIUS * n / 256.0, dx)
    dist, mask = O.distance_transform(u, BAND, dx)
    from sklearn.ensemble import RandomForestRegressor
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    X = O.feature_matrix(SPECS, u, img_, dist, mask, dx)
    idx = np.flatnonzero(mask.ravel())
    y = target_distance_on_band(idx, n, radius, c)
    rs = np.random.RandomState(seed)
    sub = rs.choice(len(idx), size=min(len(idx), 20000), replace=False)
    reg = Pipeline([("standardscaler", StandardScaler()),
                    ("regressor", RandomForestRegressor(n_estimators=100, max_samples=300,
                                                        random_state=rs))])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        reg.fit(X[sub], y[sub])
    del X
    use_ref = O.reference_lib() is not None
    times, counts = [], []
    gap = None
    t_loop = time.perf_counter()
    for it in range(max(1, steps + warmup)):
        nb = int(mask.sum())
        t1 = time.perf_counter()
        u, dist, mask, _ = O.evolve_step(u, img_, dist, mask, dx, SPECS, reg, STEP, BAND,
                                         use_reference=use_ref)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append(t2 - t1)
            counts.append(nb)
        if log:
            log("  cpu iteration %d: band %d, %.2f s" % (it, nb, t2 - t1))
        if literal_gap and gap is None:
            # the same rebuild input through the upstream-literal scikit-fmm restatement
            _, lmask = O.distance_transform(u, BAND, dx, literal=True)
            gap = {"mask_mismatch_voxels": int((lmask != mask).sum()), "band_voxels": int(mask.sum()),
                   "what": "band of the first evolved level set: product marcher vs "
                           "oracle/skfmm_literal.c (scikit-fmm restated from memory; see "
                           "profiles/r2_band_gap_vs_upstream_literal.json for the campaign)"}
        if times and time.perf_counter() - t_loop > budget_s:
            break
    return {"seconds": float(sum(times)), "band_voxel_iterations": int(sum(counts)),
            "steps": len(times), "kind": "port", "gap": gap,
            "gradient_kernel": "oracle/_ref (reference C)" if use_ref else "oracle C port",
            "setup_s": time.perf_counter() - t0 - float(sum(times))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.size
    res = cpu_reference_run(n, 1234, args.steps, args.warmup, budget_s=170.0)
    value = res["band_voxel_iterations"] / res["seconds"]
    sample = ("%d of the %d requested timed steps, each ONE evolution iteration (features + "
              "RF-100 velocity + update + band rebuild) of the same %d^3 volume, single thread "
              "(the reference is single-threaded Python/numpy/C)" % (res["steps"], args.steps, n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / max(1, res["steps"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": res["kind"],
                         "sample": sample, "timed_steps": res["steps"],
                         "gradient_kernel": res["gradient_kernel"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# helpers shared by the sharded sections
# --------------------------------------------------------------------------------------------
class Ctx:
    """rank / world / device and the timing protocol: barrier + synchronize on both sides, CUDA
    events on the launching stream, max over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)

    def init_pg(self, force=False):
        if (self.world > 1 or force) and not self.dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
            self.dist.init_process_group("nccl", rank=self.rank, world_size=self.world,
                                         device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, reps):
        """Run fn() `reps` times between barriers; returns (max-over-ranks ms, list of results)."""
        torch = self.torch
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = [fn() for _ in range(reps)]
        ev1.record()
        self.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), out

    def sum_ints(self, values):
        t = self.torch.tensor([int(v) for v in values], dtype=self.torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(v) for v in t.tolist()]


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


def _blob_items(count, shape, seed, r_lo, r_hi, jitter):
    """`count` items of `shape` on the HOST: a noisy ball (radius in [r_lo, r_hi] voxels,
    jittered centre).  (c2 and the CPU-side tools; the c4 section generates on the device.)"""
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


def c4_crop_geometry(index):
    """(centre[3], radius) of crop `index` of the C4 dataset (nodule-like ball of
    radius U[6,18], centre jitter +-8)."""
    rs = np.random.RandomState(10000 + index)
    return 32.0 + rs.uniform(-8, 8, size=3), rs.uniform(6, 18)


def c4_crops_device(lo, hi, device):
    """Crops [lo, hi) of the C4 dataset as a (hi-lo, 64, 64, 64) float64 CUDA tensor: the ball of
    c4_crop_geometry plus N(0, 0.2^2) noise.  The noise of block k (crops [256k, 256k+256)) comes
    from a device generator seeded 5000 + k -- so a crop is the same whichever rank, engine batch
    or GPU count evolves it.  lo and hi must be multiples of 256."""
    import torch
    assert lo % C4_BLOCK == 0 and hi % C4_BLOCK == 0
    n = 64
    out = torch.empty((hi - lo, n, n, n), dtype=torch.float64, device=device)
    ax = torch.arange(n, dtype=torch.float64, device=device)
    for k in range(lo // C4_BLOCK, hi // C4_BLOCK):
        par = np.empty((C4_BLOCK, 4))
        for q in range(C4_BLOCK):
            par[q, :3], par[q, 3] = c4_crop_geometry(k * C4_BLOCK + q)
        p = torch.from_numpy(par).to(device)
        r2 = ((ax[None, :, None, None] - p[:, 0, None, None, None]) ** 2 +
              (ax[None, None, :, None] - p[:, 1, None, None, None]) ** 2 +
              (ax[None, None, None, :] - p[:, 2, None, None, None]) ** 2)
        gen = torch.Generator(device=device)
        gen.manual_seed(5000 + k)
        blk = out[(k * C4_BLOCK - lo):(k * C4_BLOCK - lo) + C4_BLOCK]
        blk.copy_((r2 < (p[:, 3] ** 2)[:, None, None, None]).to(torch.float64))
        blk.add_(torch.randn(blk.shape, dtype=torch.float64, device=device, generator=gen), alpha=0.2)
        del r2
    return out


def c4_initial_level_set(device):
    import torch
    shape = (64, 64, 64)
    g = np.indices(shape).astype(np.float64)
    rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(3)))
    return torch.from_numpy(np.where(rc < 5.0, 1.0, -1.0)).to(device)


def c4_forest(device):
    """One forest for the C4 run, trained (untimed, identically on every rank) on the initial
    band of block 0; returns (host pipeline, device regressor)."""
    import torch
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.regressors import export_regressor
    shape = (64, 64, 64)
    eng0 = BandEngine(shape, C4_BLOCK, dx=np.ones(3), band=BAND, specs=SPECS, device=device)
    eng0.load_images(c4_crops_device(0, C4_BLOCK, device), normalize=True)
    eng0.set_level_sets(c4_initial_level_set(device).expand((C4_BLOCK,) + shape).contiguous())
    X = eng0.compute_features().cpu().numpy()
    y = X[:, 1] + 0.2 * np.random.RandomState(1).standard_normal(X.shape[0])
    host = _fit_forest(X, y, 5)
    del eng0
    torch.cuda.empty_cache()
    return host, export_regressor(host, device)


def section_c4(ctx, total=C4_TOTAL, chunk=C4_CHUNK, log=None):
    """BASELINE configs[3]: `total` crops of 64^3, strong scaling over the ranks, no collective on
    the data path.  Raw crops resident in HBM; the timed pass covers, per engine batch of
    `chunk` crops: normalisation + the Gaussian / edge planes, ball initialisation + first band,
    30 evolution iterations."""
    import torch
    from lsml_b200 import _lib
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.sharding import shard_range
    dev = ctx.device
    log = log or (lambda m: None)
    shape = (64, 64, 64)
    # contiguous ranges in units of 256-crop blocks
    blo, bhi = shard_range(total // C4_BLOCK, ctx.rank, ctx.world)
    lo, hi = blo * C4_BLOCK, bhi * C4_BLOCK
    u0_one = c4_initial_level_set(dev)
    _, reg = c4_forest(dev)
    log("c4: forest trained; generating crops [%d, %d)" % (lo, hi))
    raw = c4_crops_device(lo, hi, dev)
    per_engine = min(chunk, max(hi - lo, C4_BLOCK))
    eng = BandEngine(shape, per_engine, dx=np.ones(3), band=BAND, specs=SPECS, device=dev)
    _DIAG["engine"] = eng
    u0 = u0_one.expand((per_engine,) + shape).contiguous()

    def evolve(a, b):
        if b - a != per_engine:          # ragged last batch: pad with copies of the first crop
            imgs = torch.cat([raw[a:b], raw[a:a + 1].expand((per_engine - (b - a),) + shape)])
        else:
            imgs = raw[a:b]
        eng.load_images(imgs, normalize=True)
        eng.set_level_sets(u0)
        eng.units_dev.zero_()
        for _ in range(C4_ITERS):
            eng.step(reg, 0.4)
        return eng.units_dev.clone()

    def one_pass():
        units = [evolve(a, min(a + per_engine, hi - lo)) for a in range(0, hi - lo, per_engine)]
        return int(torch.stack(units).sum().item()) if units else 0

    if hi > lo:                           # warm-up: one engine batch
        evolve(0, min(per_engine, hi - lo))
    log("c4: warm-up done")
    l0 = _lib.launch_count()
    ms, units = ctx.timed(one_pass, 1)
    launches = _lib.launch_count() - l0
    eng.check_converged()
    tot_units, tot_launch, failed = ctx.sum_ints([units[0], launches, int(eng._fmm_status[0].item())])
    # per-phase device times of one engine batch on rank 0 (CUDA-event brackets; evidence pass)
    phases, upd = None, None
    if hi > lo:
        eng.timers = {}
        nb_batch = int(evolve(0, min(per_engine, hi - lo)).item())
        ph = eng.phase_times_ms()
        eng.timers = None
        phases = {k: v[1] for k, v in ph.items()}
        upd_n, upd_ms = ph.get("update", (1, 0.0))
        if upd_ms > 0:
            peak, peak_src = measured_peak_gbs()
            gbs = 64.0 * nb_batch / (upd_ms * 1e-3) / 1e9        # ~64 B per band voxel
            upd = {"bound": "hbm", "kernel": "band_update_kernel + band_apply_kernel over one engine "
                                             "batch (%d band voxels per iteration on average)"
                                             % (nb_batch // C4_ITERS),
                   "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                   "algorithmic_bytes_per_launch": 64.0 * nb_batch / C4_ITERS,
                   "ms": upd_ms / max(1, upd_n), "peak_source": peak_src}
            # DRAM traffic of the pair at this scale: one ncu capture per kernel at iteration ~20 of
            # a 2048-crop batch (a launch of fewer band voxels than the average: compare the
            # BYTES PER MICROSECOND, 2.6-2.9 TB/s, not the totals)
            tu, su = ncu_traffic("c4_band_update_kernel")
            ta, _ = ncu_traffic("c4_band_apply_kernel")
            upd["traffic"] = (tu + ta) if (tu is not None and ta is not None) else None
            upd["traffic_source"] = su
    del eng, raw
    torch.cuda.empty_cache()
    return {"value": tot_units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": 1,
            "phases_ms_per_engine_batch": phases, "roofline_update": upd,
            "n_gpus": ctx.world, "scaling": "strong",
            "config": {"workload": "C4: %d synthetic 64^3 crops in total (BASELINE configs[3]), "
                                   "contiguous index ranges per rank, engine batches of %d crops, "
                                   "RF-100 velocity, F=16, band=3, %d iterations; no data-path "
                                   "collective; one step = one pass over ALL crops incl. "
                                   "normalisation and image planes" % (total, per_engine, C4_ITERS),
                       "crops_on_rank0": hi - lo},
            "band_voxel_iterations_per_step": tot_units, "gpu_launches": tot_launch,
            "rebuilds_not_converged": failed}


def section_c5(ctx, n=C5_SIZE, log=None):
    """BASELINE configs[4]: ONE n^3 volume slab-decomposed over the ranks along axis 0."""
    import torch
    from sklearn.linear_model import LinearRegression
    from lsml_b200 import _lib
    from lsml_b200.core.regressors import export_regressor
    from lsml_b200.core.slab import SlabVolume, TorchCollective
    dev = ctx.device
    log = log or (lambda m: None)
    ctx.init_pg(force=True)
    shape = (n, n, n)
    specs = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0),
              ("image_edge", 2.0), ("band_mean", 2.0), ("band_std", 2.0), ("dist_to_com",),
              ("moments", (0, 1, 2), (1, 2)), ("size",)])
    vol = SlabVolume(shape, specs=specs, device=dev, collective=TorchCollective())
    _DIAG["engine"] = vol.ranks[0].eng
    # every rank only materialises the planes it holds (owned + halo), on the device; the noise
    # of a plane depends on the plane's index only, so that halo planes equal the neighbour's
    sl = vol.ranks[0].local_slice()
    ax = torch.arange(n, dtype=torch.float64, device=dev) - n / 2.0
    a0 = ax[sl]
    part = torch.empty((a0.numel(), n, n), dtype=torch.float64, device=dev)
    u0_part = torch.empty_like(part)
    gen = torch.Generator(device=dev)
    for q, plane in enumerate(range(sl.start, sl.stop)):
        # ellipsoid with semi-axes 300/260/220 at 1024, ball init r = 150
        e2 = (a0[q] / 0.30) ** 2 + (ax[:, None] / 0.26) ** 2 + (ax[None, :] / 0.22) ** 2
        gen.manual_seed(7000 + plane)
        part[q] = (e2 < float(n) ** 2).to(torch.float64) + 0.25 * torch.randn(
            (n, n), dtype=torch.float64, device=dev, generator=gen)
        r2 = a0[q] ** 2 + ax[:, None] ** 2 + ax[None, :] ** 2
        u0_part[q] = torch.where(r2 < (0.146 * n) ** 2, 1.0, -1.0)
    vol.load_image_parts([part], normalize=True)
    del part
    X = np.random.RandomState(0).randn(64, 14)
    w = np.zeros(14)
    w[1] = 0.8
    reg = export_regressor(LinearRegression().fit(X, X @ w + 0.15), dev)
    rounds = []

    def one_pass():
        vol.set_level_sets_parts([u0_part])
        tot = 0
        for _ in range(C5_ITERS):
            tot += vol.step(reg, 0.4)
            rounds.append(vol.rebuild_rounds)
        return tot

    one_pass()
    log("c5: warm-up done")
    del rounds[:]
    l0 = _lib.launch_count()
    reps = 2
    ms, units = ctx.timed(one_pass, reps)
    launches = _lib.launch_count() - l0
    tot_launch, = ctx.sum_ints([launches])
    r0 = vol.ranks[0]
    seg = ctx.sum_ints([int((r0.eng.u[0, r0.lo:r0.hi] > 0).sum().item())])[0]
    del vol, u0_part, r0
    torch.cuda.empty_cache()
    return {"value": sum(units) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / reps, "steps": reps,
            "ms_per_iteration": ms / reps / C5_ITERS, "n_gpus": ctx.world, "scaling": "strong",
            "config": {"workload": "C5: ONE synthetic %d^3 ellipsoid volume (BASELINE configs[4]) "
                                   "slab-decomposed over %d GPU(s) along axis 0, linear velocity, "
                                   "F=14, band=3, %d iterations per step; NCCL halo exchange of u "
                                   "and of the marcher's T/state planes + all-reduce of the shape "
                                   "statistics" % (n, ctx.world, C5_ITERS)},
            "band_voxel_iterations_per_step": sum(units) // reps, "gpu_launches": tot_launch,
            "exchange_rounds_per_rebuild": float(np.mean(rounds)) if rounds else None,
            "final_segmentation_voxels": seg}


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


C2_SPECS = ([("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0),
             ("image_edge", 2.0), ("band_mean", 0.0), ("band_mean", 2.0),
             ("band_std", 0.0), ("band_std", 2.0), ("dist_to_com",),
             ("moments", (0, 1), (1, 2)), ("size",), ("boundary_size",), ("isoperimetric",)])


def section_c2(ctx, total=1024, features="basic", log=None):
    """BASELINE configs[1]: `total` images of 101^2 sharded over the ranks, 10 iterations."""
    import torch
    from lsml_b200 import _lib
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.regressors import export_regressor
    from lsml_b200.core.sharding import shard_range
    dev = ctx.device
    shape, iters = (101, 101), 10
    lo, hi = shard_range(total, ctx.rank, ctx.world)
    imgs = _blob_items(hi - lo, shape, 20000 + ctx.rank, 15, 35, 10)
    specs = gestalt32_specs() if features == "gestalt32" else C2_SPECS
    eng = BandEngine(shape, hi - lo, dx=np.ones(2), band=BAND, specs=specs, device=dev)
    _DIAG["engine"] = eng
    eng.load_images(imgs, normalize=True)
    g = np.indices(shape).astype(np.float64)
    rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(2)))
    u0 = torch.from_numpy(np.where(rc < 10.0, 1.0, -1.0)).to(dev).expand(
        (hi - lo,) + shape).contiguous()
    eng.set_level_sets(u0)
    X = eng.compute_features().cpu().numpy()
    y = X[:, 1] + 0.2 * np.random.RandomState(1).standard_normal(X.shape[0])
    reg = export_regressor(_fit_forest(X, y, 5), dev)
    nfeat = X.shape[1]

    def one_pass():
        eng.set_level_sets(u0)
        eng.units_dev.zero_()
        for _ in range(iters):
            eng.step(reg, 0.4)
        return int(eng.units_dev.item())

    one_pass()
    l0 = _lib.launch_count()
    reps = 3
    ms, units = ctx.timed(one_pass, reps)
    launches = _lib.launch_count() - l0
    eng.check_converged()
    tot_units, tot_launch = ctx.sum_ints([sum(units), launches])
    eng.timers = {}
    one_pass()
    phases = {k: v[1] for k, v in eng.phase_times_ms().items()}
    eng.timers = None
    return {"value": tot_units / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / reps, "steps": reps,
            "n_gpus": ctx.world, "scaling": "strong", "phases_ms_per_step": phases,
            "config": {"workload": "C2: %d synthetic 101x101 images in total (BASELINE configs[1]), "
                                   "%d on rank 0 in one engine, RF-100 velocity, F=%d (%s), band=3, "
                                   "%d iterations per step; no data-path collective"
                                   % (total, hi - lo, nfeat, features, iters)},
            "band_voxel_iterations_per_step": tot_units // reps, "gpu_launches": tot_launch}


def run_single_workload(args):
    """--workload c2 | c4 | c5: one sharded workload alone, printed as the whole JSON line."""
    ctx = Ctx()
    ctx.init_pg(force=args.workload == "c5")
    t_start = time.perf_counter()
    log = (lambda m: print("[rank %d %7.1f s] %s" % (ctx.rank, time.perf_counter() - t_start, m),
                           file=sys.stderr, flush=True)) if args.verbose else None
    with ClockSampler(ctx.local_rank) as clocks:
        if args.workload == "c4":
            res = section_c4(ctx, total=args.items or C4_TOTAL, log=log)
        elif args.workload == "c5":
            res = section_c5(ctx, n=args.size or C5_SIZE, log=log)
        else:
            res = section_c2(ctx, total=args.items or 1024, features=args.features, log=log)
    if ctx.rank == 0:
        line = {"metric": METRIC, "warmup": 1, "higher_is_better": True, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "clocks": clocks.summary()}
        line.update(res)
        print(json.dumps(line))
    if ctx.dist.is_initialized():
        ctx.dist.destroy_process_group()
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
    ap.add_argument("--size", type=int, default=None, help="volume edge (c3: 256, c5: 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--sections", default="c4,c5",
                    help="sharded workloads timed after the headline (comma list of c4, c5; "
                         "'none' skips them)")
    ap.add_argument("--items", type=int, default=0, help="c2 / c4: total items")
    ap.add_argument("--c5-size", type=int, default=C5_SIZE)
    ap.add_argument("--features", default="basic", choices=["basic", "gestalt32"],
                    help="c2: `gestalt32` = the full F = 32 list of examples/gestalt2d (corner and "
                         "previous-iterate features through the user-feature registry)")
    ap.add_argument("--watchdog", type=float, default=1500.0,
                    help="seconds after which the process kills itself (a hung kernel cannot be "
                         "interrupted from Python; exiting releases the GPU box instead of "
                         "holding it until an outer limit); 0 disables")
    args = ap.parse_args()
    if args.watchdog > 0:
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
    if args.impl == "reference":
        args.size = args.size or 256
        return run_reference_arm(args)
    if args.workload != "c3":
        return run_single_workload(args)
    args.size = args.size or 256

    import torch
    ctx = Ctx()
    ctx.init_pg()
    dist = ctx.dist
    rank, world, device, local_rank = ctx.rank, ctx.world, ctx.device, ctx.local_rank
    t_start = time.perf_counter()
    log = (lambda m: print("[rank %d %7.1f s] %s" % (rank, time.perf_counter() - t_start, m),
                           file=sys.stderr, flush=True)) if args.verbose else None

    from lsml_b200 import LevelSetMachineLearning, _lib
    from lsml_b200.core.engine import BandEngine
    from lsml_b200.core.model import _ball_u0
    from lsml_b200.initializer import BallInitializer

    n = args.size
    seed = 1234                  # every rank evolves the SAME volume (replicas)
    img, radius, c = make_volume(n, seed)
    dx = np.ones(3)
    r_init = INIT_RADIUS * n / 256.0
    eng = BandEngine((n, n, n), 1, dx=dx, band=BAND, specs=SPECS, device=device)
    _DIAG["engine"] = eng
    eng.load_images(img[None], normalize=True)
    centre = 0.5 * np.array([n, n, n], dtype=np.float64)
    u0 = _ball_u0(eng, centre[None], [r_init]).clone()
    regs = train_regressors(eng, u0, n, radius, c, seed, log=log)
    model = LevelSetMachineLearning.from_regressors(make_features(), BallInitializer(radius=r_init),
                                                    regs, step=STEP, band=BAND)
    dev_regs = [model._device_model(i, device) for i in range(N_ITERS)]

    def one_pass():
        eng.set_level_sets(u0)
        for i in range(N_ITERS):
            eng.step(dev_regs[i], STEP)

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        one_pass()
    ctx.barrier()
    eng.units_dev.zero_()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            one_pass()
        ev1.record()
        ctx.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    bvi = eng.band_voxel_iterations
    failed_rebuilds = int(eng._fmm_status[0].item())

    # ---- end to end through the public API, host buffers --------------------------------------
    img_pinned = torch.from_numpy(img[None]).pin_memory()
    out_host = torch.empty((1, n, n, n), dtype=torch.bool).pin_memory()

    def e2e_pass():
        u_fin, m_fin = model.segment_batch(img_pinned, dx=dx, n_iters=N_ITERS, device=device)
        out_host.copy_(u_fin > 0, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_pass()
    ctx.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_pass()
    ctx.barrier()
    e2e_s = time.perf_counter() - t0
    seg_voxels = int(out_host.sum())

    # the reference-signature call: segment(img) -> all iterates as one host float64 array
    e2e_segment = None
    if rank == 0:
        try:
            t0 = time.perf_counter()
            us = model.segment(img, dx=dx, verbose=False, iterate_until_validation_max=False,
                               device=device)
            seg_s = time.perf_counter() - t0
            e2e_segment = {"value": (bvi / args.steps) / seg_s, "unit": UNIT,
                           "ms_per_step": 1e3 * seg_s, "h2d_bytes_per_step": int(img.nbytes),
                           "d2h_bytes_per_step": int(2 * 8 * (bvi / args.steps) + 8 * n ** 3),
                           "host_result_bytes": int(us.nbytes),
                           "api": "LevelSetMachineLearning.segment(host image) -> us "
                                  "(n_iters+1, n, n, n) float64 on the host, rebuilt from the "
                                  "per-iteration sparse records (model.py:283-411 signature)",
                           "same_final_mask_as_segment_batch":
                               bool(int((us[-1] > 0).sum()) == seg_voxels)}
            del us
        except Exception as exc:      # e.g. a host without 7 GB to spare
            e2e_segment = {"error": repr(exc)}

    # ---- per-phase evidence: one more pass with CUDA-event brackets around every phase ---------
    # (events on the launching stream; the brackets add a few microseconds per phase, which is
    # why this is a separate pass and not the timed region)
    eng.timers = {}
    eng.set_level_sets(u0)
    cand_sum, sweeps_sum, replay_sum = 0, 0, 0
    l_iter0 = _lib.launch_count()
    for i in range(N_ITERS):
        eng.step(dev_regs[i], STEP)
        if i % 5 == 4:                      # counters of a sample of the rebuilds (host sync)
            ctr = eng.fmm_counters()
            cand_sum += ctr["n_cand"]; sweeps_sum += ctr["sweeps"]; replay_sum += ctr["replays"]
    launches_per_iter = (_lib.launch_count() - l_iter0) / N_ITERS
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
    upd_n, upd_ms = phases.get("update", (1, 0.0))
    # the velocity kernel (features gathered in-kernel -> forest): dominant kernel of region (A),
    # i.e. of BASELINE.json's metric.  Algorithmic bytes per band voxel: band index
    # (8) + one float64 per image plane (8 P) + velocity out (8).
    vel_n, vel_ms = phases.get("predict", (1, 0.0))
    n_planes = int(eng.planes.shape[0]) if getattr(eng, "planes", None) is not None else 0
    velocity_bytes = (16.0 + 8.0 * n_planes) * (bvi / max(1, args.steps)) / N_ITERS
    velocity_gbs = velocity_bytes / (vel_ms / max(1, vel_n) * 1e-3) / 1e9 if vel_ms > 0 else 0.0
    update_bytes = 64.0 * (bvi / max(1, args.steps)) / N_ITERS      # ~64 B/band voxel
    update_gbs = update_bytes / (upd_ms / max(1, upd_n) * 1e-3) / 1e9 if upd_ms > 0 else 0.0

    # ---- the HBM-bound kernel named by north_star: dense stencil at the same grid ---------------
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
    grads = torch.zeros((3, n, n, n), dtype=torch.float64, device=device)

    def centered_call():
        _lib.check(_lib.lib.lsml_gradient_centered_f64(3, shape_c, _lib.c_i64(1), _lib.ptr(A), _lib.ptr(full),
                                                       _lib.ptr(grads), _lib.ptr(out), dx_c, 0,
                                                       _lib.current_stream()), "gradient_centered")
    for _ in range(3):
        centered_call()
    gc_ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        centered_call()
        b.record()
        torch.cuda.synchronize()
        gc_ms.append(a.elapsed_time(b))
    gc_ms = float(np.median(gc_ms))
    centered_bytes = 41.0 * n ** 3         # 8 (A) + 1 (mask) + 3 x 8 (gradient) + 8 (magnitude)
    centered_gbs = centered_bytes / (gc_ms * 1e-3) / 1e9
    del A, nu, full, out, flush, grads

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
    clocks_summary = clocks.summary()

    # ---- the sharded workloads, at this N -----------------------------------------------------
    del eng, model, dev_regs
    _DIAG.clear()
    torch.cuda.empty_cache()
    sections = {}
    wanted = [s for s in args.sections.split(",") if s and s != "none"]
    for name in wanted:
        try:
            if name == "c4":
                sections["c4"] = section_c4(ctx, total=args.items or C4_TOTAL, log=log)
            elif name == "c5":
                sections["c5"] = section_c5(ctx, n=args.c5_size, log=log)
        except Exception as exc:      # the headline line must survive a failing section
            import traceback
            traceback.print_exc(file=sys.stderr)
            sections[name] = {"error": repr(exc)[:300]}
        _DIAG.clear()
        torch.cuda.empty_cache()
        if log:
            log("section %s done" % name)

    if rank == 0:
        march_traffic, march_src = ncu_traffic("fmm_march_kernel")
        stencil_traffic, stencil_src = ncu_traffic("gmag_os_tma_kernel")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(n),
            "results": {"band_voxel_iterations_per_step": bvi_all // max(1, args.steps) // world,
                        "final_segmentation_voxels": seg_voxels,
                        "rebuilds_not_converged": "%d of %d (every rebuild of the timed region "
                                                  "and the warm-up)" % (
                                                      failed_rebuilds,
                                                      (args.steps + args.warmup) * (N_ITERS + 1)),
                        "launches_per_iteration": launches_per_iter,
                        "replicas": "every rank evolves the same volume" if world > 1 else None},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(img_pinned.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel()),
                    "api": "LevelSetMachineLearning.segment_batch(host image) -> host mask",
                    "ms_per_step": e2e_ms_max / e2e_steps},
            "e2e_segment": e2e_segment,
            "gpu_launches": launches_all,
            "clocks": clocks_summary,
            "roofline": {"bound": "hbm",
                         "kernel": "band rebuild: fmm_march_kernel<3> (+ clear/prepare/init/seed), "
                                   "the dominant kernel of the step",
                         "achieved": rebuild_gbs, "peak": peak, "unit": "GB/s",
                         "frac": rebuild_gbs / peak, "traffic": march_traffic,
                         "traffic_source": march_src,
                         "peak_source": peak_src, "ms": rebuild_avg_ms,
                         "share_of_step": rebuild_ms / max(phase_total, 1e-9),
                         "algorithmic_bytes_per_launch": rebuild_bytes,
                         "candidates_per_launch": avg_cand,
                         "sweeps_per_launch": sweeps_sum / max(1, n_sampled),
                         "replays_per_launch": replay_sum / max(1, n_sampled),
                         "note": "latency-bound (dependency depth of the fast-marching order x "
                                 "replay latency), not HBM-bound: DESIGN.md 4.6"},
            "roofline_stencil": {"bound": "hbm",
                                 "kernel": "dense gmag_os: gmag_os_tma_kernel (TMA plane pipeline; %d^3 "
                                           "float64, full mask, L2 flushed between launches)" % n,
                                 "achieved": stencil_gbs, "peak": peak, "unit": "GB/s",
                                 "frac": stencil_gbs / peak, "traffic": stencil_traffic,
                                 "traffic_source": stencil_src,
                                 "algorithmic_bytes_per_launch": stencil_bytes,
                                 "peak_source": peak_src, "ms": st_ms},
            "roofline_centered": {"bound": "hbm",
                                  "kernel": "dense gradient_centered (%d^3 float64, full mask, L2 "
                                            "flushed between launches): the same TMA plane pipeline" % n,
                                  "achieved": centered_gbs, "peak": peak, "unit": "GB/s",
                                  "frac": centered_gbs / peak, "traffic": None,
                                  "algorithmic_bytes_per_launch": centered_bytes,
                                  "peak_source": peak_src, "ms": gc_ms},
            "roofline_velocity": {"bound": "hbm",
                                  "kernel": "forest_ranked_kernel: feature gather + RF-100 walked from "
                                            "shared memory (rank-quantised 4-byte nodes), the dominant "
                                            "kernel of region (A)",
                                  "achieved": velocity_gbs, "peak": peak, "unit": "GB/s",
                                  "frac": velocity_gbs / peak,
                                  "traffic": ncu_traffic("forest_ranked_kernel")[0],
                                  "traffic_source": ncu_traffic("forest_ranked_kernel")[1],
                                  "algorithmic_bytes_per_launch": velocity_bytes,
                                  "peak_source": peak_src, "ms": vel_ms / max(1, vel_n),
                                  "note": "not HBM-bound: ~1100 dependent shared-memory node visits "
                                          "per voxel; the limiter is the L1/shared data pipe "
                                          "(bank-conflict wavefronts of divergent node fetches) and "
                                          "instruction issue, DESIGN.md 4.4"},
            "roofline_update": {"bound": "hbm",
                                "kernel": "band_update_kernel + band_apply_kernel (banded upwind "
                                          "|Du|, u += step*v*|Du|, exact statistics)",
                                "achieved": update_gbs, "peak": peak, "unit": "GB/s",
                                "frac": update_gbs / peak, "traffic": None,
                                "algorithmic_bytes_per_launch": update_bytes,
                                "ms": upd_ms / max(1, upd_n),
                                "note": "two launches over ~0.27 M band voxels: launch-latency "
                                        "bound at this band size; see c4.roofline_update for the "
                                        "same kernels on a 2048-crop batch"},
            "phases_ms_per_step": {k: v[1] for k, v in phases.items()},
        }
        line.update(sections)
        # SURVEY.md 8(d) region (A): features + velocity + update with the band given -- the part
        # BASELINE.json's metric names; `value` above is region (B), the full step with band
        # rebuild and compaction.  From the per-phase device times of the evidence pass (one
        # step on rank 0's volume).
        ms_a = sum(phases.get(k, (0, 0.0))[1] for k in ("features", "predict", "update"))
        if ms_a > 0:
            line["region_a"] = {"value": (bvi_all / world / max(1, args.steps)) / (ms_a * 1e-3),
                                "unit": UNIT, "ms_per_step": ms_a,
                                "what": "features + velocity + update, band given, one GPU"}
        if not args.no_cpu_baseline:
            res = cpu_reference_run(n, 1234, 1, 0, budget_s=25.0, log=log, literal_gap=True)
            v = res["band_voxel_iterations"] / res["seconds"]
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": res["kind"],
                "sample": "1 evolution iteration (features+RF-100+update+band rebuild) of the same "
                          "%d^3 volume, single thread, %.1f s" % (n, res["seconds"]),
                "gradient_kernel": res["gradient_kernel"]}
            line["results"]["band_mismatch_vs_upstream_literal"] = res["gap"]
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
