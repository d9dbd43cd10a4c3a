"""Batched workloads (many independent items in one engine): per-iteration device time.
  python tools/batch_probe.py 2d   -> 1024 x 101^2 (C2-like)
  python tools/batch_probe.py 3d [items] -> items x 64^3 (C4-like)"""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.ensemble import RandomForestRegressor
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import StandardScaler
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.regressors import export_regressor

kind = sys.argv[1] if len(sys.argv) > 1 else "2d"
rs = np.random.RandomState(0)
dev = torch.device("cuda")
if kind == "2d":
    B, shape, r0 = 1024, (101, 101), 10.0
    specs = [("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
             ("band_mean", 0.0), ("band_std", 2.0), ("dist_to_com",), ("moments", (0, 1), (1, 2)),
             ("size",), ("boundary_size",), ("isoperimetric",)]
else:
    B, shape, r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 512, (64, 64, 64), 5.0
    specs = [("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
             ("band_mean", 0.0), ("band_std", 2.0), ("dist_to_com",), ("moments", (0, 1, 2), (1, 2)),
             ("size",)]
nd = len(shape)
grid = np.indices(shape).astype(np.float64)
imgs = np.empty((B,) + shape)
u0 = np.empty((B,) + shape)
for b in range(B):
    c = np.array(shape) / 2.0 + rs.uniform(-8, 8, size=nd)
    rad = rs.uniform(0.1, 0.28) * shape[0]
    r = np.sqrt(sum((grid[a] - c[a]) ** 2 for a in range(nd)))
    imgs[b] = (r < rad).astype(np.float64) + 0.2 * rs.standard_normal(shape)
    rc = np.sqrt(sum((grid[a] - shape[a] / 2.0) ** 2 for a in range(nd)))
    u0[b] = np.where(rc < r0, 1.0, -1.0)
eng = BandEngine(shape, B, dx=np.ones(nd), band=3.0, specs=specs, device=dev)
eng.load_images(imgs, normalize=True)
eng.set_level_sets(u0)
X = eng.compute_features().cpu().numpy()
y = X[:, 1] + 0.2 * rs.standard_normal(X.shape[0])           # grow where the smoothed image is bright
sub = rs.choice(X.shape[0], size=min(X.shape[0], 20000), replace=False)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    reg = Pipeline([("s", StandardScaler()), ("r", RandomForestRegressor(n_estimators=100, max_samples=300, random_state=0))]).fit(X[sub], y[sub])
dreg = export_regressor(reg, dev)
for rep in range(2):
    eng.set_level_sets(u0)
    eng.timers = {} if rep == 1 else None
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tot = 0
    iters = 10
    for it in range(iters):
        tot += eng.n_band; eng.step(dreg, 0.4)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("rep %d: %s x %s, %d iterations: %.2f ms/iteration, %d band voxel iterations, %.1f M bvi/s"
          % (rep, B, shape, iters, ms / iters, tot, tot / ms / 1e3))
print({k: round(v[1] / 10, 3) for k, v in eng.phase_times_ms().items()}, "ms per iteration by phase")
print("fmm:", eng.fmm_counters())
