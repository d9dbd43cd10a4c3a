"""A few evolution iterations of the bench workload (256^3, F=16, RF-100) for profiling:
  python tools/iter_probe.py [n] [iters]                      # event-timed phases
  ncu --metrics gpu__time_duration.sum ... python tools/iter_probe.py   # launch list"""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.model import _ball_u0
from lsml_b200.core.regressors import export_regressor
from sklearn.ensemble import RandomForestRegressor
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import StandardScaler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda")
img, radius, c = bench.make_volume(n, 1234)
eng = BandEngine((n, n, n), 1, dx=np.ones(3), band=bench.BAND, specs=bench.SPECS, device=dev)
eng.load_images(img[None], normalize=True)
u0 = _ball_u0(eng, 0.5 * np.array([[n, n, n]], dtype=np.float64), [bench.INIT_RADIUS * n / 256]).clone()
eng.set_level_sets(u0)
X = eng.compute_features().cpu().numpy()
idx = eng.band_idx[:eng.n_band].cpu().numpy()
y = bench.target_distance_on_band(idx, n, radius, c)
rs = np.random.RandomState(0)
sub = rs.choice(eng.n_band, size=min(eng.n_band, 20000), replace=False)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    reg = Pipeline([("s", StandardScaler()), ("r", RandomForestRegressor(n_estimators=100, max_samples=300, random_state=rs))]).fit(X[sub], y[sub])
dreg = export_regressor(reg, dev)
for rep in range(2):
    eng.set_level_sets(u0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tot = 0
    for it in range(iters):
        tot += eng.n_band; eng.step(dreg, bench.STEP)
    b.record(); torch.cuda.synchronize()
    print("rep %d: %d iterations, %d band voxel iterations, %.1f us/iteration (device), %.1f us/iteration (host wall)"
          % (rep, iters, tot, a.elapsed_time(b) * 1e3 / iters, (time.perf_counter() - t0) * 1e6 / iters))
