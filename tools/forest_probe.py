"""Times the velocity kernel (feature program -> forest) on the bench's C3 state and checks it
against scikit-learn.  LSML_FOREST_VARIANT = global | staged2 | staged4 selects the kernel
(regress.cu); run once per variant:  python tools/forest_probe.py [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.model import _ball_u0
from lsml_b200.core.regressors import export_regressor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
img, radius, c = bench.make_volume(n, 1234)
eng = BandEngine((n, n, n), 1, dx=np.ones(3), band=3.0, specs=bench.SPECS, device=dev)
eng.load_images(img[None], normalize=True)
u0 = _ball_u0(eng, 0.5 * np.array([[n, n, n]], dtype=np.float64), [60.0 * n / 256]).clone()
eng.set_level_sets(u0)
nb = eng.n_band
X = eng.compute_features().cpu().numpy()
idx = eng.band_idx[:nb].cpu().numpy()
y = bench.target_distance_on_band(idx, n, radius, c)
host = bench._fit_forest(X, y, 7)
reg = export_regressor(host, dev)
eng._ensure_band_buffers()
eng._item_level_features()
def run():
    reg.predict_band(eng.band_source(), eng.n_band_dev, eng.vel)
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
got = eng.vel[:nb].cpu().numpy()
want = host.predict(X)
if os.environ.get("LSML_FOREST_DEPTHS"):
    # how deep the walks are: mean leaf depth, and the mean over warps (32 consecutive band
    # voxels) of the deepest lane -- what a lock-step warp pays per tree
    forest = host.steps[-1][1]
    Xs = host.steps[0][1].transform(X).astype(np.float32)
    tot = mx = 0.0
    for est in forest.estimators_[:20]:
        t = est.tree_
        depth = np.zeros(t.node_count, dtype=np.int64)
        for i in range(t.node_count):
            if t.children_left[i] != -1:
                depth[t.children_left[i]] = depth[t.children_right[i]] = depth[i] + 1
        d = depth[t.apply(Xs)]
        tot += d.mean()
        mx += d[:len(d) // 32 * 32].reshape(-1, 32).max(axis=1).mean()
    print("leaf depth: mean %.2f, mean of the per-warp maximum %.2f (20 trees)" % (tot / 20, mx / 20))
print("variant %-8s band %d trees %d max_nodes %d: %.1f us (median of 10), equal to sklearn: %s"
      % (os.environ.get("LSML_FOREST_VARIANT", "default") + os.environ.get("LSML_FOREST_RK", ""), nb, reg.n_trees, reg.max_nodes,
         float(np.median(ts)), bool(np.array_equal(got, want))))
