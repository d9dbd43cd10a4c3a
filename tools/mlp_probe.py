"""Times the MLP velocity kernel (LSML_MLP_VARIANT=scalar: the one-thread-per-voxel kernel; default: FP64
tensor cores) on a bench-sized band and checks it against scikit-learn."""
import os, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from sklearn.neural_network import MLPRegressor
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import StandardScaler
from lsml_b200.core.regressors import export_regressor
rs = np.random.RandomState(0)
X = rs.randn(4000, 16); y = np.sin(X[:, 0]) + X[:, 3]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = Pipeline([("s", StandardScaler()), ("r", MLPRegressor(hidden_layer_sizes=(100,), max_iter=30, random_state=0))]).fit(X, y)
reg = export_regressor(m)
Xt = torch.from_numpy(rs.randn(285188, 16)).cuda()
nb = torch.tensor([Xt.shape[0]], dtype=torch.int64, device="cuda"); out = torch.empty(Xt.shape[0], dtype=torch.float64, device="cuda")
for _ in range(3): reg.predict_into(Xt, nb, out)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); reg.predict_into(Xt, nb, out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
want = m.predict(Xt.cpu().numpy())
print("mlp %s: 285188 x 16 -> 100 -> 1: %.1f us, max rel err %.2e" % (os.environ.get("LSML_MLP_VARIANT", "dmma"), float(np.median(ts)), float(np.abs(out.cpu().numpy() - want).max() / np.abs(want).max())))
