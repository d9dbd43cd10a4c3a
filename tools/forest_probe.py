"""Times forest inference alone on a synthetic band feature matrix (B x 16, RF-100)."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.ensemble import RandomForestRegressor
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import StandardScaler
from lsml_b200.core.regressors import export_regressor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 270000
rs = np.random.RandomState(0)
X = rs.randn(B, 16); X[:, 8:] = np.cumsum(rs.randn(B, 8) * 0.01, axis=0)   # spatially coherent columns
y = X[:, 0] + np.sin(3 * X[:, 9]) + 0.3 * rs.randn(B)
sub = rs.choice(B, 20000, replace=False)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    reg = Pipeline([("s", StandardScaler()), ("r", RandomForestRegressor(n_estimators=100, max_samples=300, random_state=0))]).fit(X[sub], y[sub])
dev = torch.device("cuda")
dreg = export_regressor(reg, dev)
Xd = torch.from_numpy(X).to(dev)
out = dreg.predict(Xd)
want = reg.predict(X)
print("max abs diff vs sklearn:", np.abs(out.cpu().numpy() - want).max(), "nodes", dreg.n_nodes)
nb = torch.tensor([B], dtype=torch.int64, device=dev); o = torch.empty(B, dtype=torch.float64, device=dev)
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dreg.predict_into(Xd, nb, o); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print("forest predict %d x 16, 100 trees: %.1f us (median)" % (B, np.median(ts)))
