"""Times the band rebuild (fast marching) and the other per-iteration phases on an evolving
volume.  Usage: python tools/fmm_probe.py [n] [iters]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.linear_model import LinearRegression
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.regressors import export_regressor
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda")
img, radius, c = bench.make_volume(n, 1234)
specs = [("image_sample", 2.0)]
eng = BandEngine((n, n, n), 1, dx=np.ones(3), band=3.0, specs=specs, device=dev)
eng.load_images(img[None], normalize=True)
from lsml_b200.core.model import _ball_u0
u0 = _ball_u0(eng, 0.5 * np.array([[n, n, n]], dtype=np.float64), [40.0 * n / 256]).clone()
reg = LinearRegression().fit(np.array([[0.0], [1.0]]), np.array([0.0, 1.0]))   # v = smoothed image
dreg = export_regressor(reg, dev)

def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3

for rep in range(2):
    eng.set_level_sets(u0)
    for it in range(iters):
        nb = eng.n_band
        t_step = timed(lambda: eng.step(dreg, 0.35, rebuild=False)); eng._n_band_host = None
        eng._band_is_rebuilt = True
        t_fmm = timed(lambda: eng._rebuild_band(from_band=True))
        ctr = eng.fmm_counters()
        if rep == 1:
            print("it %2d band %7d  step(no rebuild) %8.1f us  rebuild+compact %8.1f us  init %d cand %d sweeps %d (tail %d) replays %d conv %s"
                  % (it, nb, t_step, t_fmm, ctr["n_init"], ctr["n_cand"], ctr["sweeps"], ctr["tail_sweeps"], ctr["replays"], ctr["converged"]))
ss = eng.fmm_sweep_sizes()
print("sweep sizes:", ss[:32])
print("sweep us   :", [round((ss[32+i+1]-ss[32+i])/1e3,1) for i in range(31) if ss[32+i+1] > 0])
from lsml_b200 import _lib
_lib.lib.lsml_fmm_phase_offset.restype = _lib.c_i64
_lib.lib.lsml_fmm_counters_offset.restype = _lib.c_i64
off = int(_lib.lib.lsml_fmm_counters_offset(_lib.c_i64(eng.total))) + int(_lib.lib.lsml_fmm_phase_offset())
ph = eng._ws_fmm.view(torch.uint8)[off:off + 64].view(torch.int64).cpu().numpy()
if ph[7] > 0:   # -DFMM_PHASE_CLOCKS build: cycles per phase of a single-lane grid-wide sweep
    names = ["top->entry", "replay", "record", "publish", "fence", "barrier"]
    print("phase cycles per sampled sweep (%d sweeps):" % ph[7],
          {n: round(float(ph[i]) / ph[7]) for i, n in enumerate(names)})
