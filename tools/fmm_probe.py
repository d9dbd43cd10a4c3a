import sys, time, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.model import _ball_u0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
img, radius, c = bench.make_volume(n, 1234)
eng = BandEngine((n,n,n), 1, dx=np.ones(3), band=3.0, specs=bench.SPECS)
eng.load_images(img[None])
u0 = _ball_u0(eng, 0.5*np.array([[n,n,n]], dtype=float), [40.0*n/256]).clone()
eng.set_level_sets(u0)
print('init band', eng.n_band, eng.fmm_counters())
# perturb u on the band like an evolution step and time rebuilds
g = torch.Generator(device='cuda').manual_seed(0)
for it in range(5):
    idx = eng.band_idx[:eng.n_band]
    eng.u.view(-1)[idx] += 0.3*torch.randn(idx.numel(), device='cuda', dtype=torch.float64, generator=g)
    eng._global_stats()
    torch.cuda.synchronize(); t=time.perf_counter()
    eng._rebuild_band()
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(it, 'band', eng.n_band, 'rebuild+compact %.3f ms' % (dt*1e3), eng.fmm_counters())
