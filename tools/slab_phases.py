"""Wall-clock breakdown of one slab-decomposed evolution iteration (synchronising after every
phase; diagnostic only):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_phases.py [n]"""
import os, sys, time, collections
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.linear_model import LinearRegression
from lsml_b200.core import slab as S
from lsml_b200.core.regressors import export_regressor

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
specs = [("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
         ("band_mean", 2.0), ("band_std", 2.0), ("dist_to_com",), ("moments", (0, 1, 2), (1, 2)), ("size",)]
vol = S.SlabVolume((n, n, n), specs=specs, device=dev, collective=S.TorchCollective())
sl = vol.ranks[0].local_slice()
ax = torch.arange(n, dtype=torch.float64, device=dev) - n / 2.0
a0 = ax[sl]
r2 = a0[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
gen = torch.Generator(device=dev); gen.manual_seed(5 + sl.start)
part = (r2 < (0.33 * n) ** 2).to(torch.float64) + 0.25 * torch.randn(r2.shape, dtype=torch.float64, device=dev, generator=gen)
u0 = torch.where(r2 < (0.2 * n) ** 2, 1.0, -1.0)
vol.load_image_parts([part], normalize=True)
X = np.random.RandomState(0).randn(64, 14); w = np.zeros(14); w[1] = 0.8
reg = export_regressor(LinearRegression().fit(X, X @ w + 0.15), dev)
vol.set_level_sets_parts([u0])
acc = collections.OrderedDict()
def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return out
    return wrapper
vol._global_stats = timed("global_stats(allreduce)", vol._global_stats)
vol._exchange = timed("exchange", vol._exchange)
vol.coll.allreduce_sum = timed("allreduce_sum", vol.coll.allreduce_sum)
for r in vol.ranks:
    for nm in ("rebuild_begin", "rebuild_march", "rebuild_requeue", "rebuild_finish", "compact_owned",
               "band_sums", "assemble", "predict_and_update"):
        setattr(r, nm, timed(nm, getattr(r, nm)))
for it in range(3):
    vol.step(reg, 0.4)
acc.clear()
iters = 10
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for it in range(iters):
    vol.step(reg, 0.4)
torch.cuda.synchronize(); total = time.perf_counter() - t0
if rank == 0:
    print("%d^3 over %d ranks: %.2f ms/iteration with per-phase synchronisation; rounds %d" % (n, world, 1e3 * total / iters, vol.rebuild_rounds))
    for k, v in acc.items():
        print("  %-28s %8.3f ms/iteration" % (k, 1e3 * v / iters))
    print("  %-28s %8.3f ms/iteration" % ("(sum)", 1e3 * sum(acc.values()) / iters))
dist.destroy_process_group()
