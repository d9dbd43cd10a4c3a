"""One volume slab-decomposed over the GPUs of a torchrun job (NCCL): correctness against the
single-GPU engine (rank 0, small volume) and per-iteration device time at a larger size.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_probe.py [n] [iters]"""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.linear_model import LinearRegression
from lsml_b200.core.engine import BandEngine
from lsml_b200.core.slab import SlabVolume, TorchCollective

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def volume(shape, seed):
    rs = np.random.RandomState(seed)
    ax = [np.arange(s, dtype=np.float64) - s / 2.0 for s in shape]
    r = np.sqrt(ax[0][:, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[2][None, None, :] ** 2)
    img = (r < 0.33 * min(shape)).astype(np.float64) + 0.25 * rs.standard_normal(shape)
    u0 = np.where(r < 0.2 * min(shape), 1.0, -1.0)
    return img, u0

specs = [("image_sample", 0.0), ("image_sample", 2.0), ("image_edge", 0.0), ("image_edge", 2.0),
         ("band_mean", 2.0), ("band_std", 2.0), ("dist_to_com",), ("moments", (0, 1, 2), (1, 2)), ("size",)]
rs = np.random.RandomState(0)
X = rs.randn(64, 14); w = np.zeros(14); w[1] = 0.8
reg = LinearRegression().fit(X, X @ w + 0.15)           # grow where the smoothed image is bright

# ---- correctness at a small size: every rank's owned planes vs rank-local single-GPU run
small = (72, 40, 44)
img, u0 = volume(small, 1)
slab = SlabVolume(small, specs=specs, device=dev, collective=TorchCollective())
slab.load_image(img, normalize=True); slab.set_level_sets(u0)
one = BandEngine(small, 1, specs=specs, device=dev)
one.load_images(img[None], normalize=True); one.set_level_sets(u0[None])
ok, rounds = True, 0
for it in range(5):
    one.step(reg, 0.4); slab.step(reg, 0.4); rounds = max(rounds, slab.rebuild_rounds)
    r = slab.ranks[0]
    ok &= bool(np.array_equal(slab.gather("mask"), one.mask[0, r.A:r.B].cpu().numpy()))
    ok &= bool(np.allclose(slab.gather("u"), one.u[0, r.A:r.B].cpu().numpy(), rtol=1e-10, atol=1e-10))
flag = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(flag)
if rank == 0:
    print("slab vs single GPU on %s over %d ranks: %s (max exchange rounds per rebuild %d)"
          % (small, world, "IDENTICAL masks, u to 1e-10" if int(flag) == 0 else "MISMATCH", rounds))
del slab, one

# ---- timing at n^3
shape = (n, n, n)
img, u0 = volume(shape, 2)
slab = SlabVolume(shape, specs=specs, device=dev, collective=TorchCollective())
slab.load_image(img, normalize=True)
for rep in range(2):
    slab.set_level_sets(u0)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tot = 0; rr = 0
    for it in range(iters):
        tot += slab.step(reg, 0.4); rr += slab.rebuild_rounds
    b.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("rep %d: %d^3 over %d GPUs: %.2f ms/iteration, %.1f M band-voxel-iterations/s, %.1f exchange rounds per rebuild"
              % (rep, n, world, float(t) / iters, tot / float(t) / 1e3, rr / iters))
dist.destroy_process_group()
