"""Times the one-time image planes (normalisation + separable Gaussians + edge magnitudes) of the bench's
feature list on one volume.  LSML_GAUSS_NO_MARCH=1 selects the tap-gather kernel for every axis.
Usage: python tools/planes_probe.py [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lsml_b200.core.engine import BandEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
img, radius, c = bench.make_volume(n, 1234)
eng = BandEngine((n, n, n), 1, dx=np.ones(3), band=3.0, specs=bench.SPECS, device=dev)
imgs = torch.from_numpy(img[None]).to(dev)
ts = []
for rep in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); eng.load_images(imgs, normalize=True); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print("planes %s: %d^3, %d planes: %.1f us (median of the last 5), checksum %.17g"
      % ("tap-gather" if os.environ.get("LSML_GAUSS_NO_MARCH") else "march", n, eng.planes.shape[0],
         float(np.median(ts[1:])), float(eng.planes.double().sum().item())))
