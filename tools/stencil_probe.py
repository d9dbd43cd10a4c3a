"""Times the dense stencil kernels alone (pre-allocated outputs, L2 flushed between launches).
Usage: python tools/stencil_probe.py [n] ; under ncu: ncu -k regex:gmag_os ... python tools/stencil_probe.py"""
import ctypes, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsml_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")
A = torch.randn((n, n, n), dtype=torch.float64, device=dev)
nu = torch.randn_like(A)
mask = torch.ones((n, n, n), dtype=torch.uint8, device=dev)
out = torch.zeros_like(A)
grads = torch.zeros((3, n, n, n), dtype=torch.float64, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
shape = _lib.int_array((n, n, n))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6530.3

def run(name, fn, bytes_per_voxel):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    ms = float(np.median(ts)); gbs = bytes_per_voxel * n ** 3 / (ms * 1e-3) / 1e9
    print("%-34s %8.1f us  %7.1f GB/s  %5.1f%% of %.0f" % (name, ms * 1e3, gbs, 100 * gbs / peak, peak))

for dx in ((1.0, 1.0, 1.0), (0.7, 1.3, 1.1)):
    dxc = _lib.double_array(dx)
    tag = "dx=1" if dx[0] == 1.0 else "dx generic"
    _lib.lib.lsml_stencil_use_tma(0)
    run("gmag_os f64 %s (register march)" % tag, lambda: _lib.check(_lib.lib.lsml_gmag_os_f64(3, shape, _lib.c_i64(1), _lib.ptr(A), _lib.ptr(mask), _lib.ptr(nu), _lib.ptr(out), dxc, _lib.current_stream())), 25)
    ref_out = out.clone()
    _lib.lib.lsml_stencil_use_tma(1)
    out.zero_()
    run("gmag_os f64 %s (TMA, NULL mask)" % tag, lambda: _lib.check(_lib.lib.lsml_gmag_os_f64(3, shape, _lib.c_i64(1), _lib.ptr(A), _lib.c_p(0), _lib.ptr(nu), _lib.ptr(out), dxc, _lib.current_stream())), 24)
    print("    TMA == register march: %s" % bool(torch.equal(out, ref_out)))
    run("gmag_os f64 %s" % tag, lambda: _lib.check(_lib.lib.lsml_gmag_os_f64(3, shape, _lib.c_i64(1), _lib.ptr(A), _lib.ptr(mask), _lib.ptr(nu), _lib.ptr(out), dxc, _lib.current_stream())), 25)
    run("gradient_centered f64 %s" % tag, lambda: _lib.check(_lib.lib.lsml_gradient_centered_f64(3, shape, _lib.c_i64(1), _lib.ptr(A), _lib.ptr(mask), _lib.ptr(grads), _lib.ptr(out), dxc, 0, _lib.current_stream())), 8 + 1 + 24 + 8)
A32, nu32, out32 = A.float(), nu.float(), out.float()
dxc = _lib.double_array((1.0, 1.0, 1.0))
_lib.lib.lsml_stencil_use_tma(0)
run("gmag_os f32 dx=1 (register march)", lambda: _lib.check(_lib.lib.lsml_gmag_os_f32(3, shape, _lib.c_i64(1), _lib.ptr(A32), _lib.ptr(mask), _lib.ptr(nu32), _lib.ptr(out32), dxc, _lib.current_stream())), 13)
ref32 = out32.clone()
_lib.lib.lsml_stencil_use_tma(1)
out32.zero_()
run("gmag_os f32 dx=1", lambda: _lib.check(_lib.lib.lsml_gmag_os_f32(3, shape, _lib.c_i64(1), _lib.ptr(A32), _lib.ptr(mask), _lib.ptr(nu32), _lib.ptr(out32), dxc, _lib.current_stream())), 13)
print("    f32 TMA == register march: %s" % bool(torch.equal(out32, ref32)))
b2 = torch.empty_like(A)
run("torch copy (reference)", lambda: b2.copy_(A), 16)
