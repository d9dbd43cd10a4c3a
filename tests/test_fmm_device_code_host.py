"""CPU-only: the DEVICE code of the band rebuild, compiled for the host.

The marked regions of lsml_b200/csrc/fmm.cu (``[host-test:begin]`` .. ``[host-test:end]``:
``init_distance``, ``solve_quadratic``, ``replay`` and what they use) are extracted verbatim,
compiled with g++ behind a small shim for the CUDA qualifiers/intrinsics (tests/native/
cuda_host_shim.h) and iterated to their fixed point over whole grids.  The result must equal the
oracle's sequential marcher bit for bit.  oracle/fmm_replay_model.c checks the ALGORITHM with a
separate C restatement; this test checks the kernel's own source text, where no GPU exists.

Both contributor rules of DESIGN.md 4.6 are covered: the adopted one and the legacy one
(-DLSML_FMM_RULE_LEGACY / LSML_ORACLE_RULE_R2=0, run in a subprocess because the oracle reads
the switch once).
"""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build_host_lib(out_dir, legacy):
    src = open(os.path.join(ROOT, "lsml_b200", "csrc", "fmm.cu")).read()
    regions = re.findall(r"// \[host-test:begin\][^\n]*\n(.*?)// \[host-test:end\]", src, re.S)
    assert len(regions) == 4, "fmm.cu: expected 4 host-testable regions"
    text = ('#include "cuda_host_shim.h"\nnamespace lsml {\n' + "\n".join(regions) +
            open(os.path.join(HERE, "native", "fmm_host_driver.inc")).read())
    cpp = os.path.join(out_dir, "fmm_device_host%s.cpp" % ("_legacy" if legacy else ""))
    so = cpp[:-4] + ".so"
    with open(cpp, "w") as f:
        f.write(text)
    cmd = ["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC",
           "-I", os.path.join(HERE, "native"), cpp, "-o", so]
    if legacy:
        cmd.insert(1, "-DLSML_FMM_RULE_LEGACY")
    subprocess.check_call(cmd)
    return so


def device_fixed_point(lib, u, dx, band, jacobi, max_sweeps=500):
    u = np.ascontiguousarray(u, dtype=np.float64)
    dx = np.asarray(dx, dtype=np.float64).copy()
    shape = np.asarray(u.shape, dtype=np.int32)
    dist = np.zeros_like(u)
    frozen = np.zeros(u.shape, dtype=np.uint8)
    p = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
    lib.fmm_device_fixed_point.restype = ctypes.c_long
    n = lib.fmm_device_fixed_point(
        ctypes.c_int(u.ndim), p(shape, ctypes.c_int), p(u, ctypes.c_double),
        p(dx, ctypes.c_double), ctypes.c_double(float(band)), ctypes.c_int(jacobi),
        ctypes.c_int(max_sweeps), p(dist, ctypes.c_double), p(frozen, ctypes.c_uint8))
    return dist, frozen.astype(bool), n


def campaign(so_path, n_cases, seed):
    """Random level sets (the generator of tests/test_fmm_algorithm.py) and evolved step
    functions: device code fixed point == oracle marcher.  Returns (checked, mismatches)."""
    from oracle import oracle as O
    from test_fmm_algorithm import _random_case
    O.build()
    lib = ctypes.CDLL(so_path)
    rs = np.random.RandomState(seed)
    checked, bad = 0, []
    for case in range(n_cases):
        u, dx, band = _random_case(rs)
        if not ((u > 0).any() and (u < 0).any()):
            continue
        wd, wm = O.distance_transform(u, band, dx)
        for jacobi in (0, 1):
            d, m, sweeps = device_fixed_point(lib, u, dx, band, jacobi)
            if sweeps <= 0 or not np.array_equal(m, wm) or not np.array_equal(d, wd):
                bad.append((case, jacobi, int(sweeps)))
        checked += 1
    # the contributor-rule fixtures: exact under the adopted rule, differ under the legacy one
    limit = []
    for name in ("fmm_nonmonotone_case.npz", "fmm_c4_evolved_case.npz"):
        z = np.load(os.path.join(HERE, "golden", name))
        wd, wm = O.distance_transform(z["u"], float(z["band"]), z["dx"])
        d, m, sweeps = device_fixed_point(lib, z["u"], z["dx"], float(z["band"]), 0)
        limit.append((int(sweeps), bool(np.array_equal(m, wm)), int((d != wd).sum())))
    return checked, bad, limit


@pytest.mark.parametrize("legacy", [False, True])
def test_device_replay_source_equals_sequential_marcher(tmp_path, legacy):
    so = build_host_lib(str(tmp_path), legacy)
    env = dict(os.environ)
    env["LSML_ORACLE_RULE_R2"] = "0" if legacy else "1"
    env["PYTHONPATH"] = os.pathsep.join([ROOT, HERE, env.get("PYTHONPATH", "")])
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "from test_fmm_device_code_host import campaign\n"
            "print(repr(campaign(%r, 150, 7)))\n" % (ROOT, HERE, so))
    out = subprocess.check_output([sys.executable, "-c", code], env=env, cwd=ROOT, text=True)
    checked, bad, limit = eval(out.strip().splitlines()[-1])
    assert checked > 100
    assert bad == [], bad
    for sweeps, mask_equal, n_diff in limit:
        assert sweeps > 0 and mask_equal
        assert (0 < n_diff < 100) if legacy else (n_diff == 0)
