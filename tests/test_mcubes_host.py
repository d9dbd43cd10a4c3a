"""CPU-only: the marching-cubes cube routine of the PRODUCT (lsml_b200/csrc/mcubes.cu, marked
region, compiled for the host with g++) against the oracle's independently written restatement of
the same specification (oracle/lsml_oracle.c: orc_marching_cubes_area), and the reference's own
known answer (lsml/feature/provided/test/test_shape.py:59-74: a sphere's area is 4 pi to 0
places)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "lsml_b200", "csrc")

DRIVER = r'''
}  // namespace lsml
extern "C" double mcubes_host(int n0, int n1, int n2, const double* u, const double* dx) {
    const double d[3] = {dx[0], dx[1], dx[2]};
    double total = 0.0;
    for (long i = 0; i + 1 < n0; ++i)
        for (long j = 0; j + 1 < n1; ++j)
            for (long k = 0; k + 1 < n2; ++k) {
                const double* q = u + (i * n1 + j) * n2 + k;
                const long s0 = (long)n1 * n2, s1 = n2;
                const double v[8] = {q[0], q[1], q[s1], q[s1 + 1], q[s0], q[s0 + 1], q[s0 + s1],
                                     q[s0 + s1 + 1]};
                total += lsml::mc_cube_area(v, d);
            }
    return total;
}
'''


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    src = open(os.path.join(CSRC, "mcubes.cu")).read()
    regions = re.findall(r"// \[host-test:begin\][^\n]*\n(.*?)// \[host-test:end\]", src, re.S)
    assert len(regions) == 1
    out = tmp_path_factory.mktemp("mcubes_host")
    cpp, so = str(out / "m.cpp"), str(out / "m.so")
    with open(cpp, "w") as f:
        f.write('#include "cuda_host_shim.h"\nnamespace lsml {\n' + regions[0] + DRIVER)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC",
                           "-I", os.path.join(HERE, "native"), cpp, "-o", so])
    lib = ctypes.CDLL(so)
    lib.mcubes_host.restype = ctypes.c_double
    return lib


def product_area(lib, u, dx):
    u = np.ascontiguousarray(u, dtype=np.float64)
    dx = np.asarray(dx, dtype=np.float64)
    P = ctypes.POINTER(ctypes.c_double)
    return lib.mcubes_host(*[ctypes.c_int(s) for s in u.shape], u.ctypes.data_as(P),
                           dx.ctypes.data_as(P))


def test_cube_routine_matches_the_oracle(host_lib, oracle):
    rs = np.random.RandomState(0)
    g = np.indices((18, 21, 16)).astype(float)
    fields = [rs.randn(12, 13, 11),                                   # every ambiguous case
              rs.randn(9, 9, 9) ** 3,
              6.3 - np.sqrt(((g - 8.2) ** 2).sum(0)) + 0.4 * rs.randn(18, 21, 16),
              np.where(rs.rand(10, 10, 10) > 0.5, 1.0, -1.0),          # +-1 step function
              g[0] - 4.0]                                              # exact zeros on a plane
    for u in fields:
        for dx in ([1.0, 1.0, 1.0], [0.7, 1.3, 2.0]):
            want = oracle.boundary_size(u, dx)
            got = product_area(host_lib, u, dx)
            assert want > 0
            assert abs(got - want) <= 1e-12 * want, (got, want)


def test_sphere_known_answer(host_lib):
    x, dx = np.linspace(-2, 2, 126, retstep=True)
    y, dy = np.linspace(-2, 2, 101, retstep=True)
    z, dz = np.linspace(-2, 2, 151, retstep=True)
    xx, yy, zz = np.meshgrid(x, y, z, indexing="ij")
    w = 1 - np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)
    area = product_area(host_lib, w, [dx, dy, dz])
    assert abs(area - 4 * np.pi) < 0.02          # the reference asks for 0 places (< 0.5)
    # watertight: the patch areas of u and -u agree except where u == 0 exactly (none here)
    assert abs(product_area(host_lib, -w, [dx, dy, dz]) - area) < 1e-9
