"""CPU-only: the band rebuild's SCHEDULER source on one host thread.

All marked regions of lsml_b200/csrc/fmm.cu are compiled with g++ behind tests/native/
cuda_host_shim.h: the device functions (claim, enqueue, push_entry, expose_neighbours,
publish_change, process_candidate, replay ...) and the kernels that have no block-level
cooperation (fmm_clear_kernel, fmm_init_kernel, fmm_init_band_kernel, fmm_seed_kernel) run as
written; tests/native/fmm_sched_driver.inc restates only fmm_prepare_kernel's scan and the sweep
loop of fmm_march_kernel.  Batches of level sets are evolved over many rebuilds (hints, parking
buckets, tag wrap and hint wipes included) and every rebuild must equal the oracle's sequential
marcher bit for bit, with the histogram invariant the bucket offsets rest on intact.

oracle/fmm_sched_model.c checks the same rules in a separate C restatement of one item; this
test runs the product's own text, on batches."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def sched_lib(tmp_path_factory):
    src = open(os.path.join(ROOT, "lsml_b200", "csrc", "fmm.cu")).read()
    regions = re.findall(r"// \[host-(?:test|sched):begin\][^\n]*\n(.*?)// \[host-(?:test|sched):end\]",
                         src, re.S)
    assert len(regions) == 8, "fmm.cu: expected 8 marked regions"
    text = ('#include "cuda_host_shim.h"\nnamespace lsml {\n' + "\n".join(regions) +
            open(os.path.join(HERE, "native", "fmm_sched_driver.inc")).read())
    out = tmp_path_factory.mktemp("fmm_sched")
    cpp, so = str(out / "fmm_sched_host.cpp"), str(out / "fmm_sched_host.so")
    with open(cpp, "w") as f:
        f.write(text)
    # LSML_HOST_SCHED_DEFS: extra -D switches of fmm.cu's build variants (tools/r2_fmm_variants.sh)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-unused"] +
                          os.environ.get("LSML_HOST_SCHED_DEFS", "").split() + [
                           "-shared", "-fPIC", "-I", os.path.join(HERE, "native"), cpp, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hs_create.restype = ctypes.c_void_p
    lib.hs_rebuild.restype = ctypes.c_long
    return lib


class HostScheduler:
    def __init__(self, lib, shape, batch, dx, ring_min=65536):
        self.lib, self.shape, self.batch = lib, tuple(shape), batch
        self.dx = np.asarray(dx, dtype=np.float64).copy()
        sh = np.asarray(shape, dtype=np.int32)
        P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
        self.h = ctypes.c_void_p(lib.hs_create(ctypes.c_int(len(shape)), P(sh, ctypes.c_int),
                                               ctypes.c_longlong(batch), P(self.dx, ctypes.c_double),
                                               ctypes.c_longlong(ring_min)))

    def set_gen(self, gen):
        self.lib.hs_set_gen(self.h, ctypes.c_longlong(gen))

    def set_slack(self, slack):
        """Hint slack of batch rebuilds: hints are written as depth + (depth >> slack), 0 = off."""
        self.lib.hs_set_slack(self.h, ctypes.c_longlong(slack))

    def set_order(self, order):
        """Evaluation order inside a sweep: 0 as queued, 1 reversed, 2 shuffled."""
        self.lib.hs_set_order(self.h, ctypes.c_int(order))

    def rebuild(self, us, band, prev_band=None, allow_hints=1):
        us = np.ascontiguousarray(us, dtype=np.float64)
        stats = np.zeros((self.batch, 8), dtype=np.uint64)
        for b in range(self.batch):
            stats[b, 0], stats[b, 7] = (us[b] > 0).sum(), (us[b] == 0).sum()
        dist = np.zeros_like(us)
        mask = np.zeros(us.shape, dtype=np.uint8)
        counters = np.zeros(4, dtype=np.int64)
        P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
        if prev_band is None:
            pb, npb = None, 0
        else:
            prev_band = np.ascontiguousarray(prev_band, dtype=np.int64)
            pb, npb = P(prev_band, ctypes.c_longlong), len(prev_band)
        n = self.lib.hs_rebuild(self.h, ctypes.c_int(len(self.shape)), P(us, ctypes.c_double),
                                P(stats, ctypes.c_ulonglong), ctypes.c_double(float(band)), pb,
                                ctypes.c_longlong(npb), ctypes.c_int(allow_hints),
                                P(dist, ctypes.c_double), P(mask, ctypes.c_uint8),
                                P(counters, ctypes.c_longlong))
        return dist, mask.astype(bool), n, counters

    def close(self):
        self.lib.hs_destroy(self.h)


def _items(rs, shape, batch):
    g = np.indices(shape).astype(float)
    us = []
    for _ in range(batch):
        c = [s / 2 + rs.uniform(-2, 2) for s in shape]
        r = np.sqrt(sum((g[a] - c[a]) ** 2 for a in range(len(shape))))
        us.append(2.0 * (r < rs.uniform(0.2, 0.33) * min(shape)) - 1.0)
    return np.array(us)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scheduler_source_reproduces_marcher_on_evolving_batches(oracle, sched_lib, seed):
    rs = np.random.RandomState(seed)
    parked = replays = voxels = 0
    for case in range(6):
        shape = (15, 16, 14) if case % 2 == 0 else (40, 36)
        batch = int(rs.randint(1, 4))
        dx = np.ones(len(shape)) if rs.rand() < 0.5 else rs.uniform(0.6, 1.6, size=len(shape))
        band = float(rs.choice([2.0, 3.0, 4.5]))
        hs = HostScheduler(sched_lib, shape, batch, dx)
        hs.set_gen(int(rs.choice([0, 14, 120, 124, 250, 252, 380, 507])))     # tag wrap / wipes
        hs.set_order(case % 3)              # the result may not depend on the order within a sweep
        hs.set_slack(2 if case >= 3 else 0)  # ... nor on the hint slack batch rebuilds run with
        us = _items(rs, shape, batch)
        if batch > 1 and case % 3 == 0:
            us[-1] = -1.0                           # an item without a contour stays untouched
        prev = None
        for it in range(10):
            d, m, sweeps, ctr = hs.rebuild(us, band, prev)
            assert sweeps > 0, ("no convergence", case, it)
            assert ctr[3] & 15 == 0, ("scheduler invariant broken", int(ctr[3]), case, it)
            for b in range(batch):
                wd, wm = oracle.distance_transform(us[b], band, dx)
                wd = np.where(np.isfinite(wd), wd, 0.0)
                assert np.array_equal(m[b], wm), (case, it, b)
                assert np.array_equal(d[b], wd), (case, it, b, np.abs(d[b] - wd).max())
            parked += int(ctr[2]); replays += int(ctr[1]); voxels += int(m.sum())
            vel = rs.standard_normal(us.shape) * rs.choice([0.1, 0.4, 1.0]) + rs.uniform(-0.3, 0.6)
            us = us.copy()
            us[m] += 0.4 * vel[m] * np.abs(rs.standard_normal(int(m.sum())))
            prev = np.flatnonzero(m.ravel())
        hs.close()
    assert parked > 0                       # the hints were really used
    assert replays < 8 * voxels             # and the reach filter keeps the work bounded


def test_ring_overflow_falls_back_to_scanning_the_candidate_list(oracle, sched_lib):
    """A ring far too small for the work sets: pushes are dropped and reported, the next sweep
    finds its candidates by their epoch (the kernel's fallback, restated in the driver) -- same
    result."""
    rs = np.random.RandomState(7)
    shape, batch, dx, band = (15, 16, 14), 3, np.ones(3), 3.0
    hs = HostScheduler(sched_lib, shape, batch, dx, ring_min=64)
    us = _items(rs, shape, batch)
    prev, fell_back = None, 0
    for it in range(6):
        d, m, sweeps, ctr = hs.rebuild(us, band, prev)
        assert sweeps > 0 and ctr[3] & 15 == 0, (it, sweeps, int(ctr[3]))
        fell_back += int(bool(ctr[3] & 16))
        for b in range(batch):
            wd, wm = oracle.distance_transform(us[b], band, dx)
            assert np.array_equal(m[b], wm) and np.array_equal(d[b], np.where(np.isfinite(wd), wd, 0.0))
        vel = rs.standard_normal(us.shape) * 0.4 + 0.2
        us = us.copy()
        us[m] += 0.4 * vel[m] * np.abs(rs.standard_normal(int(m.sum())))
        prev = np.flatnonzero(m.ravel())
    hs.close()
    assert fell_back == 6


def test_reference_known_answers_through_the_scheduler_source(oracle, sched_lib):
    """lsml/util/test/test_distance_transform.py:11-21 (1-D, band 1) and a full-distance (band 0)
    3-D case on the scheduler source."""
    hs = HostScheduler(sched_lib, (5,), 1, [1.0])
    d, m, sweeps, ctr = hs.rebuild(np.r_[-1, -1, 1, -1, -1.][None], 1.0)
    hs.close()
    assert sweeps > 0 and ctr[3] & 15 == 0
    assert (d[0] == np.r_[0., -0.5, 0.5, -0.5, 0.]).all()
    assert (m[0] == np.r_[False, True, True, True, False]).all()
    arr = np.ones((4, 5, 6))
    arr[2] = -1
    hs = HostScheduler(sched_lib, arr.shape, 1, [1, 1, 1.])
    d, m, sweeps, ctr = hs.rebuild(arr[None], 0.0)
    hs.close()
    wd, wm = oracle.fmm_distance(arr, np.ones(3), 0.0)
    assert sweeps > 0 and m.all() and np.array_equal(d[0], wd)


def test_edge_cases_through_the_scheduler_source(sched_lib):
    """distance_transform.py:36-47: all-zero input -> full mask, zero distances; single-sign
    input -> empty mask (the items are skipped from their exact sign counts)."""
    hs = HostScheduler(sched_lib, (4, 5, 6), 3, [1, 1, 1.])
    us = np.stack([np.zeros((4, 5, 6)), np.ones((4, 5, 6)), -np.ones((4, 5, 6))])
    d, m, sweeps, ctr = hs.rebuild(us, 0.0)
    hs.close()
    assert sweeps >= 0 and ctr[3] & 15 == 0
    assert m[0].all() and (d[0] == 0).all()
    assert not m[1].any() and not m[2].any()
