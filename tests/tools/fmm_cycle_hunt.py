"""CPU hunt for inputs on which the band rebuild's replay iteration does not converge.

Background: a C4 run with 2048 crops per GPU did not finish in round 1.
One candidate cause is an input on which the replay operator has no fixed point (or a 2-cycle in
synchronous order), which the first contributor rule did not exclude (DESIGN.md 4.6
"Contributor rule"; `LSML_ORACLE_RULE_R2=0` runs the hunt with that rule).  This script evolves C4-like crops on the CPU (same generator, features and
forest recipe as bench.py --workload c4, oracle arithmetic) and after every iteration runs the CPU
model of the device scheduler (oracle/fmm_sched_model.c), in place and synchronously, with a
generous sweep budget.  Any non-convergent or mismatching rebuild is saved as an .npz.

    python tests/tools/fmm_cycle_hunt.py --worker 0 --items 400 --out /tmp/hunt

TEST INFRASTRUCTURE (uses oracle/; lives under tests/ for that reason).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", type=int, default=0)
    ap.add_argument("--items", type=int, default=100)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="/tmp/fmm_hunt")
    ap.add_argument("--minutes", type=float, default=60.0)
    ap.add_argument("--size", type=int, default=64, help="edge of the cubic crops")
    ap.add_argument("--radius", type=float, nargs=2, default=(6.0, 18.0), help="object radius range")
    ap.add_argument("--init-radius", type=float, default=5.0)
    ap.add_argument("--jitter", type=float, default=8.0)
    ap.add_argument("--scheduler-source", action="store_true",
                    help="also rebuild every band with the kernel's scheduler source on one host "
                         "thread (tests/test_fmm_scheduler_host.py harness), hints carried along")
    ap.add_argument("--device-source-every", type=int, default=0,
                    help="also iterate the kernel's own source, compiled for the host "
                         "(tests/test_fmm_device_code_host.py), on every N-th rebuild input")
    args = ap.parse_args()
    import bench
    from oracle import oracle as O
    from test_fmm_algorithm import sched_rebuild
    O.build()
    os.makedirs(args.out, exist_ok=True)
    dev_lib = None
    if args.device_source_every:
        import ctypes
        import tempfile
        from test_fmm_device_code_host import build_host_lib, device_fixed_point
        legacy = os.environ.get("LSML_ORACLE_RULE_R2") == "0"
        dev_lib = ctypes.CDLL(build_host_lib(tempfile.mkdtemp(), legacy))
    n_dev = 0
    sched_lib = None
    if args.scheduler_source:
        import ctypes
        import re
        import subprocess
        import tempfile
        from test_fmm_scheduler_host import HostScheduler
        src = open(os.path.join(ROOT, "lsml_b200", "csrc", "fmm.cu")).read()
        regions = re.findall(r"// \[host-(?:test|sched):begin\][^\n]*\n(.*?)// \[host-(?:test|sched):end\]",
                             src, re.S)
        native = os.path.join(ROOT, "tests", "native")
        tmpd = tempfile.mkdtemp()
        with open(os.path.join(tmpd, "s.cpp"), "w") as f:
            f.write('#include "cuda_host_shim.h"\nnamespace lsml {\n' + "\n".join(regions) +
                    open(os.path.join(native, "fmm_sched_driver.inc")).read())
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-unused",
                               "-shared", "-fPIC", "-I", native, os.path.join(tmpd, "s.cpp"),
                               "-o", os.path.join(tmpd, "s.so")])
        sched_lib = ctypes.CDLL(os.path.join(tmpd, "s.so"))
        sched_lib.hs_create.restype = ctypes.c_void_p
        sched_lib.hs_rebuild.restype = ctypes.c_long
    n_sched = 0
    shape, band, dx = (args.size,) * 3, 3.0, np.ones(3)
    CHUNK = 16          # items are generated 16 at a time
    g = np.indices(shape).astype(np.float64)
    rc = np.sqrt(sum((g[a] - shape[a] / 2.0) ** 2 for a in range(3)))
    u0 = np.where(rc < args.init_radius, 1.0, -1.0)
    specs = bench.SPECS
    # u-independent planes once per item; everything else per iteration (oracle expressions)
    def planes_of(img):
        return [O.image_sample_plane(img, 0.0, dx), O.image_sample_plane(img, 2.0, dx),
                O.image_edge_plane(img, 0.0, dx), O.image_edge_plane(img, 2.0, dx)]

    def features(u, pl, mask):
        nb = int(mask.sum())
        cols = [p[mask] for p in pl]
        for p in (pl[0], pl[1]):
            cols.append(np.full(nb, p[mask].mean()))
        for p in (pl[0], pl[1]):
            cols.append(np.full(nb, p[mask].std()))
        pos = u > 0
        cnt = pos.sum()
        com = np.array([g[a][pos].sum() / cnt for a in range(3)])
        cols.append(np.sqrt(sum((g[a][mask] - com[a]) ** 2 for a in range(3))))
        for a in range(3):
            cols.append(np.full(nb, com[a]))
            cols.append(np.full(nb, ((g[a][pos] - com[a]) ** 2).sum() / cnt))
        cols.append(np.full(nb, float(cnt)))
        return np.stack(cols, axis=1)

    t_end = time.time() + 60 * args.minutes
    reg = None
    n_reb = n_bad = 0
    worst = 0
    log = open(os.path.join(args.out, "worker%d.log" % args.worker), "w")
    imgs = None
    for b in range(args.items):
        if b % CHUNK == 0:
            imgs = bench._blob_items(CHUNK, shape, 10000 + 97 * args.worker + 7919 * (b // CHUNK),
                                     args.radius[0], args.radius[1], args.jitter)
        img = O.normalize_image(imgs[b % CHUNK])
        pl = planes_of(img)
        u = u0.copy()
        dist, mask = O.distance_transform(u, band, dx)
        lvl = [np.zeros(u.size, dtype=np.uint16) for _ in range(2)]
        hist = [np.zeros(512, dtype=np.int64) for _ in range(2)]
        gen = [np.array([b * 37], dtype=np.int64) for _ in range(2)]
        prev = None
        hs = None
        if sched_lib is not None:
            hs = HostScheduler(sched_lib, shape, 1, dx)
            hs.set_gen(b * 37)
            hs.set_order(b % 3)
            hs.rebuild(u[None], band, None)          # the band of u0 (dense search), as the engine does
        for it in range(args.iters):
            if not mask.any():
                break
            X = features(u, pl, mask)
            if reg is None:
                y = X[:, 1] + 0.2 * np.random.RandomState(1).standard_normal(X.shape[0])
                reg = bench._fit_forest(X, y, 5 + args.worker)
            v = np.zeros(u.shape)
            v[mask] = reg.predict(X)
            gm = O.gmag_os(u, v, mask, dx)
            u = u.copy()
            u[mask] += 0.4 * v[mask] * gm[mask]
            prev = np.flatnonzero(mask.ravel())
            wd, wm = O.distance_transform(u, band, dx)
            for mode in (0, 1):
                d, m, sweeps, ctr = sched_rebuild(O, u, dx, band, lvl[mode], hist[mode], gen[mode],
                                                  prev, mode, max_sweeps=3000)
                n_reb += 1
                worst = max(worst, int(sweeps))
                ok = sweeps > 0 and ctr[3] == 0 and np.array_equal(m, wm)
                exact = ok and np.array_equal(d, wd)
                if not exact:
                    n_bad += 1
                    kind = "NOCONV" if sweeps <= 0 else ("MASK" if not ok else "DIST")
                    name = os.path.join(args.out, "case_w%d_b%d_it%d_m%d_%s.npz"
                                        % (args.worker, b, it, mode, kind))
                    np.savez_compressed(name, u=u, prev=prev, dx=dx, band=band)
                    log.write("%s item %d it %d mode %d sweeps %d ctr %s maxdiff %g\n" % (
                        kind, b, it, mode, sweeps, ctr.tolist(),
                        float(np.abs(d - wd).max()) if sweeps > 0 else -1.0))
                    log.flush()
            if dev_lib is not None and (n_reb // 2) % args.device_source_every == 0:
                d, m, sweeps = device_fixed_point(dev_lib, u, dx, band, 0)
                n_dev += 1
                if sweeps <= 0 or not np.array_equal(m, wm) or not np.array_equal(d, wd):
                    n_bad += 1
                    np.savez_compressed(os.path.join(args.out, "case_w%d_b%d_it%d_DEVSRC.npz"
                                                     % (args.worker, b, it)), u=u, prev=prev, dx=dx,
                                        band=band)
                    log.write("DEVSRC item %d it %d sweeps %d\n" % (b, it, sweeps))
                    log.flush()
            if hs is not None:
                d, m, sweeps, ctr = hs.rebuild(u[None], band, prev)
                n_sched += 1
                if (sweeps <= 0 or (ctr[3] & 15) or not np.array_equal(m[0], wm)
                        or not np.array_equal(d[0], np.where(np.isfinite(wd), wd, 0.0))):
                    n_bad += 1
                    np.savez_compressed(os.path.join(args.out, "case_w%d_b%d_it%d_SCHED.npz"
                                                     % (args.worker, b, it)), u=u, prev=prev, dx=dx,
                                        band=band)
                    log.write("SCHED item %d it %d sweeps %d ctr %s\n" % (b, it, sweeps, ctr.tolist()))
                    log.flush()
            dist, mask = wd, wm
        if hs is not None:
            hs.close()
        log.write("item %d done: rebuilds %d bad %d worst sweeps %d scheduler-source rebuilds %d "
                  "device-source checks %d\n" % (b, n_reb, n_bad, worst, n_sched, n_dev))
        log.flush()
        if time.time() > t_end:
            break
    log.write("END rebuilds %d bad %d worst sweeps %d scheduler-source rebuilds %d "
              "device-source checks %d\n" % (n_reb, n_bad, worst, n_sched, n_dev))
    log.close()


if __name__ == "__main__":
    main()
