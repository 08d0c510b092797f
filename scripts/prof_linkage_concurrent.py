"""Profiling helper: K independent linkages (pdist + merge loop, cfg2 size) enqueued from ONE host thread on K
contexts / streams; per-problem event time and the wall time of the group.  Isolates how much concurrent merge loops
slow each other (no STFT, no host threads).  usage: python scripts/prof_linkage_concurrent.py [N] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg = ge.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1683
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
D = 192
KMAX = 16
ctxs, dx, dz = [], [], []
for i in range(KMAX):
    rng = np.random.default_rng(100 + i)
    cen = rng.standard_normal((4, D))
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    x = cen[rng.integers(0, 4, N)] + 0.04 * rng.standard_normal((N, D))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    c = pkg.Context(0)
    for opt in os.environ.get("SDB_OPTS", "").split(","):
        if "=" in opt:
            k, v = opt.split("=")
            c.set_option(int(k), int(v))
    ctxs.append(c)
    dx.append(c.to_device(np.ascontiguousarray(x, np.float64)))
    dz.append(c.malloc(8 * 4 * (N - 1)))
for c, a, b in zip(ctxs, dx, dz):  # warm-up (workspace allocation, attribute opt-in)
    c._check(c.L.sd_linkage_dev(c.h, a, N, D, b))
    c.sync()
for K in (1, 2, 4, 8, 12, 16):
    best = None
    for _ in range(reps):
        for c in ctxs[:K]:
            c.sync()
        t0 = time.perf_counter()
        for c, a, b in zip(ctxs[:K], dx[:K], dz[:K]):
            c.timer_start(0)
            c._check(c.L.sd_linkage_dev(c.h, a, N, D, b))
            c.timer_stop(0)
        for c in ctxs[:K]:
            c.sync()
        wall = (time.perf_counter() - t0) * 1e3
        per = [c.timer_ms(0) for c in ctxs[:K]]
        if best is None or wall < best[0]:
            best = (wall, per)
    print("K=%2d concurrent linkages N=%d: wall %.2f ms; per problem min %.2f / mean %.2f / max %.2f ms (%.2f us/merge)"
          % (K, N, best[0], min(best[1]), float(np.mean(best[1])), max(best[1]), float(np.mean(best[1])) * 1e3 / (N - 1)),
          flush=True)
