"""Profiling helper: linkage + fcluster at the BASELINE config sizes (cfg2 1.8k, cfg3 10.8k x 192, cfg5 50k x 256).
usage: python scripts/prof_cluster_sizes.py [max_n] [linkage_threads]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg = ge.load_package()
ctx = pkg.Context(0)
max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
if len(sys.argv) > 2:
    ctx.set_option(2, int(sys.argv[2]))  # SD_OPT_LINKAGE_THREADS
for name, N, D, k in (("cfg2", 1773, 192, 4), ("cfg3", 10773, 192, 6), ("cfg5", 50000, 256, 12)):
    if N > max_n:
        continue
    rng = np.random.default_rng(N)
    cen = rng.standard_normal((k, D))
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    x = cen[rng.integers(0, k, N)] + 0.04 * rng.standard_normal((N, D))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    d_x = ctx.to_device(np.ascontiguousarray(x, np.float64))
    d_Z = ctx.malloc(8 * 4 * (N - 1))
    ctx.debug_counters()
    for rep in range(2):
        ctx.timer_start(0)
        ctx._check(ctx.L.sd_linkage_dev(ctx.h, d_x, N, D, d_Z))
        ctx.timer_stop(0)
        ms = ctx.timer_ms(0)
    c = ctx.debug_counters().astype(float)
    print("%s N=%d D=%d: linkage (pdist + merges) %.1f ms = %.2f us/merge; %.2f revalidations/merge, refills %d, "
          "fallbacks %d" % (name, N, D, ms, ms * 1e3 / (N - 1), c[0] / max(c[7], 1), int(c[1]), int(c[2])), flush=True)
    m = max(c[7], 1)  # counters accumulate over the two repetitions, and so does the merge count
    print("    cycles per merge (control warp view): decide %.0f, revalidation %.0f, sweep+fence+publish %.0f, "
          "mbarrier wait %.0f; stage times: pdist %.2f ms, merges %.2f ms"
          % (c[3] / m, c[4] / m, c[5] / m, c[6] / m, *ctx.linkage_stage_ms()), flush=True)
    ctx.free(d_x)
    ctx.free(d_Z)
