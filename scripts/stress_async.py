"""Stress helper: many contexts on their own streams / host threads, deep queues of sd_clustering_async_dev calls.
usage: python scripts/stress_async.py [contexts] [calls] [chunks]"""
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg, synth = ge.load_package(), ge.load_synth()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
Cn = int(sys.argv[3]) if len(sys.argv) > 3 else 120
S, D, F = 3, 192, 293
jobs = []
for i in range(K):
    ctx = pkg.Context(0)
    ctx.set_option(4, int(os.environ.get("CLUSTER", "1")))  # SD_OPT_LINKAGE_CLUSTER: 0 = one-CTA merge loop
    emb, _ = synth.embeddings(100 + i, Cn, S, D, n_speakers=3 + i % 3, nan_frac=0.05)
    seg = synth.segmentations(200 + i, Cn, F, S)
    b = (seg > 0.5).astype(np.float64)
    keep = np.flatnonzero(~np.isnan(emb.reshape(Cn * S, D)[:, 0])).astype(np.int32)
    want, kw = ctx.clustering(emb, b)
    jobs.append(dict(ctx=ctx, emb=emb, keep=keep, want=want, kw=kw, d_e=ctx.to_device(emb), d_b=ctx.to_device(b),
                     d_h=ctx.malloc(4 * Cn * S), d_k=ctx.malloc(4), p=ctx.cluster_params()))


def run(j):
    ctx = j["ctx"]
    ctx._check(ctx.L.sd_status_reset(ctx.h))
    kc = C.c_int()
    for _ in range(calls):
        if os.environ.get("SYNC", "0") == "1":  # the synchronous entry point under the same load
            ctx._check(ctx.L.sd_clustering_dev(ctx.h, j["d_e"], Cn, S, D, C.byref(j["p"]), j["d_b"], F, j["d_h"], None, 0,
                                               C.byref(kc)))
            ctx.h2d(j["d_k"], np.array([kc.value], np.int32))
            continue
        ctx._check(ctx.L.sd_clustering_async_dev(ctx.h, j["d_e"], Cn, S, D, C.byref(j["p"]), pkg._ptr(j["keep"]),
                                                 j["keep"].size, j["d_b"], F, j["d_h"], None, 0, j["d_k"]))
    ctx._check(ctx.L.sd_status_check(ctx.h))
    hard, k = np.empty((Cn, S), np.int32), np.empty(1, np.int32)
    ctx.d2h(hard, j["d_h"])
    ctx.d2h(k, j["d_k"])
    assert k[0] == j["kw"] and np.array_equal(hard, j["want"])
    return True


with ThreadPoolExecutor(max_workers=K) as pool:
    ok = list(pool.map(run, jobs))
print("stress ok:", len(ok), "contexts x", calls, "asynchronous clusterings of", Cn * S, "rows")
