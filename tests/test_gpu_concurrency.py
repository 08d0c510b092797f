"""Many library contexts in flight on one GPU (own stream + host thread each).  The single-context suites cannot see
block-level races that only show when other kernels share the SMs (the fcluster loop-flag race fixed in round 1 was
found this way); this is the reduced form of scripts/stress_async.py."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_async", [False, True])
def test_clustering_from_many_contexts(pkg, synth, oracle, use_async):
    K, calls, Cn, S, D, F = 12, 12, 120, 3, 192, 293
    jobs = []
    for i in range(K):
        ctx = pkg.Context(0)
        emb, _ = synth.embeddings(100 + i, Cn, S, D, n_speakers=3 + i % 3, nan_frac=0.05)
        b = (synth.segmentations(200 + i, Cn, F, S) > 0.5).astype(np.float64)
        keep = np.flatnonzero(~np.isnan(emb.reshape(Cn * S, D)[:, 0])).astype(np.int32)
        _, want, _, kw = oracle.clustering_stage(emb, b)
        jobs.append(dict(ctx=ctx, keep=keep, want=want, kw=kw, d_e=ctx.to_device(emb), d_b=ctx.to_device(b),
                         d_h=ctx.malloc(4 * Cn * S), d_k=ctx.malloc(4), p=ctx.cluster_params()))

    def run(j):
        ctx = j["ctx"]
        kc = C.c_int(-1)
        ctx._check(ctx.L.sd_status_reset(ctx.h))
        for _ in range(calls):
            if use_async:
                ctx._check(ctx.L.sd_clustering_async_dev(ctx.h, j["d_e"], Cn, S, D, C.byref(j["p"]), pkg._ptr(j["keep"]),
                                                         j["keep"].size, j["d_b"], F, j["d_h"], None, 0, j["d_k"]))
            else:
                ctx._check(ctx.L.sd_clustering_dev(ctx.h, j["d_e"], Cn, S, D, C.byref(j["p"]), j["d_b"], F, j["d_h"],
                                                   None, 0, C.byref(kc)))
        ctx._check(ctx.L.sd_status_check(ctx.h))
        hard = np.empty((Cn, S), np.int32)
        ctx.d2h(hard, j["d_h"])
        if use_async:
            k = np.empty(1, np.int32)
            ctx.d2h(k, j["d_k"])
            kc.value = int(k[0])
        return kc.value == j["kw"] and np.array_equal(hard, j["want"])

    with ThreadPoolExecutor(max_workers=K) as pool:
        ok = list(pool.map(run, jobs))
    for j in jobs:
        for q in ("d_e", "d_b", "d_h", "d_k"):
            j["ctx"].free(j[q])
        j["ctx"].close()
    assert all(ok), ok
