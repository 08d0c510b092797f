"""Profiling helper: K files processed concurrently (one thread + context + stream each), per-stage event times
under contention.  usage: python scripts/prof_concurrent.py [files] [steps] [free]"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as ge
import bench

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
free_running = len(sys.argv) > 3 and sys.argv[3] == "free"
pkg, synth, geo = ge.load_package(), ge.load_synth(), bench.geometry()
torch.cuda.set_device(0)
base = synth.fbank_items(102, geo["items"], geo["L"])
jobs = [bench.FileJob(pkg, synth, geo, 0, 17 * i, base, i, torch) for i in range(K)]
for j in jobs:
    j.use_async = os.environ.get("ASYNC", "0") == "1"  # sd_clustering_async_dev instead of sd_clustering_dev
pool = ThreadPoolExecutor(max_workers=K)
acc = np.zeros((K, 5))


def run(j, idx):
    t0 = time.perf_counter()
    j.step(timed=True)
    j.ctx.sync()
    acc[idx, 0] += time.perf_counter() - t0
    for s in (1, 2, 3, 4):
        acc[idx, s] += j.ctx.timer_ms(s)


def run_many(j, idx, n):
    for _ in range(n):
        run(j, idx)


if free_running:
    for f in [pool.submit(run_many, j, i, 3) for i, j in enumerate(jobs)]:
        f.result()
    acc[:] = 0
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for f in [pool.submit(run_many, j, i, steps) for i, j in enumerate(jobs)]:
        f.result()
else:
    for it in range(3 + steps):
        if it == 3:
            acc[:] = 0
            torch.cuda.synchronize()
            t_all = time.perf_counter()
        for f in [pool.submit(run, j, i) for i, j in enumerate(jobs)]:
            f.result()
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / steps
m = acc.mean(0) / steps
print("free-running" if free_running else "lock-step", end=" ")
print("files=%d: %.2f ms per step (%.2f ms per file); per file under contention: host wall %.2f ms, stft %.2f, "
      "binarize+count %.2f, clustering %.2f, aggregate %.3f ms" % (K, wall * 1e3, wall * 1e3 / K, m[0] * 1e3, m[1], m[2], m[3], m[4]))
