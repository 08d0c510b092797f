"""Profiling helper: clustering stage of the bench workload (cfg2), a few repetitions, with linkage counters.
usage: python scripts/prof_cluster.py [n_speakers] [linkage_threads]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge
import bench

pkg = ge.load_package()
synth = ge.load_synth()
geo = bench.geometry()
n_spk = int(sys.argv[1]) if len(sys.argv) > 1 else 4
emb, _ = synth.embeddings(202, geo["C"], geo["S"], geo["D"], n_speakers=n_spk)
ctx = pkg.Context(0)
if len(sys.argv) > 2:
    ctx.set_option(2, int(sys.argv[2]))
ctx.clustering(emb)
ctx.debug_counters()
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
    hard, k = ctx.clustering(emb)
dt = (time.perf_counter() - t0) / reps
c = ctx.debug_counters().astype(float)
N = int((~np.isnan(emb[:, :, 0])).sum())
m = max(c[7], 1.0)
print("N=%d clusters=%d: %.2f ms per clustering; per merge: %.2f stale revalidations, cycles decide %.0f / revalidate %.0f / "
      "sweep+publish %.0f / wait %.0f; refills %d; fallbacks to heap kernel %d"
      % (N, k, dt * 1e3, c[0] / m, c[3] / m, c[4] / m, c[5] / m, c[6] / m, int(c[1]), int(c[2])))
