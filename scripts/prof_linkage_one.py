"""One linkage (pdist + merge loop) at a given size, for ncu captures.  usage: prof_linkage_one.py N D [threads]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
pkg, synth = ge.load_package(), ge.load_synth()
N, D = int(sys.argv[1]), int(sys.argv[2])
ctx = pkg.Context(0)
if len(sys.argv) > 3:
    ctx.set_option(2, int(sys.argv[3]))
x, _ = synth.stress_embeddings(200 + D + N % 97, N, D, 6 if N < 20000 else 12)
x /= np.linalg.norm(x, axis=1, keepdims=True)
d_x = ctx.to_device(np.ascontiguousarray(x, np.float64))
d_Z = ctx.malloc(8 * 4 * (N - 1))
for _ in range(2):
    ctx._check(ctx.L.sd_linkage_dev(ctx.h, d_x, N, D, d_Z))
    ctx.sync()
print("done", ctx.linkage_stage_ms())
