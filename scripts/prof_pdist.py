"""Tensor-core pdist (3xTF32 Gram GEMM, tcgen05 + TMA) vs the exact fp64 kernel: error and time.
usage: python scripts/prof_pdist.py [N] [D]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg = ge.load_package()
synth = ge.load_synth()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1683
D = int(sys.argv[2]) if len(sys.argv) > 2 else 192
ctx = pkg.Context(0)
C = (N + 2) // 3
emb, _ = synth.embeddings(7, C, 3, D, n_speakers=6, nan_frac=0.0, tiny=())
x = emb.reshape(-1, D)[:N]
xn = ctx.normalize_embeddings(x)
ex = ctx.pdist(xn, 0)
tc = ctx.pdist(xn, 1)
err = np.abs(tc - ex)
print("N=%d D=%d: max |d_tc - d_exact| = %.3e, mean %.3e, pairs with err > 1e-6: %d of %d; min distance %.3f"
      % (N, D, err.max(), err.mean(), int((err > 1e-6).sum()), err.size, ex.min()))
for mode in (0, 1):
    ctx.pdist(xn, mode)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.pdist(xn, mode)
    print("mode %d host-call time %.2f ms (includes H2D of x and D2H of the condensed matrix)" % (mode, (time.perf_counter() - t0) / 3 * 1e3))
