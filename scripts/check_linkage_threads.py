import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as ge
from oracle.oracle import Oracle
pkg, synth, o = ge.load_package(), ge.load_synth(), Oracle()
ctx = pkg.Context(0)
ok = True
for threads in (512, 1024, 256):
    ctx.set_option(2, threads)
    for N, D, seed in ((2, 4, 0), (33, 16, 2), (319, 192, 3), (1136, 192, 4), (1773, 192, 5), (4800, 64, 6)):
        x, _ = synth.stress_embeddings(seed, N, D, 5)
        xn = o.normalize(x)
        same = np.array_equal(ctx.linkage(xn), o.linkage(xn))
        ok &= same
        print("threads", threads, "N", N, "bit-exact" if same else "MISMATCH", flush=True)
    rng = np.random.default_rng(11)
    base = rng.standard_normal((40, 6))
    xx = np.concatenate([base, base, base[:9]])
    same = np.array_equal(ctx.linkage(xx), o.linkage(xx))
    ok &= same
    print("threads", threads, "ties", "bit-exact" if same else "MISMATCH", flush=True)
    x = o.normalize(np.random.default_rng(9).standard_normal((700, 24)))
    same = np.array_equal(ctx.linkage(x), o.linkage(x)); ok &= same
    print("threads", threads, "unclustered", "bit-exact" if same else "MISMATCH", flush=True)
print("PRE TEST", "PASSED" if ok else "FAILED")
