"""Profiling helper: the STFT kernel at the bench workload size (cfg2), resident buffers.
usage: python scripts/prof_stft.py [variant] [items]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg = ge.load_package()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
items = int(sys.argv[2]) if len(sys.argv) > 2 else 1773
L = 160000
T = 1 + L // 160
ctx = pkg.Context(0)
ctx.set_option(3, variant)
rng = np.random.default_rng(0)
wav = (0.1 * rng.standard_normal((items, L))).astype(np.float32)
d_in = ctx.to_device(wav)
d_out = ctx.malloc(items * T * 402 * 4)
p = ctx.stft_params()
for _ in range(3):
    ctx.stft_dev(d_in, items, L, d_out, p)
ctx.sync()
ms = []
for _ in range(10):
    ctx.timer_start(0)
    ctx.stft_dev(d_in, items, L, d_out, p)
    ctx.timer_stop(0)
    ms.append(ctx.timer_ms(0))
b = items * (4 * L + 4 * T * 402)
print("variant %d: %d items, median %.4f ms, best %.4f ms -> %.0f GB/s (median), %.1f%% of 6548 GB/s"
      % (variant, items, np.median(ms), min(ms), b / np.median(ms) / 1e6, 100 * b / np.median(ms) / 1e6 / 6548.2))
