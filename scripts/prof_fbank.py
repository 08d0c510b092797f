"""Profiling helper: fused log-mel front-end (sd_fbank_dev) at the bench workload size, resident buffers."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg = ge.load_package()
items = int(sys.argv[1]) if len(sys.argv) > 1 else 1773
L = 160000
T = 1 + L // 160
ctx = pkg.Context(0)
if len(sys.argv) > 2:
    ctx.set_option(3, int(sys.argv[2]))  # SD_OPT_STFT_VARIANT: 8 = mel projection by (frame, part) threads
rng = np.random.default_rng(0)
wav = (0.1 * rng.standard_normal((items, L))).astype(np.float32)
d_in = ctx.to_device(wav)
d_len = ctx.to_device(np.ones(items, np.float32))
d_out = ctx.malloc(items * T * 80 * 4)
p = ctx.fbank_params()
vp = C.c_void_p
for _ in range(3):
    ctx._check(ctx.L.sd_fbank_dev(ctx.h, vp(d_in), items, L, vp(d_len), C.byref(p), vp(d_out)))
ctx.sync()
ms = []
for _ in range(10):
    ctx.timer_start(0)
    ctx._check(ctx.L.sd_fbank_dev(ctx.h, vp(d_in), items, L, vp(d_len), C.byref(p), vp(d_out)))
    ctx.timer_stop(0)
    ms.append(ctx.timer_ms(0))
b = items * (4 * L + 4 * T * 80 + 4)
print("fused fbank: %d items, median %.4f ms -> %.0f GB/s algorithmic (%.1f%% of 6548 GB/s); %.0f audio-s/s of 10 s items"
      % (items, np.median(ms), b / np.median(ms) / 1e6, 100 * b / np.median(ms) / 1e6 / 6548.2, items * 10 / (np.median(ms) / 1e3)))
