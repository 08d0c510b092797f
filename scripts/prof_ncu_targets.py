"""Workloads for `ncu --set full` captures of the kernels the round-1 verdict asked evidence for.
usage: python scripts/prof_ncu_targets.py fbank|pdist_tc|pdist_f64|aggregate|mask_compact|binarize"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge

pkg, synth = ge.load_package(), ge.load_synth()
ctx = pkg.Context(0)
what = sys.argv[1]
vp = C.c_void_p
if what == "fbank":
    items, L = 600, 160000
    wav = synth.fbank_items(1, 8, L)
    wav = np.tile(wav, (items // 8, 1))
    d_in, d_len = ctx.to_device(wav), ctx.to_device(np.ones(items, np.float32))
    d_out = ctx.malloc(items * (1 + L // 160) * 80 * 4)
    p = ctx.fbank_params()
    for _ in range(2):
        ctx._check(ctx.L.sd_fbank_dev(ctx.h, vp(d_in), items, L, vp(d_len), C.byref(p), vp(d_out)))
    ctx.sync()
elif what in ("pdist_tc", "pdist_f64"):
    N, D = 10000, 256
    x, _ = synth.stress_embeddings(205, N, D, 12)
    xn = ctx.normalize_embeddings(x)
    for _ in range(2):
        ctx.pdist(xn, 1 if what == "pdist_tc" else 0)
elif what == "aggregate":
    Cn, F, K = 3591, 589, 3  # cfg3 geometry
    sc = synth.segmentations(3, Cn, F, K).astype(np.float64)
    for _ in range(2):
        ctx.aggregate(sc, (0.0, 1.0, 10.0, 16000 * 3600), missing=0.0, skip_average=True)
elif what == "binarize":
    Cn, F, K = 3591, 589, 3
    seg = synth.segmentations(3, Cn, F, K)
    for _ in range(2):
        b = ctx.binarize_swf(seg)
        ctx.speaker_count(b, chunks=(0.0, 1.0, 10.0, 1))
elif what == "mask_compact":
    Cn, K, L, step, F = 591, 3, 160000, 16000, 589
    n = (Cn - 1) * step + L
    wave = np.tile(synth.waveform(4, 20.0), n // 320000 + 1)[:n]
    masks = (np.tile(synth.segmentations(8, 64, F, 3), (Cn // 64 + 1, 1, 1))[:Cn].transpose(0, 2, 1).reshape(Cn * K, F) > 0.5).astype(np.float32)
    for _ in range(2):
        ctx.mask_compact_file(wave, masks, Cn, K, L, step)
print("done", what)
