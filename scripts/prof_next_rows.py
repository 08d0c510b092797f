"""Profiling helper: the SURVEY 8f rows at the bench file geometry (uses bench.next_rows_timing)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge
import bench

pkg = ge.load_package()
synth = ge.load_synth()
geo = bench.geometry()
ctx = pkg.Context(0)
C_, F, S = geo["C"], geo["F"], geo["S"]
seg = synth.segmentations(1102, C_, F, S)
emb, _ = synth.embeddings(202, C_, S, geo["D"], n_speakers=4)
chunks = pkg.Window(0.0, bench.WORKLOAD["step_s"], bench.WORKLOAD["window_s"], int(bench.WORKLOAD["audio_seconds"] * 16000))
frames = pkg.Window(0.0, pkg.FRAME_STEP, pkg.FRAME_DURATION, 0)
d_seg = ctx.to_device(seg)
d_bin = ctx.malloc(C_ * F * S * 8)
ctx._check(ctx.L.sd_binarize_dev(ctx.h, C.c_void_p(d_seg), C_, F, S, pkg.ONSET, 0, C.c_void_p(d_bin)))
b = np.empty((C_, F, S), np.float64)
ctx.d2h(b, d_bin)
hard, k = ctx.clustering(emb, b)
d_hard = ctx.to_device(np.ascontiguousarray(hard, np.int32))
print(bench.next_rows_timing(ctx, pkg, synth, geo, d_seg, d_bin, d_hard, chunks, frames))
