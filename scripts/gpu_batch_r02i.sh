#!/bin/bash
# fbank v2 (4 CTAs/SM, register Hamming, width-sorted mel warps): tests + timing + ncu; launch list of the new default
# bench; ncu of the one-CTA merge loop (now the batch default)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time_v2.log
python scripts/prof_stft.py 0 1773 2>&1 | tail -2 | tee gpurun_out/r02_stft_dyn.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches_v7.csv python bench.py --files 12 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_under_ncu_v7.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank400_kernel -s 2 -c 1 -f -o gpurun_out/r02_fbank_v2 python scripts/prof_ncu_targets.py fbank > gpurun_out/r02_ncu_fbank_v2.log 2>&1
cat > /tmp/one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as ge
pkg, synth = ge.load_package(), ge.load_synth()
ctx = pkg.Context(0)
ctx.set_option(4, 0)  # one-CTA merge loop
N, D = 1683, 192
x, _ = synth.stress_embeddings(200 + D + N % 97, N, D, 6)
x /= np.linalg.norm(x, axis=1, keepdims=True)
d_x = ctx.to_device(np.ascontiguousarray(x, np.float64))
d_Z = ctx.malloc(8 * 4 * (N - 1))
for _ in range(3):
    ctx._check(ctx.L.sd_linkage_dev(ctx.h, d_x, N, D, d_Z))
    ctx.sync()
print("done", ctx.linkage_stage_ms())
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linkage_fast_kernel -s 1 -c 1 -f -o gpurun_out/r02_linkage_fast_cfg2 python /tmp/one.py > gpurun_out/r02_ncu_linkage_fast.log 2>&1
tail -2 gpurun_out/r02_ncu_linkage_fast.log
ls -la gpurun_out/*.ncu-rep
