#!/bin/bash
# STFT FIFO chain in the batch: chain depth x files in flight x merge-loop kernel
cd "$GRAFT_REPO_ROOT" || exit 1
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs"
run() {  # name, env..., files
  name=$1; files=$2; shift 2
  env "$@" SDB_BATCH_TRACE=gpurun_out/r02_trace_$name.csv $B --files $files > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  python -c "
import json,sys
for l in open('gpurun_out/r02_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3))
"
}
run chain0_f16 16 SDB_BATCH_STFT_CHAIN=0
run chain1_f16 16 SDB_BATCH_STFT_CHAIN=1
run chain2_f16 16 SDB_BATCH_STFT_CHAIN=2
run chain1_f24 24 SDB_BATCH_STFT_CHAIN=1
run chain2_f24 24 SDB_BATCH_STFT_CHAIN=2
run chain1_f32 32 SDB_BATCH_STFT_CHAIN=1
run chain1_f24_onecta 24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run chain1_f32_onecta 32 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run chain1_f16_w4 16 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=6=4
