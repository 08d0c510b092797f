#!/bin/bash
# SM partition (green contexts) in the batch: narrow size x merge-loop kernel x chain depth
cd "$GRAFT_REPO_ROOT" || exit 1
SDB_BATCH_NARROW_SMS=16 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs"
run() {  # name, files, env...
  name=$1; files=$2; shift 2
  env "$@" SDB_BATCH_TRACE=gpurun_out/r02_trace_$name.csv $B --files $files > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  grep -h "sd_batch" gpurun_out/r02_bench_$name.err | head -2
  python -c "
import json,sys
for l in open('gpurun_out/r02_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'e2e', round(d['e2e']['value']))
"
}
run gp_n16_one_c1_f16 16 SDB_BATCH_NARROW_SMS=16 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gp_n16_one_c0_f16 16 SDB_BATCH_NARROW_SMS=16 SDB_BATCH_STFT_CHAIN=0 SDB_BATCH_OPTS=4=0
run gp_n16_clu_c1_f16 16 SDB_BATCH_NARROW_SMS=16 SDB_BATCH_STFT_CHAIN=1
run gp_n8_one_c1_f16 16 SDB_BATCH_NARROW_SMS=8 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gp_n24_one_c1_f16 16 SDB_BATCH_NARROW_SMS=24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gp_n16_one_c1_f24 24 SDB_BATCH_NARROW_SMS=16 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gp_n24_clu_c1_f16 16 SDB_BATCH_NARROW_SMS=24 SDB_BATCH_STFT_CHAIN=1
run gp_n16_one_c2_f16 16 SDB_BATCH_NARROW_SMS=16 SDB_BATCH_STFT_CHAIN=2 SDB_BATCH_OPTS=4=0
