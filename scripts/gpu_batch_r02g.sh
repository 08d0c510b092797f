#!/bin/bash
# more files in flight: partitioned vs unpartitioned, one-CTA merge loop
cd "$GRAFT_REPO_ROOT" || exit 1
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
run gq_n24_one_c1_f24 24 SDB_BATCH_NARROW_SMS=24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gq_n24_one_c1_f32 32 SDB_BATCH_NARROW_SMS=24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gq_n32_one_c1_f32 32 SDB_BATCH_NARROW_SMS=32 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gq_n0_one_c1_f24 24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gq_n0_one_c0_f24 24 SDB_BATCH_STFT_CHAIN=0 SDB_BATCH_OPTS=4=0
run gq_n0_one_c1_f32 32 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
run gq_n0_one_c0_f32 32 SDB_BATCH_STFT_CHAIN=0 SDB_BATCH_OPTS=4=0
run gq_n0_one_c1_f40 40 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
