#!/bin/bash
# full GPU test suite + the default bench (24 files in flight, one-CTA merge loop, chained STFTs)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_f.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_f.log
python bench.py > gpurun_out/r02_bench_v7.json 2> gpurun_out/r02_bench_v7.err
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs"
SDB_BATCH_TRACE=gpurun_out/r02_trace_v7_f24.csv $B > gpurun_out/r02_bench_v7_f24.json 2>&1
$B --files 32 > gpurun_out/r02_bench_v7_f32.json 2>&1
$B --files 16 > gpurun_out/r02_bench_v7_f16.json 2>&1
SDB_BATCH_LINKAGE_CLUSTER=0 $B --files 16 > gpurun_out/r02_bench_v7_f16_onecta.json 2>&1
for t in v7 v7_f24 v7_f32 v7_f16 v7_f16_onecta; do
  python -c "
import json,sys
for l in open('gpurun_out/r02_bench_$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$t', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['config']['batch_config'])
"
done
