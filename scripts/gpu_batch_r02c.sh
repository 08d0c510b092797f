#!/bin/bash
# timeline of the batch step + one-CTA merge loop under the batch
cd "$GRAFT_REPO_ROOT" || exit 1
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs"
$B --files 16 > gpurun_out/r02_bench_ctl_f16.json 2> gpurun_out/r02_bench_ctl_f16.err
SDB_BATCH_TRACE=gpurun_out/r02_trace_f16.csv $B --files 16 > gpurun_out/r02_bench_trace_f16.json 2> gpurun_out/r02_bench_trace_f16.err
SDB_BATCH_OPTS=4=0 SDB_BATCH_TRACE=gpurun_out/r02_trace_f16_onecta.csv $B --files 16 > gpurun_out/r02_bench_trace_f16_onecta.json 2>&1
SDB_BATCH_OPTS=4=0 SDB_BATCH_TRACE=gpurun_out/r02_trace_f32_onecta.csv $B --files 32 > gpurun_out/r02_bench_trace_f32_onecta.json 2>&1
for t in ctl_f16 trace_f16 trace_f16_onecta trace_f32_onecta; do
  echo "== $t"; python -c "
import json,sys
for l in open('gpurun_out/r02_bench_$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['path_frac'])
"
done
for t in f16 f16_onecta f32_onecta; do
  echo "== $t"
  python scripts/prof_batch_timeline.py gpurun_out/r02_trace_$t.csv
done
