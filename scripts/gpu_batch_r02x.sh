#!/bin/bash
# final state of the round: full GPU suite + the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_i.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_i.log
python bench.py > gpurun_out/r02_bench_v12.json 2> gpurun_out/r02_bench_v12.err
python -c "
import json
for l in open('gpurun_out/r02_bench_v12.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), round(d['single_file']['ms_per_file'],2), d['gpu_launches'])
"
