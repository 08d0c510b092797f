#!/bin/bash
# STFT variants (0 = 5 CTAs/SM default, 4 = 4 CTAs/SM, 5 = 8-frame tiles), full GPU suite, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
for v in 0 4 5 0 5; do echo "--- variant $v" | tee -a gpurun_out/r02_stft_v10.log; python scripts/prof_stft.py $v 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v10.log; done
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_g.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_g.log
python bench.py > gpurun_out/r02_bench_v9.json 2> gpurun_out/r02_bench_v9.err
python -c "
import json
for l in open('gpurun_out/r02_bench_v9.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['single_file']['ms_per_file'])
"
