#!/bin/bash
# final evidence of the round: full GPU suite, default bench (complete line), cfg4 workload, ncu of the final STFT kernel
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_h.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_h.log
python bench.py > gpurun_out/r02_bench_v11.json 2> gpurun_out/r02_bench_v11.err
python bench.py --workload cfg4 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_cfg4_1gpu_v2.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for t in v11 cfg4_1gpu_v2; do python -c "
import json
for l in open('gpurun_out/r02_bench_$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$t', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), round(d['single_file']['ms_per_file'],2))
"; done
tail -c 600 gpurun_out/r02_bench_reference_arm.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft400_kernel -s 2 -c 1 -f -o gpurun_out/r02_stft_v12 python scripts/prof_stft.py 0 1773 > gpurun_out/r02_ncu_stft_v12.log 2>&1
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time_v6.log
