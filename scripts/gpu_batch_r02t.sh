#!/bin/bash
# fused speaker_count kernel: parity + drop-in + smoke; quick bench
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_batch.py tests/test_gpu_host_shim.py -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_v12_quick.json 2>/dev/null
python -c "
import json
for l in open('gpurun_out/r02_bench_v12_quick.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stages', d['single_file']['stages_ms'], 'launches', d['gpu_launches'])
"
