#!/bin/bash
# STFT v12: two sample buffers (variant 0) vs one (variant 4) vs five CTAs per SM (variant 3)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in 0 4 3 0 4; do echo "--- variant $v" | tee -a gpurun_out/r02_stft_v12.log; python scripts/prof_stft.py $v 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v12.log; done
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_v10_quick.json 2>/dev/null
python -c "
import json
for l in open('gpurun_out/r02_bench_v10_quick.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4))
"
