#!/bin/bash
# STFT v7 / fbank v3: real-pair split between adjacent lanes (shuffle), two block barriers per tile
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
python scripts/prof_stft.py 0 1773 2>&1 | tail -2 | tee gpurun_out/r02_stft_v7_shfl.log
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time_v3.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_v8_quick.json 2> gpurun_out/r02_bench_v8_quick.err
python -c "
import json
for l in open('gpurun_out/r02_bench_v8_quick.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft', round(d['single_file']['stages_ms']['stft'],4), round(d['roofline']['frac'],4))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft400_kernel -s 2 -c 1 -f -o gpurun_out/r02_stft_v7 python scripts/prof_stft.py 0 600 > gpurun_out/r02_ncu_stft_v7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank400_kernel -s 1 -c 1 -f -o gpurun_out/r02_fbank_v3 python scripts/prof_ncu_targets.py fbank > gpurun_out/r02_ncu_fbank_v3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
