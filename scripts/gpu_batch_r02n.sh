#!/bin/bash
# evidence for the final kernels: ncu --set full of the STFT at the bench size, launch list of the bench command
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft400_kernel -s 2 -c 1 -f -o gpurun_out/r02_stft_v9 python scripts/prof_stft.py 0 1773 > gpurun_out/r02_ncu_stft_v9.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_v9.csv python bench.py --files 12 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_under_ncu_v9.json 2>/dev/null
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants" 2>&1 | tail -2
ls -la gpurun_out/r02_stft_v9.ncu-rep gpurun_out/r02_launches_v9.csv
