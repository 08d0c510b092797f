#!/bin/bash
# round 2, GPU call B: full tests, default bench, launch list, ncu capture of the STFT kernel, size sweep
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r02_pytest_e.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err
tail -2 gpurun_out/r02_bench_v4.err
python scripts/prof_cluster_sizes.py 2>&1 | tee gpurun_out/r02_cluster_sizes_v2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_v4.csv python bench.py --files 4 --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_under_ncu.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft400_kernel -s 2 -c 1 -f -o gpurun_out/r02_stft_v6 python scripts/prof_stft.py 0 600 > gpurun_out/r02_ncu_stft.log 2>&1
ls -la gpurun_out/r02_stft_v6.ncu-rep gpurun_out/r02_launches_v4.csv
