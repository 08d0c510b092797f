#!/bin/bash
# fbank: mel projection by (frame, part) threads vs one filter per thread (variant 7)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py -x -q -m gpu 2>&1 | tail -2
for v in 0 7 0 7; do echo "--- variant $v" | tee -a gpurun_out/r02_fbank_time_v7.log; python scripts/prof_fbank.py 1773 $v 2>&1 | tail -1 | tee -a gpurun_out/r02_fbank_time_v7.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank400_kernel -s 1 -c 1 -f -o gpurun_out/r02_fbank_v4 python scripts/prof_ncu_targets.py fbank > gpurun_out/r02_ncu_fbank_v4.log 2>&1
ls -la gpurun_out/r02_fbank_v4.ncu-rep
