#!/bin/bash
# STFT v9: compact twiddle table; variant 3 = five CTAs per SM (72 registers, 24 B spills)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in 0 3 0 3; do echo "--- variant $v" | tee -a gpurun_out/r02_stft_v9.log; python scripts/prof_stft.py $v 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v9.log; done
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time_v5.log
