#!/bin/bash
# STFT v13: pads 0 / 20, one bulk copy per pair of hop segments (9 instead of 18 per tile)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in 0 3 0; do echo "--- variant $v" | tee -a gpurun_out/r02_stft_v13.log; python scripts/prof_stft.py $v 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v13.log; done
python scripts/prof_fbank.py 2>&1 | tail -1 | tee gpurun_out/r02_fbank_time_v9.log
