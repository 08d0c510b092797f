#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py -x -q -m gpu 2>&1 | tail -2
for v in 0 8; do echo "--- variant $v" | tee -a gpurun_out/r02_fbank_time_v8.log; python scripts/prof_fbank.py 1773 $v 2>&1 | tail -1 | tee -a gpurun_out/r02_fbank_time_v8.log; done
