#!/bin/bash
# round 2, GPU call A: tests, merge-loop interference experiment, size sweep, ncu captures
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_c.log
python scripts/prof_linkage_concurrent.py 1683 4 2>&1 | tee gpurun_out/r02_linkage_concurrent.log
python scripts/prof_cluster_sizes.py 2>&1 | tee gpurun_out/r02_cluster_sizes.log
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time.log
for t in fbank:fbank400 pdist_tc:pdist_tc_kernel aggregate:aggregate_kernel mask_compact:mask_compact_kernel binarize:binarize_kernel; do
  w=${t%%:*}; k=${t##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_$w python scripts/prof_ncu_targets.py $w > gpurun_out/r02_ncu_$w.log 2>&1
  tail -2 gpurun_out/r02_ncu_$w.log
done
ls -la gpurun_out/*.ncu-rep
