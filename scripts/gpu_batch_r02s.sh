#!/bin/bash
# compute-sanitizer over the restructured front-end kernels (dynamic tiles, pair-lane exchange, lane-parallel bulk issue)
cd "$GRAFT_REPO_ROOT" || exit 1
for tool in racecheck synccheck memcheck; do
  SDB_SANITIZE_ONLY=frontend timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02_sanitizer_frontend_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "^(ok|FAIL)|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r02_sanitizer_frontend_$tool.log | sort | uniq -c | sort -rn | head -12
done
