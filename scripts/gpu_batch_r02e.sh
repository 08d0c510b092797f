#!/bin/bash
# dynamic STFT tile scheduling: parity tests, then chain depth x files in flight
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs"
run() {  # name, files, env...
  name=$1; files=$2; shift 2
  env "$@" SDB_BATCH_TRACE=gpurun_out/r02_trace_$name.csv $B --files $files > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err
  python -c "
import json,sys
for l in open('gpurun_out/r02_bench_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['path_frac'],3), 'stft alone', round(d['single_file']['stages_ms']['stft'],4))
"
}
run dyn_chain0_f16 16 SDB_BATCH_STFT_CHAIN=0
run dyn_chain1_f16 16 SDB_BATCH_STFT_CHAIN=1
run dyn_chain2_f16 16 SDB_BATCH_STFT_CHAIN=2
run dyn_chain0_f24 24 SDB_BATCH_STFT_CHAIN=0
run dyn_chain1_f24 24 SDB_BATCH_STFT_CHAIN=1
run dyn_chain2_f24 24 SDB_BATCH_STFT_CHAIN=2
run dyn_chain1_f32 32 SDB_BATCH_STFT_CHAIN=1
run dyn_chain1_f24_onecta 24 SDB_BATCH_STFT_CHAIN=1 SDB_BATCH_OPTS=4=0
