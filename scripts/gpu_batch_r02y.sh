#!/bin/bash
# 2 GPUs of one box: the bench as the driver launches it (final code of the round)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v12_2gpu.json 2> gpurun_out/r02_bench_v12_2gpu.err
python -c "
import json
for l in open('gpurun_out/r02_bench_v12_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step'],2), d['n_gpus'], 'e2e', round(d['e2e']['value']), d['config']['host_wait'])
"
