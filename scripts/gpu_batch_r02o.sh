#!/bin/bash
# STFT v11 (lane-parallel bulk issue, 80 registers) vs the packed window multiply; e2e with more files in flight
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
echo "--- default" | tee -a gpurun_out/r02_stft_v11.log; python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v11.log
echo "--- SD_WINDOW_MUL2" | tee -a gpurun_out/r02_stft_v11.log; SDB200_LIB=pyannote-audio_speaker-diarization_cpp_b200/variants/lib_mul2.so python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v11.log
done
echo "--- 4 CTAs/SM" | tee -a gpurun_out/r02_stft_v11.log; python scripts/prof_stft.py 4 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v11.log
SDB200_LIB=pyannote-audio_speaker-diarization_cpp_b200/variants/lib_mul2.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stft" 2>&1 | tail -1
for n in 3 5; do
SDB_E2E_FILES=$n python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_e2e$n.json 2>/dev/null
python -c "
import json
for l in open('gpurun_out/r02_bench_e2e$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print('e2e files $n:', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_file'],1), round(d['e2e']['d2h_gbs_achieved'],1))
"
done
