#!/bin/bash
# STFT v8 (branch-free stores, indexed shuffle, tile coordinates from thread 0) vs the early-barrier variant
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kaldi.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
echo "--- default (barrier at the end of the tile)" | tee gpurun_out/r02_stft_v8.log
python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v8.log
echo "--- SD_STFT_BARRIER_EARLY" | tee -a gpurun_out/r02_stft_v8.log
SDB200_LIB=pyannote-audio_speaker-diarization_cpp_b200/variants/lib_early.so python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v8.log
SDB200_LIB=pyannote-audio_speaker-diarization_cpp_b200/variants/lib_early.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stft" 2>&1 | tail -2
python scripts/prof_fbank.py 2>&1 | tee gpurun_out/r02_fbank_time_v4.log
echo "--- again" | tee -a gpurun_out/r02_stft_v8.log
python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v8.log
SDB200_LIB=pyannote-audio_speaker-diarization_cpp_b200/variants/lib_early.so python scripts/prof_stft.py 0 1773 2>&1 | tail -1 | tee -a gpurun_out/r02_stft_v8.log
