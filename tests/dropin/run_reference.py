"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference program (build/dropin_reference: /root/reference's
speakerDiarizer.cpp + onnx_model.cc + clustering.cpp, libtorch STFT, -DWRITE_DATA, deterministic model stand-ins) on
the reference's own test wav and records what it produced:

  tests/golden/dropin_reference.json   SHA-256 + size of every /tmp/cpp_<stage>.txt dump
                                       (pipeline/script/verifyEveryStepResult.py:6-17), the printed speaker segments
                                       (speakerDiarizer.cpp:3437-3440), SHA-256 of each captured emd4.onnx input
  tests/golden/dropin_ort_inputs.npz   a seeded sample of frames of every captured [32,501,201,2] emd4.onnx input
                                       (the whole capture is ~280 MB) + all wav_lens
  oracle/_ref/dropin/                  (git-ignored, travels to the GPU box) the dumps themselves, gzip'ed, and the
                                       full capture as compressed npz, for a complete comparison / diff on failure

~10 CPU-minutes: EmbeddingModel1::infer copies 6.4 M elements per batch through .item<float>() and writes a 68 MB
text file (speakerDiarizer.cpp:2022-2036, 1923-1928).  Run in the container that has /root/reference:

    make -C tests/dropin reference-dumps
"""
import glob
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
WAV = "/root/reference/pipeline/data/multi-speaker_1min.wav"
EXE = os.path.join(HERE, "build", "dropin_reference")
KEEP = os.path.join(ROOT, "oracle", "_ref", "dropin")
GOLD = os.path.join(ROOT, "tests", "golden")
SAMPLE_SEED, SAMPLE_FRAMES = 20261017, 8


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def sample_frames(T):
    """frames checked on the GPU box: both edges (centre padding) + SAMPLE_FRAMES seeded interior frames per row"""
    rng = np.random.default_rng(SAMPLE_SEED)
    return np.unique(np.concatenate([[0, 1, 2, T - 3, T - 2, T - 1], rng.integers(3, T - 3, SAMPLE_FRAMES)]))


def collect(capture_dir, stdout_text, rec):
    """shared with tests/test_gpu_dropin.py: summarise a run (dump hashes, segments, capture hashes)"""
    dumps = {}
    for p in sorted(glob.glob("/tmp/cpp_*.txt")):
        dumps[os.path.basename(p)] = {"sha256": sha_file(p), "bytes": os.path.getsize(p)}
    rec["dumps"] = dumps
    lines = stdout_text.splitlines()
    rec["segments"] = [l for l in lines if l.startswith("[") and "--> Speaker_" in l]
    caps = sorted(glob.glob(os.path.join(capture_dir, "emb_input_*.f32")))
    rec["ort_inputs"] = {os.path.basename(p): sha_file(p) for p in caps}
    return caps


def main():
    for p in glob.glob("/tmp/cpp_*.txt"):
        os.remove(p)
    shutil.rmtree(KEEP, ignore_errors=True)
    cap_dir = os.path.join(KEEP, "capture")
    os.makedirs(cap_dir)
    t0 = time.time()
    r = subprocess.run([EXE, WAV, cap_dir], capture_output=True, text=True)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0, r.returncode
    rec = {"wav": "pipeline/data/multi-speaker_1min.wav", "wav_sha256": sha_file(WAV),
           "seconds": round(time.time() - t0, 1), "sample_seed": SAMPLE_SEED, "sample_frames": SAMPLE_FRAMES}
    caps = collect(cap_dir, r.stdout, rec)
    open(os.path.join(KEEP, "stdout.txt"), "w").write(r.stdout)
    # sample + full capture
    full, lens = [], []
    for p in caps:
        full.append(np.fromfile(p, np.float32).reshape(32, -1, 201, 2))
        lens.append(np.fromfile(p.replace("emb_input", "emb_lens"), np.float32))
    full = np.stack(full)
    frames = sample_frames(full.shape[2])
    np.savez_compressed(os.path.join(GOLD, "dropin_ort_inputs.npz"), frames=frames, sample=full[:, :, frames],
                        wav_lens=np.stack(lens))
    np.savez_compressed(os.path.join(KEEP, "ort_inputs_full.npz"), x=full, wav_lens=np.stack(lens))
    shutil.rmtree(cap_dir)
    for p in sorted(glob.glob("/tmp/cpp_*.txt")):
        with open(p, "rb") as f, gzip.open(os.path.join(KEEP, os.path.basename(p) + ".gz"), "wb", 6) as g:
            shutil.copyfileobj(f, g)
    with open(os.path.join(GOLD, "dropin_reference.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("reference run %.0f s: %d dumps, %d segments, %d emd4 inputs" %
          (rec["seconds"], len(rec["dumps"]), len(rec["segments"]), len(caps)))
    print("\n".join(rec["segments"]))


if __name__ == "__main__":
    main()
