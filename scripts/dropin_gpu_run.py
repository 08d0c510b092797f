"""Run the libsdb200 drop-in build of the reference pipeline on the GPU box and bring back what it wrote
(development aid; the judged check is tests/test_gpu_dropin.py).  Output: gpurun_out/dropin_b/."""
import glob
import gzip
import importlib.util
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("rr", os.path.join(ROOT, "tests", "dropin", "run_reference.py"))
rr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rr)

out = os.path.join(ROOT, "gpurun_out", "dropin_b")
shutil.rmtree(out, ignore_errors=True)
os.makedirs(out)
for p in glob.glob("/tmp/cpp_*.txt"):
    os.remove(p)
cap = tempfile.mkdtemp()
t0 = time.time()
r = subprocess.run([os.path.join(ROOT, "tests/dropin/build/dropin_sdb200"),
                    os.path.join(ROOT, "oracle/_ref/multi-speaker_1min.wav"), cap], capture_output=True, text=True)
open(os.path.join(out, "stdout.txt"), "w").write(r.stdout)
open(os.path.join(out, "stderr.txt"), "w").write(r.stderr)
rec = {"returncode": r.returncode, "seconds": round(time.time() - t0, 2)}
caps = rr.collect(cap, r.stdout, rec)
json.dump(rec, open(os.path.join(out, "run.json"), "w"), indent=1)
for p in sorted(glob.glob("/tmp/cpp_*.txt")):
    if os.path.getsize(p) < 4 << 20:
        with open(p, "rb") as f, gzip.open(os.path.join(out, os.path.basename(p) + ".gz"), "wb") as g:
            shutil.copyfileobj(f, g)
if caps:
    frames = rr.sample_frames(501)
    full = np.stack([np.fromfile(p, np.float32).reshape(32, 501, 201, 2) for p in caps])
    lens = np.stack([np.fromfile(p.replace("emb_input", "emb_lens"), np.float32) for p in caps])
    np.savez_compressed(os.path.join(out, "ort_inputs_sample.npz"), frames=frames, sample=full[:, :, frames], wav_lens=lens)
print("rc", r.returncode, "seconds", rec["seconds"], "dumps", len(rec["dumps"]), "segments", len(rec["segments"]))
print("\n".join(rec["segments"]))
sys.stderr.write(r.stderr[-1500:])
