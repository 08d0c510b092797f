"""CPU: the C-ABI library loads, exports every symbol include/sdb200.h declares, refuses to run without a GPU
(no CPU fallback), and its host-only helpers agree with the reference's golden file.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = header_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(pkg.EXPORTS) == names
    assert L.sd_version() == 100


def test_library_is_sm100a_and_torch_free(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "oracle" not in ldd and "sdref" not in ldd


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.SdError) as e:
        pkg.Context(0)
    assert e.value.code == pkg.SD_ERR_CUDA


def test_host_helpers(pkg, golden_dir):
    L = pkg.lib()
    frames = np.load(os.path.join(golden_dir, "closest_frame.npz"))["frames"]
    w = pkg.Window(0.0, 0.016875, 0.016875, 0)
    t = 0.0
    for k in range(0, 10000):
        assert L.sd_closest_frame(C.byref(w), t) == frames[k]
        t += 0.5
    assert [L.sd_np_rint(v) for v in (0.5, 1.5, 2.5, -1.5, 3.5)] == [0, 2, 2, -2, 4]
    assert L.sd_trim_num_frames(293, 0.1, 0.1) == 235 and L.sd_trim_num_frames(589, 0.1, 0.1) == 473
    assert L.sd_stft_num_frames(80000, 160) == 501 and L.sd_stft_num_frames(160000, 160) == 1001
    cw = pkg.Window(0.0, 0.5, 5.0, 944000)
    assert L.sd_aggregate_num_frames(109, C.byref(cw), C.byref(w)) == 3497
    tw = pkg.Window(0.5, 0.5, 4.0, 235)
    assert L.sd_aggregate_num_frames(109, C.byref(tw), C.byref(w)) == 3438
    lens = np.array([0.5, 0.25], np.float32)
    out = np.zeros(32, np.float32)
    assert L.sd_pack_wav_lens(lens.ctypes.data_as(pkg.c_fp), 2, 32, out.ctypes.data_as(pkg.c_fp)) == 0
    assert list(out[:3]) == [0.5, 0.25, 1.0] and np.all(out[2:] == 1.0)
    p = pkg.ClusterParams()
    L.sd_cluster_default_params(C.byref(p))
    assert p.threshold == np.float32(0.7153814381597874) and p.min_cluster_size == 15 and p.num_clusters == -1


def test_stft_phase_emulation_on_cpu(tmp_path):
    """Replays the three kernel phases of csrc/fft400.cuh thread by thread on the CPU (index maps, exchange
    layouts, real-pair split) against a direct fp64 DFT."""
    exe = str(tmp_path / "emulate_stft")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "emulate_stft.cpp"),
                           "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_mel_projection_emulation_on_cpu(tmp_path):
    """Replays the fused fbank kernel's mel projection (csrc/mel_table.h: (frame, 20-bin part) threads, even / odd
    accumulators, flush columns) on the CPU against the plain per-filter sums, for the speechbrain and Kaldi banks."""
    exe = str(tmp_path / "emulate_mel")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "emulate_mel.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "speechbrain 80" in out.stdout and "UNEXPECTED" not in out.stdout


def test_batch_has_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.SdError) as e:
        pkg.Batch(0, 2)
    assert e.value.code == pkg.SD_ERR_CUDA


def test_batch_timeline_script_reads_a_committed_trace():
    """scripts/prof_batch_timeline.py on a trace written by SDB_BATCH_TRACE on a B200 (profiles/): the analysis the
    batch scheduling of DESIGN.md section 5 was derived from must keep running."""
    trace = os.path.join(ROOT, "profiles", "r02_trace_v7_f24.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_batch_timeline.py"), trace],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "workers 24" in out.stdout and "merge loop" in out.stdout and "STFT completions" in out.stdout
