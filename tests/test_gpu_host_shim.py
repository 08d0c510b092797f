"""GPU: the C++ host shim (host/sdb200_host.hpp) against the reference translation unit compiled next to it.
The binary is built by __graft_entry__.build() in the container that has /root/reference and travels with the
snapshot (tests/cpp/build/test_host_shim)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "test_host_shim")


def test_host_shim_matches_reference_functions():
    if not os.path.exists(EXE):
        pytest.skip("tests/cpp/build/test_host_shim not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, os.path.join(ROOT, "tests", "golden", "tiny_list.wav")], capture_output=True, text=True, timeout=600)
    print(r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "PASSED" in r.stderr
    for name in ("reconstruct", "to_annotation", "masked_signals", "read_wav", "crop (SegmentModel::crop)"):
        assert "ok   " + name in r.stderr, name
