"""CPU: the numpy restatement of the Kaldi-compatible front-end (oracle.kaldi_*) against
torchaudio.compliance.kaldi.fbank -- frozen in tests/golden/kaldi_fbank.npz by oracle/make_golden_kaldi.py, and
re-computed live when torchaudio is importable.  (BASELINE north_star bullet 1; not on the reference's own path.)"""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaldi_fbank.npz")


def cases(g):
    for k in g.files:
        if k.startswith("fbank_"):
            yield k, bool(int(k[10])), bool(int(k[14])), int(k.split("pre")[1]) / 100.0


def test_restatement_matches_torchaudio_golden():
    g = np.load(GOLD)
    wav = g["wav"]
    n = 0
    for k, snip, dc, pre in cases(g):
        for b in range(wav.shape[0]):
            want = g[k][b]
            got = O.kaldi_fbank(wav[b], 80, snip, pre, dc)
            assert got.shape == want.shape
            err = np.abs(got - want)
            loud = want > -8.0  # mel energies above 3e-4: fp32 (torchaudio) vs fp64 agree to 1e-4 in the log domain
            assert err[loud].max() < 1e-4
            assert err.max() < 2e-3  # near-silent bins carry torchaudio's fp32 FFT noise
            n += 1
    assert n == 24


def test_frame_counts_and_reflection():
    x = np.arange(1000, dtype=np.float64)
    assert O.kaldi_frames(x, snip_edges=True).shape == (1 + (1000 - 400) // 160, 400)
    assert O.kaldi_frames(x, snip_edges=False).shape == ((1000 + 80) // 160, 400)
    # first frame of snip_edges=False starts 120 samples before the signal, mirrored: x[119], ..., x[0], x[0], ...
    f = O.kaldi_frames(x, snip_edges=False, preemph=0.0, remove_dc_offset=False)
    w = O.kaldi_frames(np.ones(1000), snip_edges=True, preemph=0.0, remove_dc_offset=False)[0]
    assert np.allclose(f[0, 1:200] / w[1:200], np.concatenate([np.arange(119, -1, -1), np.arange(0, 280)])[1:200])


def test_live_against_torchaudio():
    torch = pytest.importorskip("torch")
    K = pytest.importorskip("torchaudio.compliance.kaldi")
    rng = np.random.default_rng(5)
    x = (0.3 * rng.standard_normal(16000)).astype(np.float32)
    want = K.fbank(torch.from_numpy(x)[None], num_mel_bins=80, dither=0.0, energy_floor=0.0, window_type="povey",
                   round_to_power_of_two=False, snip_edges=False).numpy()
    got = O.kaldi_fbank(x, 80, snip_edges=False)
    assert np.abs(got - want).max() < 1e-4
