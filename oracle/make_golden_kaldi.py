"""TEST INFRASTRUCTURE -- tests/golden/kaldi_fbank.npz: torchaudio.compliance.kaldi.fbank (the Kaldi-compatible
front-end named by BASELINE.json's north_star) on seeded signals, for every framing / conditioning combination the
library offers.  The reference's own front-end is torch::stft + Hamming (speakerDiarizer.cpp:2007-2008); its only
Kaldi-style parameter set is embeddings/threeModel.py:7-66 (25 ms / 10 ms / n_fft 400), which is what is used here
(round_to_power_of_two=False keeps the 400-point transform).

    python oracle/make_golden_kaldi.py
"""
import os
import sys

import numpy as np
import torch
import torchaudio.compliance.kaldi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
import __graft_entry__ as ge  # noqa: E402

CASES = [(snip, dc, pre) for snip in (False, True) for dc in (False, True) for pre in (0.97, 0.0)]


def main():
    synth = ge.load_synth()
    wav = synth.fbank_items(77, 3, 24000)
    wav[0] = synth.waveform(78, 1.5)[:24000]          # speech-like harmonic stacks with pauses
    wav[2] += 0.05                                     # a DC offset, so that remove_dc_offset matters
    out = {"wav": wav}
    for snip, dc, pre in CASES:
        feats = [K.fbank(torch.from_numpy(w)[None], num_mel_bins=80, frame_length=25.0, frame_shift=10.0, dither=0.0,
                         energy_floor=0.0, preemphasis_coefficient=pre, remove_dc_offset=dc, window_type="povey",
                         round_to_power_of_two=False, snip_edges=snip, sample_frequency=16000.0, low_freq=20.0,
                         high_freq=0.0, use_energy=False).numpy() for w in wav]
        out["fbank_snip%d_dc%d_pre%d" % (snip, dc, int(pre * 100))] = np.stack(feats)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kaldi_fbank.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
