"""TEST INFRASTRUCTURE -- regenerates tests/golden/*.npz by running the reference itself (oracle/_ref/libsdref.so,
the unmodified reference sources) on seeded inputs, plus the reference's own golden file
pipeline/src/test/closest_frame.txt and torch's window table.  Run in a container that has /root/reference:

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT  # replace the script directory (it would shadow the `oracle` package)
import __graft_entry__ as ge  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference"


def main():
    os.makedirs(OUT, exist_ok=True)
    synth = ge.load_synth()
    r = Ref()

    # 1. the reference's own fixture: SlidingWindow(0, .016875, .016875).closest_frame(0.5*k)
    rows = [l.strip().split(",") for l in open(os.path.join(REF_ROOT, "pipeline/src/test/closest_frame.txt"))]
    frames = np.array([int(a) for a, _ in rows], np.int32)
    np.savez_compressed(os.path.join(OUT, "closest_frame.npz"), frames=frames)

    # 2. torch::hamming_window(400) as the reference gets it (speakerDiarizer.cpp:2007)
    import torch
    np.save(os.path.join(OUT, "hamming400_torch.npy"), torch.hamming_window(400).numpy())

    # 3. STFT: EmbeddingModel1::infer as written, ORT input captured
    wav = synth.fbank_items(3, 2, 16000)
    wav[1, 9000:] = 0.0
    out, wl = r.stft(wav, lens=np.array([1.0, 0.5625], np.float32))
    np.savez_compressed(os.path.join(OUT, "stft_ref.npz"), wav=wav, out=out[:2], pad_is_zero=not out[2:].any(),
                        wav_lens=wl)

    # 4. segmentation post-processing at the reference constants (cfg1 shape)
    C, F, K = 109, 293, 3
    seg = synth.segmentations(101, C, F, K)
    b = r.binarize(seg)
    trimmed, tw = r.trim(b)
    count, cf = r.speaker_count(b)
    clean = r.clean_segmentations(b)
    sc = seg.astype(np.float64)
    sc[::5, :, 1] = np.nan
    sc[10:20] = np.nan
    sf = (0.0, 0.5, 5.0, 944000)
    agg_avg, post = r.aggregate(sc, sf, missing=np.nan, skip_average=False)
    agg_sum, _ = r.aggregate(sc, sf, missing=0.0, skip_average=True)
    s64 = seg.transpose(0, 2, 1).reshape(C * K, F).astype(np.float64)[:64].copy()
    s64[::7, 5:9] = 0.5
    s64[3, :4] = 0.5
    rows_bin = r.binarize_ndarray(s64, 0.5, False)
    rows_bin_init = r.binarize_ndarray(s64, 0.5, True)
    np.savez_compressed(os.path.join(OUT, "segpost_ref.npz"), seg=seg, binarized=b.astype(np.uint8), tw=tw, count=count,
                        cf=cf, clean=clean.astype(np.uint8), sc_nan_mask=np.isnan(sc), agg_avg=agg_avg, agg_sum=agg_sum,
                        post=post, rows_scores=s64, rows_bin=rows_bin, rows_bin_init=rows_bin_init,
                        trimmed_shape=np.array(trimmed.shape))

    # 5. clustering
    emb, spk = synth.embeddings(201, C, 3, 192)
    rc, hard = r.clustering_stage(emb, b)
    assert rc == 0
    x = emb.reshape(-1, 192)
    x = x[~np.isnan(x[:, 0])]
    xn = r.normalize(x)
    Z = r.linkage(xn)
    T = r.fcluster(Z, float(np.float32(0.7153814381597874)))
    rc, labels = r.cluster_labels(x)
    assert rc == 0
    np.savez_compressed(os.path.join(OUT, "cluster_ref.npz"), emb=emb.astype(np.float32), hard=hard, Z=Z, T=T,
                        labels=labels, xn_first=xn[:4])

    # 6. toy of pipeline/src/clustering/cluster.cpp:8-13 and a duplicate-rows tie case
    toy = np.array([[0, 0], [0, 1], [1, 0], [0, 4], [0, 3], [1, 4], [4, 0], [3, 0], [4, 1], [4, 4], [3, 4], [4, 3]],
                   np.float64)
    rng = np.random.default_rng(7)
    base = rng.standard_normal((12, 8))
    ties = np.concatenate([base, base[:6], base[3:9], np.zeros((2, 8)) + 0.25])
    np.savez_compressed(os.path.join(OUT, "linkage_small.npz"), toy=toy, toy_Z=r.linkage(toy),
                        toy_T=r.clustering_cluster(toy, 1.1), ties=ties, ties_Z=r.linkage(ties),
                        ties_T=r.clustering_cluster(ties, 1.0))

    # 7. next rows: mask/compaction, reconstruct, to_annotation
    wav4 = synth.fbank_items(5, 4, 80000)
    masks = (synth.segmentations(9, 4, 293, 1)[:, :, 0] > 0.5).astype(np.float32)
    masks[2] = 0
    masks[2, :2] = 1
    rc, sig, lens, ts = r.mask_compact(wav4, masks)
    rec, fr = r.reconstruct(seg, sf, hard, count, cf)
    segs, labs = r.to_annotation(rec, fr)
    np.savez_compressed(os.path.join(OUT, "next_ref.npz"), masks=masks, mc_rc=rc, sig_nonzero=(sig != 0).sum(1),
                        sig_sum=sig.astype(np.float64).sum(1), lens=lens, too_short=ts, rec=rec.astype(np.uint8), fr=fr,
                        segs=segs, labs=labs)
    # 8. ingest: a PCM16 mono WAV with a LIST sub-chunk between "fmt " and "data" (wav.h:84-91), read by WavReader
    import struct
    pcm = (synth.waveform(12, 0.25)[:4000] * 20000).astype(np.int16)
    pcm[:4] = [-32768, 32767, 0, -1]
    lst = b"INFOISFT" + struct.pack("<I", 6) + b"synth\0"
    body = (b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + b"LIST" +
            struct.pack("<I", len(lst)) + lst + b"data" + struct.pack("<I", pcm.size * 2) + pcm.tobytes())
    with open(os.path.join(OUT, "tiny_list.wav"), "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    wav_f, meta = r.wav_load(os.path.join(OUT, "tiny_list.wav"))
    assert wav_f.size == 4000 and meta == (1, 16, 16000)
    long_wave = synth.waveform(13, 7.3)
    crop_starts = np.array([0.0, 0.5, 2.30001, 3.0, 6.9, 7.29])
    crops = np.stack([r.crop(long_wave, t) for t in crop_starts])
    np.savez_compressed(os.path.join(OUT, "ingest_ref.npz"), pcm=pcm, wav=wav_f, meta=np.array(meta),
                        crop_starts=crop_starts, crop_sum=crops.astype(np.float64).sum(1),
                        crop_nonzero=(crops != 0).sum(1), crop_first=crops[:, :4], crop_last=crops[:, -4:])
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden written to", OUT, "%.1f KB" % (tot / 1024))


if __name__ == "__main__":
    main()
