"""CPU: the C restatement (oracle/sd_oracle.c) against the committed golden vectors.

Fixtures were produced by the reference itself (oracle/make_golden.py, oracle/_ref) plus the reference's own
golden file pipeline/src/test/closest_frame.txt.  Integer / index / fp64 results must be bit-identical."""
import os

import numpy as np

THRESH = float(np.float32(0.7153814381597874))


def g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_closest_frame_reference_fixture(oracle, golden_dir):
    frames = g(golden_dir, "closest_frame.npz")["frames"]
    assert len(frames) == 10000
    t = 0.0
    for k in range(len(frames)):
        assert oracle.closest_frame(t) == frames[k], k
        t += 0.5  # the reference accumulates (speakerDiarizer.cpp:3258)
    assert frames[1] == 29 and frames[2] == 59 and frames[-1] == 296266


def test_np_rint_half_even(oracle):
    assert [oracle.np_rint(v) for v in (0.5, 1.5, 2.5, -1.5, 3.5, -2.5, 1.2, 3.6)] == [0, 2, 2, -2, 4, -2, 1, 4]


def test_hamming_window_close_to_torch(oracle, golden_dir):
    w = np.load(os.path.join(golden_dir, "hamming400_torch.npy"))
    mine = oracle.hamming_window(400)
    assert np.abs(mine - w).max() <= 6e-8  # a few entries differ by one ulp (vectorised cos in ATen)
    assert abs(float(w[0]) - 0.08000001311) < 1e-9 and w[200] == 1.0


def test_stft_vs_reference_capture(oracle, golden_dir):
    d = g(golden_dir, "stft_ref.npz")
    w = np.load(os.path.join(golden_dir, "hamming400_torch.npy"))
    got = oracle.stft(d["wav"], window=w)
    assert got.shape == d["out"].shape == (2, 101, 201, 2)
    assert np.abs(got - d["out"]).max() < 1e-6  # fp64 DFT vs fp64 FFT, both rounded to fp32
    assert bool(d["pad_is_zero"])
    assert np.allclose(d["wav_lens"][:2], [1.0, 0.5625]) and np.all(d["wav_lens"][2:] == 1.0)
    # spot frames helper agrees with the full transform
    fr = oracle.stft_frames(d["wav"][1], 40, 45, window=w)
    assert np.array_equal(fr, got[1, 40:45])


def test_segmentation_postprocessing(oracle, golden_dir):
    d = g(golden_dir, "segpost_ref.npz")
    seg = d["seg"]
    b = oracle.binarize(seg)
    assert np.array_equal(b, d["binarized"].astype(np.float64))
    tr, tw = oracle.trim(b)
    assert tuple(tr.shape) == tuple(d["trimmed_shape"]) and np.array_equal(tw, d["tw"])
    cnt, cf = oracle.speaker_count(b)
    assert np.array_equal(cnt, d["count"]) and np.array_equal(cf, d["cf"])
    assert np.array_equal(oracle.clean_segmentations(b), d["clean"].astype(np.float64))
    sc = seg.astype(np.float64)
    sc[d["sc_nan_mask"]] = np.nan
    sf = (0.0, 0.5, 5.0, 944000)
    a1, post = oracle.aggregate(sc, sf, missing=np.nan, skip_average=False)
    a2, _ = oracle.aggregate(sc, sf, missing=0.0, skip_average=True)
    assert np.array_equal(a1, d["agg_avg"], equal_nan=True)
    assert np.array_equal(a2, d["agg_sum"])
    assert np.array_equal(post, d["post"])


def test_binarize_rows_tie_handling(oracle, golden_dir):
    """binarize_ndarray with scores exactly equal to onset (undefined frames carry the previous state)."""
    d = g(golden_dir, "segpost_ref.npz")
    s = d["rows_scores"]
    R, F = s.shape
    # the oracle's entry takes [C][F][K] float32; exact-0.5 values survive the float round trip
    got = oracle.binarize(s.astype(np.float32).reshape(R, F, 1), onset=0.5, initial_state=False)[:, :, 0]
    assert np.array_equal(got.astype(np.uint8), d["rows_bin"])
    got = oracle.binarize(s.astype(np.float32).reshape(R, F, 1), onset=0.5, initial_state=True)[:, :, 0]
    assert np.array_equal(got.astype(np.uint8), d["rows_bin_init"])


def test_linkage_toy_and_ties(oracle, golden_dir):
    d = g(golden_dir, "linkage_small.npz")
    Z = oracle.linkage(d["toy"])
    assert np.array_equal(Z, d["toy_Z"])
    assert list(oracle.fcluster(Z, 1.1)) == [5, 5, 6, 7, 7, 8, 1, 1, 2, 3, 3, 4] == list(d["toy_T"])
    Zt = oracle.linkage(d["ties"])
    assert np.array_equal(Zt, d["ties_Z"])  # heap-order tie breaking
    assert np.array_equal(oracle.fcluster(Zt, 1.0), d["ties_T"])


def test_linkage_matches_scipy(oracle):
    from scipy.cluster.hierarchy import fcluster, linkage
    rng = np.random.default_rng(3)
    x = rng.standard_normal((400, 64))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    Z = oracle.linkage(x)
    Zs = linkage(x, method="centroid")
    assert np.array_equal(Z, Zs)
    for cut in (0.9, 1.2, THRESH):
        assert np.array_equal(oracle.fcluster(Z, cut), fcluster(Zs, cut, criterion="distance"))


def test_clustering_stage(oracle, golden_dir):
    d = g(golden_dir, "cluster_ref.npz")
    s = g(golden_dir, "segpost_ref.npz")
    emb = d["emb"].astype(np.float64)
    x = emb.reshape(-1, 192)
    x = x[~np.isnan(x[:, 0])]
    xn = oracle.normalize(x)
    assert np.array_equal(xn[:4], d["xn_first"])
    Z = oracle.linkage(xn)
    assert np.array_equal(Z, d["Z"])
    assert np.array_equal(oracle.fcluster(Z, THRESH), d["T"])
    rc, lab = oracle.cluster_labels(x)
    assert rc == 0 and np.array_equal(lab, d["labels"])
    rc, hard, _, k = oracle.clustering_stage(emb, s["binarized"].astype(np.float64))
    assert rc == 0 and np.array_equal(hard, d["hard"]) and k == d["labels"].max() + 1


def test_next_rows(oracle, synth, golden_dir):
    d = g(golden_dir, "next_ref.npz")
    s = g(golden_dir, "segpost_ref.npz")
    c = g(golden_dir, "cluster_ref.npz")
    wav4 = synth.fbank_items(5, 4, 80000)
    rc, sig, lens, ts = oracle.mask_compact(wav4, d["masks"])
    assert rc == d["mc_rc"] and np.array_equal(lens, d["lens"]) and np.array_equal(ts, d["too_short"])
    assert np.array_equal((sig != 0).sum(1), d["sig_nonzero"])
    assert np.array_equal(sig.astype(np.float64).sum(1), d["sig_sum"])
    rec, fr = oracle.reconstruct(s["seg"], (0.0, 0.5, 5.0, 944000), c["hard"], s["count"], s["cf"])
    assert np.array_equal(rec, d["rec"].astype(np.float64)) and np.array_equal(fr, d["fr"])
    segs, labs = oracle.to_annotation(rec, fr)
    key = np.lexsort((labs, segs[:, 1], segs[:, 0]))
    keyr = np.lexsort((d["labs"], d["segs"][:, 1], d["segs"][:, 0]))
    assert np.array_equal(segs[key], d["segs"][keyr]) and np.array_equal(labs[key], d["labs"][keyr])


def test_cosine_zero_magnitude_is_an_error(oracle):
    a = np.ones((2, 8))
    b = np.zeros((1, 8))
    rc, _ = oracle.cosine_cdist(a, b)
    assert rc == 2


def test_ingest_golden(oracle, golden_dir):
    """f4: WavReader + /32768 scaling and SegmentModel::crop, vectors generated by the reference itself."""
    import os
    g = np.load(os.path.join(golden_dir, "ingest_ref.npz"))
    assert np.array_equal(oracle.ingest_pcm16(g["pcm"]), g["wav"])
    raw = open(os.path.join(golden_dir, "tiny_list.wav"), "rb").read()
    at = raw.index(b"data") + 8
    assert np.array_equal(np.frombuffer(raw[at:at + 8000], np.int16), g["pcm"])
    assert oracle.slide_geometry(16000 * 60) == (110, 880000, 80000)
    assert oracle.slide_geometry(80000) == (0, 0, 80000)
    assert oracle.slide_geometry(80001) == (1, 8000, 72001)
    assert oracle.slide_geometry(1) == (0, -1, 0)
