"""CPU: the C restatement against the reference itself (oracle/_ref/libsdref.so = unmodified reference sources).
Skipped when the reference library has not been built (it needs /root/reference at build time)."""
import numpy as np
import pytest

THRESH = float(np.float32(0.7153814381597874))


@pytest.mark.parametrize("seed,C", [(1, 7), (2, 40)])
def test_segmentation_chain(oracle, ref, synth, seed, C):
    seg = synth.segmentations(seed, C, 293, 3)
    b = oracle.binarize(seg)
    assert np.array_equal(b, ref.binarize(seg))
    assert np.array_equal(oracle.trim(b)[0], ref.trim(b)[0])
    co, cfo = oracle.speaker_count(b)
    cr, cfr = ref.speaker_count(b)
    assert np.array_equal(co, cr) and np.array_equal(cfo, cfr)
    assert np.array_equal(oracle.clean_segmentations(b), ref.clean_segmentations(b))
    sc = seg.astype(np.float64)
    sc[::3, :, 2] = np.nan
    sf = (0.0, 0.5, 5.0, 16000 * 30)
    for skip in (False, True):
        for miss in (0.0, np.nan):
            ao, po = oracle.aggregate(sc, sf, missing=miss, skip_average=skip)
            ar, pr = ref.aggregate(sc, sf, missing=miss, skip_average=skip)
            assert np.array_equal(ao, ar, equal_nan=True) and np.array_equal(po, pr)


def test_aggregate_other_geometry(oracle, ref, synth):
    """10 s / 1 s chunks, 589 frames (cfg2 geometry) and a non-zero window start."""
    # (the reference indexes out of bounds when a chunk has more frames than its duration covers, so the
    # frame counts must be consistent with the window: 589 frames / 10 s, 473 trimmed frames / 8 s)
    for F, sf in ((589, (0.0, 1.0, 10.0, 16000 * 21)), (473, (1.0, 1.0, 8.0, 473))):
        seg = synth.segmentations(5, 12, F, 3).astype(np.float64)
        ao, po = oracle.aggregate(seg, sf, missing=0.0)
        ar, pr = ref.aggregate(seg, sf, missing=0.0)
        assert np.array_equal(ao, ar) and np.array_equal(po, pr)


@pytest.mark.parametrize("N,D,seed", [(2, 4, 0), (3, 192, 1), (50, 16, 2), (257, 192, 3), (700, 192, 4)])
def test_linkage_fcluster(oracle, ref, N, D, seed):
    rng = np.random.default_rng(seed)
    raw = rng.standard_normal((N, D)) * rng.uniform(0.01, 30.0, (N, 1))  # un-normalised, like ECAPA embeddings
    x = oracle.normalize(raw)
    assert np.array_equal(x, ref.normalize(raw))  # Helper::normalizeEmbeddings, float-rounded norm (SD:330-357)
    assert np.array_equal(oracle.pdist(x), ref.pdist(x))
    Zo, Zr = oracle.linkage(x), ref.linkage(x)
    assert np.array_equal(Zo, Zr)
    for cut in (0.5, THRESH, 1.35, 1.45, 5.0):
        assert np.array_equal(oracle.fcluster(Zo, cut), ref.fcluster(Zr, cut))


def test_linkage_with_duplicates_and_grid(oracle, ref):
    rng = np.random.default_rng(11)
    base = rng.standard_normal((20, 6))
    x = np.concatenate([base, base, base[:7]])
    assert np.array_equal(oracle.linkage(x), ref.linkage(x))
    gx, gy = np.meshgrid(np.arange(6.0), np.arange(5.0))
    grid = np.stack([gx.ravel(), gy.ravel()], 1)  # many exactly equal distances
    assert np.array_equal(oracle.linkage(grid), ref.linkage(grid))


@pytest.mark.parametrize("seed,C,nspk,tiny", [(1, 30, 2, ()), (2, 60, 3, (4,)), (3, 109, 4, (3, 5)), (4, 40, 5, (2, 2, 2))])
def test_clustering_stage(oracle, ref, synth, seed, C, nspk, tiny):
    emb, _ = synth.embeddings(seed, C, 3, 192, n_speakers=nspk, tiny=tiny)
    seg = synth.segmentations(seed + 50, C, 293, 3)
    b = oracle.binarize(seg)
    rc_o, ho, _, _ = oracle.clustering_stage(emb, b)
    rc_r, hr = ref.clustering_stage(emb, b)
    assert rc_o == rc_r == 0 and np.array_equal(ho, hr)
    x = emb.reshape(-1, 192)
    x = x[~np.isnan(x[:, 0])]
    assert np.array_equal(oracle.cluster_labels(x)[1], ref.cluster_labels(x)[1])


def test_cosine_cdist(oracle, ref):
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((9, 192)) * 20, rng.standard_normal((4, 192))
    ro, do = oracle.cosine_cdist(a, b)
    rr, dr = ref.cosine_cdist(a, b)
    assert ro == rr == 0 and np.array_equal(do, dr)
    b[2] = 0.0
    assert oracle.cosine_cdist(a, b)[0] == ref.cosine_cdist(a, b)[0] == 2


def test_stft_as_written(oracle, ref, synth):
    import os
    w = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming400_torch.npy"))
    wav = synth.fbank_items(8, 3, 8000)
    out, wl = ref.stft(wav, lens=np.array([1.0, 0.25, 0.5], np.float32))
    assert out.shape == (32, 51, 201, 2) and not out[3:].any()
    assert np.abs(oracle.stft(wav, window=w) - out[:3]).max() < 1e-6
    assert np.array_equal(ref.stft_fft_only(wav), out[:3])
    assert list(wl[:4]) == [1.0, 0.25, 0.5, 1.0]


def test_next_rows(oracle, ref, synth):
    C = 30
    seg = synth.segmentations(21, C, 293, 3)
    b = oracle.binarize(seg)
    count, cf = oracle.speaker_count(b)
    emb, _ = synth.embeddings(22, C, 3, 192, n_speakers=3, tiny=())
    _, hard, _, _ = oracle.clustering_stage(emb, b)
    sf = (0.0, 0.5, 5.0, 16000 * 20)
    ro, fo = oracle.reconstruct(seg, sf, hard, count, cf)
    rr, fr = ref.reconstruct(seg, sf, hard, count, cf)
    assert np.array_equal(ro, rr) and np.array_equal(fo, fr)
    so, lo = oracle.to_annotation(ro, fo)
    sr, lr = ref.to_annotation(rr, fr)
    ko, kr = np.lexsort((lo, so[:, 1], so[:, 0])), np.lexsort((lr, sr[:, 1], sr[:, 0]))
    assert np.array_equal(so[ko], sr[kr]) and np.array_equal(lo[ko], lr[kr])
    wav = synth.fbank_items(4, 3, 80000)
    masks = (synth.segmentations(6, 3, 293, 1)[:, :, 0] > 0.4).astype(np.float32)
    a, b2 = oracle.mask_compact(wav, masks), ref.mask_compact(wav, masks)
    assert a[0] == b2[0] and all(np.array_equal(p, q) for p, q in zip(a[1:], b2[1:]))


def test_ingest_rows(oracle, ref, synth, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "ingest_ref.npz"))
    wav, meta = ref.wav_load(os.path.join(golden_dir, "tiny_list.wav"))
    assert meta == (1, 16, 16000) and np.array_equal(wav, oracle.ingest_pcm16(g["pcm"]))
    wave = synth.waveform(13, 7.3)
    for t, s, nz in zip(g["crop_starts"], g["crop_sum"], g["crop_nonzero"]):
        a, b = oracle.crop(wave, t), ref.crop(wave, t)
        assert a.shape == b.shape == (80000,) and np.array_equal(a, b)
        assert a.astype(np.float64).sum() == s and (a != 0).sum() == nz


@pytest.mark.parametrize("seed", range(6))
def test_next_rows_randomised(oracle, ref, synth, seed):
    """Randomised geometry for the rows either side of the hot path: masks with empty / one-frame / full rows,
    different numbers of chunks and speakers, non-binary scores and all four annotation parameters."""
    rng = np.random.default_rng(100 + seed)
    B, L, F = int(rng.integers(1, 6)), int(rng.integers(1200, 9000)), int(rng.integers(5, 60))
    wav = synth.fbank_items(seed, B, L)
    masks = (rng.random((B, F)) < rng.random()).astype(np.float32)
    if B > 1:
        masks[0] = 0
    if B > 2:
        masks[1] = 1
    a, b2 = oracle.mask_compact(wav, masks), ref.mask_compact(wav, masks)
    assert a[0] == b2[0] and np.array_equal(a[1], b2[1])
    if a[0] == 0:
        assert np.array_equal(a[2], b2[2]) and np.array_equal(a[3], b2[3])
    C = int(rng.integers(12, 40))
    seg = synth.segmentations(300 + seed, C, 293, 3)
    bb = oracle.binarize(seg)
    count, cf = oracle.speaker_count(bb)
    emb, _ = synth.embeddings(400 + seed, C, 3, 64, n_speakers=int(rng.integers(2, 5)), tiny=())
    _, hard, _, _ = oracle.clustering_stage(emb, bb)
    sf = (0.0, 0.5, 5.0, 16000 * (C // 2 + 6))
    ro, fo = oracle.reconstruct(seg, sf, hard, count, cf)
    rr, fr = ref.reconstruct(seg, sf, hard, count, cf)
    assert np.array_equal(ro, rr) and np.array_equal(fo, fr)
    scores = np.clip(ro * 0.7 + rng.uniform(0, 0.45, ro.shape), 0, 1)
    for onset, offset, on, off in ((0.5, 0.5, 0.0, 0.5817029476165771), (0.6, 0.35, 0.1, 0.05), (0.5, 0.5, 0.0, 0.0)):
        so, lo = oracle.to_annotation(scores, fo, onset, offset, on, off)
        sr, lr = ref.to_annotation(scores, fr, onset, offset, on, off)
        ko, kr = np.lexsort((lo, so[:, 1], so[:, 0])), np.lexsort((lr, sr[:, 1], sr[:, 0]))
        assert np.array_equal(so[ko], sr[kr]) and np.array_equal(lo[ko], lr[kr])


@pytest.mark.parametrize("seed", range(4))
def test_clustering_stage_randomised(oracle, ref, synth, seed):
    rng = np.random.default_rng(500 + seed)
    C = int(rng.integers(8, 70))
    emb, _ = synth.embeddings(600 + seed, C, 3, 192, n_speakers=int(rng.integers(1, 6)), nan_frac=float(rng.random() * 0.2),
                              tiny=(2,) if seed % 2 else ())
    seg = synth.segmentations(700 + seed, C, 293, 3)
    b = oracle.binarize(seg)
    rc_o, ho, _, _ = oracle.clustering_stage(emb, b)
    rc_r, hr = ref.clustering_stage(emb, b)
    assert rc_o == rc_r == 0 and np.array_equal(ho, hr)
