"""SURVEY 8f "next" rows on the device, bit-exact against the oracle and the reference-generated golden vectors."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _masks(synth, seed, B, F, thr=0.5):
    return (synth.segmentations(seed, B, F, 1)[:, :, 0] > thr).astype(np.float32)


@pytest.mark.parametrize("B,L,F,seed", [(4, 80000, 293, 3), (32, 160000, 589, 5), (3, 4001, 17, 7), (1, 1000, 999, 9)])
def test_mask_compact_matches_oracle(ctx, oracle, synth, B, L, F, seed):
    wav = synth.fbank_items(seed, B, L)
    masks = _masks(synth, seed + 1, B, F)
    if B > 2:
        masks[1] = 0            # empty item -> too short
        masks[2] = 0
        masks[2, :1] = 1        # one frame only
    rc_o, sig_o, lens_o, ts_o = oracle.mask_compact(wav, masks)
    rc_g, sig_g, lens_g, ts_g = ctx.mask_compact(wav, masks)
    assert rc_o == rc_g
    assert np.array_equal(sig_o, sig_g)
    if rc_o == 0:
        assert np.array_equal(lens_o, lens_g) and np.array_equal(ts_o, ts_g)


def test_mask_compact_all_too_short(ctx, oracle, synth):
    wav = synth.fbank_items(2, 3, 80000)
    masks = np.zeros((3, 293), np.float32)
    masks[:, 5] = 1
    rc_o = oracle.mask_compact(wav, masks)[0]
    rc_g, sig_g, _, _ = ctx.mask_compact(wav, masks)
    assert rc_o == 1 and rc_g == 1
    assert np.array_equal(sig_g, oracle.mask_compact(wav, masks)[1])


def test_mask_compact_golden(ctx, synth, golden_dir):
    g = np.load(os.path.join(golden_dir, "next_ref.npz"))
    wav4 = synth.fbank_items(5, 4, 80000)
    rc, sig, lens, ts = ctx.mask_compact(wav4, g["masks"])
    assert rc == int(g["mc_rc"])
    assert np.array_equal((sig != 0).sum(1), g["sig_nonzero"])
    assert np.array_equal(sig.astype(np.float64).sum(1), g["sig_sum"])
    assert np.array_equal(lens, g["lens"]) and np.array_equal(ts, g["too_short"])


def test_select_masks(ctx, synth):
    C_, F, K = 21, 589, 3
    b = (synth.segmentations(11, C_, F, K) > 0.5).astype(np.float64)
    b[3, :, 1] = 0
    min_num_frames = 3.0
    clean = b * (b.sum(2, keepdims=True) < 2)
    want = np.empty((C_ * K, F), np.float32)
    for c in range(C_):
        for k in range(K):
            use = clean[c, :, k].astype(np.float32).sum() > min_num_frames
            want[c * K + k] = (clean if use else b)[c, :, k]
    got = ctx.select_masks(b, min_num_frames)
    assert np.array_equal(got, want)


def test_mask_compact_file_equals_batches(ctx, oracle, synth):
    """The whole-file entry point equals the reference's per-batch calls on cropped, zero-padded chunks."""
    C_, K, Ls, step, F = 23, 3, 80000, 8000, 293
    n = (C_ - 1) * step + Ls - 12345          # the last chunks run past the end of the file -> zero padding
    wave = synth.waveform(4, n / 16000.0)[:n]
    assert wave.shape[0] == n
    masks = _masks(synth, 8, C_ * K, F)
    masks[5] = 0
    masks[40:64] = 0                           # one whole batch of 32 ends up too short
    masks[32:40] = 0
    sig, lens, ts, inv = ctx.mask_compact_file(wave, masks, C_, K, Ls, step)
    padded = np.concatenate([wave, np.zeros(Ls, np.float32)])
    items = np.stack([padded[c * step:c * step + Ls] for c in range(C_) for _ in range(K)])
    R = C_ * K
    for g0 in range(0, R, 32):
        sl = slice(g0, min(R, g0 + 32))
        rc, s_o, l_o, t_o = oracle.mask_compact(items[sl], masks[sl])
        assert rc == inv[g0 // 32]
        assert np.array_equal(s_o, sig[sl])
        if rc == 0:
            assert np.array_equal(l_o, lens[sl]) and np.array_equal(t_o, ts[sl])
    assert inv.tolist() == [0, 1, 0]


# ------------------------------------------------------------------ f2 / f3

def _pipeline_case(oracle, synth, C_, F, seed, n_speakers=3):
    seg = synth.segmentations(seed, C_, F, 3)
    b = oracle.binarize(seg)
    count, cf = oracle.speaker_count(b)
    emb, _ = synth.embeddings(seed + 1, C_, 3, 192, n_speakers=n_speakers, tiny=())
    _, hard, _, _ = oracle.clustering_stage(emb, b)
    sf = (0.0, 0.5, 5.0, 16000 * (C_ // 2 + 5))
    return seg, sf, hard, count, cf


@pytest.mark.parametrize("C_,seed", [(30, 21), (109, 33), (7, 5)])
def test_reconstruct_matches_oracle(ctx, oracle, synth, C_, seed):
    seg, sf, hard, count, cf = _pipeline_case(oracle, synth, C_, 293, seed)
    ro, fo = oracle.reconstruct(seg, sf, hard, count, cf)
    rg, fg = ctx.reconstruct(seg, sf, hard, count, cf)
    assert rg.shape == ro.shape and np.array_equal(rg, ro)
    assert np.array_equal(np.array([fg.start, fg.step, fg.duration]), fo)


def test_reconstruct_all_inactive_and_nan_scores(ctx, oracle, synth):
    seg, sf, hard, count, cf = _pipeline_case(oracle, synth, 20, 293, 3)
    hard2 = np.full_like(hard, -2)
    ro, _ = oracle.reconstruct(seg, sf, hard2, count, cf)
    rg, _ = ctx.reconstruct(seg, sf, hard2, count, cf)
    assert np.array_equal(rg, ro) and rg.shape[1] == 1
    seg2 = seg.copy()
    seg2[3, 10:50, 1] = np.nan
    hard3 = hard.copy()
    hard3[5] = [0, 0, 0]       # three local speakers on one cluster -> max over three
    ro, _ = oracle.reconstruct(seg2, sf, hard3, count, cf)
    rg, _ = ctx.reconstruct(seg2, sf, hard3, count, cf)
    assert np.array_equal(rg, ro)


def test_reconstruct_and_annotation_golden(ctx, golden_dir):
    g, sp, cl = (np.load(os.path.join(golden_dir, n)) for n in ("next_ref.npz", "segpost_ref.npz", "cluster_ref.npz"))
    sf = (0.0, 0.5, 5.0, 944000)
    rec, fr = ctx.reconstruct(sp["seg"], sf, cl["hard"], sp["count"], tuple(sp["cf"]))
    assert np.array_equal(rec.astype(np.uint8), g["rec"]) and set(np.unique(rec)) <= {0.0, 1.0}
    assert np.array_equal(np.array([fr.start, fr.step, fr.duration]), g["fr"])
    segs, labs = ctx.to_annotation(rec, (fr.start, fr.step, fr.duration, 0))
    ko, kr = np.lexsort((labs, segs[:, 1], segs[:, 0])), np.lexsort((g["labs"], g["segs"][:, 1], g["segs"][:, 0]))
    assert np.array_equal(segs[ko], g["segs"][kr]) and np.array_equal(labs[ko], g["labs"][kr])
    assert np.all(np.diff(segs[:, 0]) >= 0)


@pytest.mark.parametrize("rows,cols,seed", [(3439, 4, 1), (1, 3, 2), (2, 2, 3), (1025, 1, 4), (40000, 7, 5), (2050, 5, 6)])
@pytest.mark.parametrize("params", [(0.5, 0.5, 0.0, 0.5817029476165771), (0.6, 0.3, 0.2, 0.1), (0.5, 0.5, 0.0, 0.0),
                                    (0.3, 0.6, 0.05, 0.02)])
def test_to_annotation_matches_oracle(ctx, oracle, synth, rows, cols, seed, params):
    onset, offset, on, off = params
    rng = np.random.default_rng(seed)
    act = synth.turn_chain(seed, rows, cols).astype(np.float64)
    scores = np.clip(act * 0.8 + rng.uniform(0, 0.35, (rows, cols)), 0, 1)     # non-binary: exercises hysteresis
    scores[rng.random((rows, cols)) < 0.02] = 0.5                              # exact ties at the thresholds
    if seed % 2:
        scores = (scores > 0.5).astype(np.float64)
    frames = (0.0, 0.016875, 0.016875, 0)
    so, lo = oracle.to_annotation(scores, frames, onset, offset, on, off)
    sg, lg = ctx.to_annotation(scores, frames, onset, offset, on, off)
    assert sg.shape == so.shape
    assert np.array_equal(sg, so) and np.array_equal(lg, lo)     # the oracle's stable order == device merge order


def test_to_annotation_capacity_error(ctx, pkg, synth):
    scores = (np.arange(200)[:, None] % 2 == 0).astype(np.float64).repeat(2, 1)
    with pytest.raises(pkg.SdError) as e:
        ctx.to_annotation(scores, (0.0, 0.016875, 0.016875, 0), 0.5, 0.5, 0.0, 0.0, cap=3)
    assert e.value.code == 6


# ------------------------------------------------------------------ f4

def test_ingest_pcm16(ctx, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "ingest_ref.npz"))
    assert np.array_equal(ctx.ingest_pcm16(g["pcm"]), g["wav"])
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 9, 4097, 1 << 20):
        pcm = rng.integers(-32768, 32768, n).astype(np.int16)
        assert np.array_equal(ctx.ingest_pcm16(pcm), oracle.ingest_pcm16(pcm))


def test_slide_geometry_and_crop(ctx, oracle, synth, golden_dir):
    for n in (1, 2, 79999, 80000, 80001, 88000, 88001, 16000 * 60, 16000 * 600 + 123):
        assert ctx.slide_geometry(n) == oracle.slide_geometry(n), n
    assert ctx.slide_geometry(16000 * 600, 10.0, 1.0) == oracle.slide_geometry(16000 * 600, 10.0, 1.0)
    g = np.load(os.path.join(golden_dir, "ingest_ref.npz"))
    wave = synth.waveform(13, 7.3)
    got = ctx.crop_chunks(wave, g["crop_starts"])
    assert np.array_equal(got.astype(np.float64).sum(1), g["crop_sum"])
    assert np.array_equal(got[:, :4], g["crop_first"]) and np.array_equal(got[:, -4:], g["crop_last"])
    starts = np.concatenate([g["crop_starts"], [-0.75, -4.99, 7.2999]])   # starts past the end are UB in the reference
    got = ctx.crop_chunks(wave, starts)
    for row, t in zip(got, starts):
        assert np.array_equal(row, oracle.crop(wave, t))
