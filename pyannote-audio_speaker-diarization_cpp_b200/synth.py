"""Seeded synthetic inputs for tests and bench.py.

The ONNX blobs of the reference are missing, so the outputs of segment2.onnx (per-chunk speaker
activity scores) and emd4.onnx (speaker embeddings) are replaced by seeded stand-ins of the same
shape, dtype and statistics.  Everything here is numpy only.
"""
import numpy as np

SAMPLE_RATE = 16000


def chunk_geometry(duration_s, window_s=5.0, step_s=0.5):
    """Number of chunks the reference's SegmentModel::slide produces (speakerDiarizer.cpp:1407-1480):
    full windows while i + window < n, plus one zero-padded tail chunk."""
    n = int(round(duration_s * SAMPLE_RATE))
    w = int(round(window_s * SAMPLE_RATE))
    s = int(round(step_s * SAMPLE_RATE))
    c, i = 0, 0
    while i + w < n:
        c += 1
        i += s
    if i + 1 < n:
        c += 1
    return c


def frames_per_chunk(window_s):
    """293 for the reference's 5 s window; 589 for 10 s (pyannote 3.x segmentation)."""
    return {5.0: 293, 10.0: 589}.get(float(window_s), int(round(window_s * 58.9)))


def turn_chain(seed, n_steps, n_speakers, p_switch=0.02, p_overlap=0.1):
    """Markov turn-taking: active[t, s] in {0,1} at 10 ms resolution."""
    rng = np.random.default_rng(seed)
    act = np.zeros((n_steps, n_speakers), np.uint8)
    cur = 0
    second = -1
    for t in range(n_steps):
        u = rng.random()
        if u < p_switch:
            cur = int(rng.integers(n_speakers))
            second = int(rng.integers(n_speakers)) if rng.random() < p_overlap else -1
        act[t, cur] = 1
        if second >= 0:
            act[t, second] = 1
    return act


def waveform(seed, duration_s, n_speakers=4):
    """Sum of speaker-specific harmonic stacks gated by a turn chain + white noise, fp32 in [-1, 1]."""
    rng = np.random.default_rng(seed)
    n = int(round(duration_s * SAMPLE_RATE))
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    act = turn_chain(seed + 7, n // 160 + 1, n_speakers)
    gate = np.repeat(act, 160, axis=0)[:n].astype(np.float64)
    x = np.zeros(n)
    for s in range(n_speakers):
        f0 = rng.uniform(90.0, 220.0)
        v = np.zeros(n)
        for h in range(1, 9):
            v += np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / h
        x += 0.2 * v * gate[:, s]
    x += 1e-3 * rng.standard_normal(n)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def segmentations(seed, C, F, K=3, silent_frac=0.08):
    """Stand-in for segment2.onnx output [C, F, K] fp32 in (0, 1): sigmoid of a smooth AR(1) track per
    (chunk, local speaker); a fraction of (chunk, speaker) tracks are forced inactive."""
    rng = np.random.default_rng(seed)
    x = np.empty((C, F, K), np.float32)
    e = rng.standard_normal((C, F, K))
    a = 0.97
    z = np.zeros((C, K))
    bias = rng.uniform(-2.5, 1.5, size=(C, K))
    for f in range(F):
        z = a * z + np.sqrt(1 - a * a) * e[:, f, :] * 3.0
        x[:, f, :] = 1.0 / (1.0 + np.exp(-(z + bias)))
    silent = rng.random((C, K)) < silent_frac
    x[np.broadcast_to(silent[:, None, :], x.shape)] *= 0.05
    return x


def embeddings(seed, C, S=3, D=192, n_speakers=4, nan_frac=0.05, tiny=(3, 5), sigma=None):
    """Stand-in for emd4.onnx output [C, S, D] fp64 (the reference widens to double,
    speakerDiarizer.cpp:2555): unit speaker centroids * random gain r~U(5,30) + isotropic noise so that the
    intra-speaker distance after normalisation is ~0.45 (< 0.7154 threshold) and inter ~1.3-1.4;
    `nan_frac` rows are all-NaN (too-short speech); `tiny` adds speakers with < 15 items."""
    rng = np.random.default_rng(seed)
    n_all = n_speakers + len(tiny)
    cen = rng.standard_normal((n_all, D))
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    if sigma is None:
        sigma = 0.45 / np.sqrt(2.0 * D)
    R = C * S
    spk = rng.integers(0, n_speakers, size=R)
    pos = rng.permutation(R)
    o = 0
    for i, cnt in enumerate(tiny):
        spk[pos[o:o + cnt]] = n_speakers + i
        o += cnt
    gain = rng.uniform(5.0, 30.0, size=(R, 1))
    e = (cen[spk] + sigma * rng.standard_normal((R, D))) * gain
    e = e.astype(np.float32).astype(np.float64)  # values that came out of an fp32 network
    nan_rows = rng.random(R) < nan_frac
    e[nan_rows] = np.nan
    return e.reshape(C, S, D), spk.reshape(C, S)


def fbank_items(seed, n_items, L):
    """Masked/compacted chunk signals as handed to EmbeddingModel1::infer: speech-like content followed
    by a zero tail of random length (padSequence, speakerDiarizer.cpp:770-797)."""
    rng = np.random.default_rng(seed)
    x = (0.1 * rng.standard_normal((n_items, L))).astype(np.float32)
    t = np.arange(L, dtype=np.float32) / SAMPLE_RATE
    for i in range(n_items):
        f0 = rng.uniform(90.0, 220.0)
        x[i] += (0.3 * np.sin(2 * np.pi * f0 * t) + 0.15 * np.sin(2 * np.pi * 2 * f0 * t)).astype(np.float32)
        keep = int(rng.uniform(0.05, 1.0) * L)
        x[i, keep:] = 0.0
    return np.clip(x, -1.0, 1.0)


def stress_embeddings(seed=205, N=50000, D=256, S=12):
    """configs[4] (clustering stress): N un-normalised D-dimensional embeddings of S planted speakers.
    Returns (x [N, D] fp64, speaker [N])."""
    rng = np.random.default_rng(seed)
    cen = rng.standard_normal((S, D))
    cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    spk = rng.integers(0, S, N)
    x = (cen[spk] + (0.45 / np.sqrt(2 * D)) * rng.standard_normal((N, D))) * rng.uniform(5, 30, (N, 1))
    return x, spk
