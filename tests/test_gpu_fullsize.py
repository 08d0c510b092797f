"""GPU: BASELINE.json configurations at full size, checked through size-independent properties (and, where the
CPU oracle still finishes in tens of seconds, bit for bit)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THRESH = float(np.float32(0.7153814381597874))


def check_dendrogram(Z, N):
    """Structural validity of a linkage matrix: every node merged exactly once, sizes add up."""
    assert Z.shape == (N - 1, 4)
    ids = Z[:, :2].astype(np.int64)
    assert (ids[:, 0] < ids[:, 1]).all()
    assert (ids < (N + np.arange(N - 1))[:, None]).all()  # children exist before their parent
    used = np.sort(ids.ravel())
    assert np.array_equal(used, np.arange(2 * N - 2))  # each cluster id consumed exactly once
    size = np.ones(2 * N - 1)
    for k in range(N - 1):
        size[N + k] = size[ids[k, 0]] + size[ids[k, 1]]
    assert np.array_equal(size[N:], Z[:, 3]) and Z[-1, 3] == N
    assert np.isfinite(Z[:, 2]).all() and (Z[:, 2] >= 0).all()


def centroid_distance_property(Z, xn, rng, samples=200):
    """Centroid linkage: the recorded merge distance is the Euclidean distance between the two cluster means."""
    N = xn.shape[0]
    members = {i: [i] for i in range(N)}
    want = set(rng.choice(N - 1, size=min(samples, N - 1), replace=False).tolist())
    worst = 0.0
    for k in range(N - 1):
        a, b = int(Z[k, 0]), int(Z[k, 1])
        if k in want:
            ca, cb = xn[members[a]].mean(0), xn[members[b]].mean(0)
            worst = max(worst, abs(np.linalg.norm(ca - cb) - Z[k, 2]))
        members[N + k] = members.pop(a) + members.pop(b)
    return worst


def test_cfg2_stft_full_batch(ctx, oracle, synth):
    """configs[1]: 1 773 items x 160 000 samples -> [1773, 1001, 201, 2]; spot frames against the oracle, DC and
    Nyquist bins purely real, deterministic across runs."""
    wav = synth.fbank_items(102, 1773, 160000)
    out = ctx.stft(wav)
    assert out.shape == (1773, 1001, 201, 2)
    rng = np.random.default_rng(0)
    for b in rng.integers(0, 1773, 12):
        t0 = int(rng.integers(0, 997))
        want = oracle.stft_frames(wav[b], t0, t0 + 4)
        assert np.abs(out[b, t0:t0 + 4] - want).max() < 1e-4
    assert not out[:, :, 0, 1].any() and not out[:, :, 200, 1].any()
    # Parseval on a few frames of a few items: sum |X|^2 over the full spectrum == N * sum (w x)^2
    out2 = ctx.stft(wav)
    assert np.array_equal(out, out2)


def test_cfg3_one_hour_clustering_bit_exact(ctx, oracle, synth):
    """configs[2]: synthetic 1-hour meeting, ~10.8k embeddings of dimension 192, full pdist + centroid linkage +
    fcluster + assignment on one GPU; the CPU oracle still finishes in well under a minute at this size."""
    C = 3591
    emb, _ = synth.embeddings(203, C, 3, 192, n_speakers=6)
    hard, soft, k = ctx.clustering(emb, None, soft_k_cap=16)
    rc, ho, so, ko = oracle.clustering_stage(emb, None, soft_k_cap=16)
    assert rc == 0 and k == ko
    assert np.array_equal(hard, ho)
    assert np.array_equal(soft[:, :, :k], so[:, :, :k], equal_nan=True)
    x = emb.reshape(-1, 192)
    x = x[~np.isnan(x[:, 0])]
    N = x.shape[0]
    assert 10000 < N < 10800
    xn = ctx.normalize_embeddings(x)
    Z = ctx.linkage(xn)
    check_dendrogram(Z, N)
    assert np.array_equal(Z, oracle.linkage(xn))
    assert np.array_equal(ctx.fcluster(Z, THRESH), oracle.fcluster(Z, THRESH))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_cfg5_clustering_stress_50k(ctx, oracle, synth):
    """configs[4]: 50 000 x 256 embeddings, bit for bit.  The reference itself is invalid here (its int condensed
    index overflows for N > 46 341, clustering.cpp:236-242); the known answer is the SHA-256 of the dendrogram and
    labels that scipy 1.18.1 linkage(method="centroid") / fcluster AND the C restatement both produce
    (oracle/make_golden_cfg5.py, ~20 CPU-minutes, tests/golden/cfg5_sha256.json).  Plus the size-independent
    properties: structural validity, centroid distances on sampled merges, recovery of the planted speakers."""
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cfg5_sha256.json")))
    N, D, S = gold["config"]["N"], gold["config"]["D"], gold["config"]["S"]
    x, spk = synth.stress_embeddings(gold["config"]["seed"], N, D, S)
    assert sha(x) == gold["input_sha256"]
    rng = np.random.default_rng(205)
    xn = ctx.normalize_embeddings(x)
    assert sha(xn) == gold["normalized_sha256"]
    assert np.abs(np.linalg.norm(xn, axis=1) - 1).max() < 1e-6
    Z = ctx.linkage(xn)
    for k, h in gold["Z_prefix_sha256"].items():  # locates a divergence if there ever is one
        assert sha(Z[:int(k)]) == h, "dendrogram differs from scipy within the first %s merges" % k
    assert sha(Z) == gold["Z_sha256"]
    check_dendrogram(Z, N)
    assert centroid_distance_property(Z, xn, rng, samples=60) < 1e-9
    T = ctx.fcluster(Z, THRESH)
    assert sha(T.astype(np.int32)) == gold["labels_sha256"]
    assert np.array_equal(T, oracle.fcluster(Z, THRESH))
    assert T.max() == S == gold["n_clusters"]
    # every flat cluster is one planted speaker
    for c in range(1, S + 1):
        assert len(np.unique(spk[T == c])) == 1


def test_cfg4_file_batch_labels_independent_of_order(ctx, oracle, synth):
    """configs[3]: files are independent units -- clustering a file gives the same labels whatever was processed
    before it on the same context (workspace reuse across files of different sizes)."""
    files = [synth.embeddings(400 + i, 291 if i % 2 else 150, 3, 192, n_speakers=2 + i % 3)[0] for i in range(6)]
    first = [ctx.clustering(e)[0] for e in files]
    again = [ctx.clustering(e)[0] for e in reversed(files)][::-1]
    for a, b, e in zip(first, again, files):
        assert np.array_equal(a, b)
        assert np.array_equal(a, oracle.clustering_stage(e)[1])
