"""GPU: the whole-GPU merge loop (linkage_wide_kernel: one CTA per SM, global-memory mailbox per round; default from
6 144 rows) forced at small sizes (SD_OPT_LINKAGE_WIDE = 2) so that the CPU oracle can check every path bit for bit:
plain runs, sizes that leave some CTAs without rows, exact ties that must hand over to the heap kernel, NaN-free
dendrograms with non-monotone merge distances.  The full-size cases (cfg3 against the oracle, cfg5 against scipy's
SHA-256) run in test_gpu_fullsize.py with the default dispatch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SD_OPT_LINKAGE_WIDE = 5


@pytest.fixture()
def wctx(pkg):
    c = pkg.Context(0)
    c.set_option(SD_OPT_LINKAGE_WIDE, 2)
    yield c
    c.close()


@pytest.mark.parametrize("N,D,seed", [(2, 4, 0), (3, 8, 1), (33, 16, 2), (319, 192, 3), (1136, 192, 4), (1773, 192, 5),
                                      (4800, 64, 6)])
def test_wide_linkage_bit_exact(wctx, oracle, synth, N, D, seed):
    x, _ = synth.stress_embeddings(seed, N, D, 5)
    xn = oracle.normalize(x)
    Z = wctx.linkage(xn)
    assert np.array_equal(Z, oracle.linkage(xn))
    c = wctx.debug_counters(reset=True)
    assert c[2] == 0  # no hand-over to the exact kernel on continuous data


def test_wide_linkage_random_unclustered(wctx, oracle):
    """no cluster structure: many stale candidates and non-monotone merge distances"""
    rng = np.random.default_rng(9)
    x = oracle.normalize(rng.standard_normal((700, 24)))
    Z = wctx.linkage(x)
    assert np.array_equal(Z, oracle.linkage(x))
    assert (np.diff(Z[:, 2]) < 0).any()


def test_wide_linkage_ties_hand_over_to_exact_kernel(wctx, oracle):
    rng = np.random.default_rng(11)
    base = rng.standard_normal((40, 6))
    x = np.concatenate([base, base, base[:9]])  # duplicate rows: exactly equal distances
    assert np.array_equal(wctx.linkage(x), oracle.linkage(x))
    gx, gy = np.meshgrid(np.arange(7.0), np.arange(6.0))
    grid = np.stack([gx.ravel(), gy.ravel()], 1)
    assert np.array_equal(wctx.linkage(grid), oracle.linkage(grid))


def test_wide_clustering_stage_equals_oracle(wctx, oracle, synth):
    emb, _ = synth.embeddings(31, 500, 3, 192, n_speakers=5)
    hard, k = wctx.clustering(emb, None)[:2]
    rc, ho, _, ko = oracle.clustering_stage(emb)
    assert rc == 0 and k == ko and np.array_equal(hard, ho)
