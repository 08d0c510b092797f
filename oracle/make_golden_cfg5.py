"""TEST INFRASTRUCTURE -- pins configs[4] (clustering stress: 50 000 x 256 embeddings) bit for bit.

The reference itself is invalid at this size (its `int` condensed index overflows for N > 46 341,
pipeline/src/clustering/clustering.cpp:236-242), so the known answer comes from the two CPU implementations of the
same algorithm that are bit-identical to the reference below that limit (tests/test_oracle_vs_ref.py):

  * scipy 1.18.1 `linkage(y, method="centroid")` + `fcluster(Z, t, "distance")` (the code clustering.cpp was ported
    from, pipeline/src/clustering/README.md:5-8), fed the condensed distances of the C restatement (sdo_pdist,
    the reference's sequential mul+add Euclidean distance, clustering.cpp:408-415);
  * the C restatement `sdo_linkage_condensed` / `sdo_fcluster` (oracle/sd_oracle.c, 64-bit condensed index).

Both must agree; the SHA-256 of Z (fp64, C order) and of the labels (int32) go to tests/golden/cfg5_sha256.json,
with prefix hashes of Z so that a diverging GPU run can be located.  ~20 CPU-minutes, ~25 GB RAM.

    python oracle/make_golden_cfg5.py [--skip-scipy]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
import __graft_entry__ as ge  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

THRESH = float(np.float32(0.7153814381597874))
PREFIXES = (1000, 10000, 25000, 40000, 49000)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    synth = ge.load_synth()
    o = Oracle()
    N, D, S, seed = 50000, 256, 12, 205
    x, spk = synth.stress_embeddings(seed, N, D, S)
    xn = o.normalize(x)
    t0 = time.time()
    y = o.pdist(xn)
    print("pdist %.0f s" % (time.time() - t0), flush=True)
    t0 = time.time()
    Z = o.linkage_condensed(y, N)
    T = o.fcluster(Z, THRESH)
    print("C restatement linkage %.0f s, %d clusters" % (time.time() - t0, T.max()), flush=True)
    rec = {"config": {"N": N, "D": D, "S": S, "seed": seed, "cutoff": THRESH},
           "input_sha256": sha(x), "normalized_sha256": sha(xn), "Z_sha256": sha(Z), "labels_sha256": sha(T),
           "Z_prefix_sha256": {str(k): sha(Z[:k]) for k in PREFIXES}, "n_clusters": int(T.max()),
           "Z_first": Z[:4].tolist(), "Z_last": Z[-4:].tolist(), "sources": ["sd_oracle.c"]}
    if "--skip-scipy" not in sys.argv:
        import scipy
        from scipy.cluster.hierarchy import fcluster, linkage
        t0 = time.time()
        Zs = linkage(y, method="centroid")
        Ts = fcluster(Zs, THRESH, "distance").astype(np.int32)
        print("scipy %s linkage %.0f s" % (scipy.__version__, time.time() - t0), flush=True)
        assert np.array_equal(Zs, Z), "scipy and the C restatement disagree on Z"
        assert np.array_equal(Ts, T), "scipy and the C restatement disagree on the labels"
        rec["sources"].append("scipy " + scipy.__version__)
    out = os.path.join(ROOT, "tests", "golden", "cfg5_sha256.json")
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
