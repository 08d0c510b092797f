"""Small pass over every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool racecheck  python scripts/racecheck_run.py      (shared-memory hazards)
    compute-sanitizer --tool memcheck   python scripts/racecheck_run.py
    compute-sanitizer --tool synccheck  python scripts/racecheck_run.py

Sizes are small (the tools slow kernels down 10-100x) but chosen so that every single-CTA kernel with flag-style
loops runs its multi-round paths: fcluster_par (pointer jumping), linkage_fast / linkage_cluster / linkage (heap),
cluster_post*, annot_* (block scans), plus the tiled front-end kernels with their mbarrier pipelines.  Results are
checked against the CPU oracle so a race that corrupts data is also seen as a mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from oracle import oracle as O  # noqa: E402

pkg, synth, o = ge.load_package(), ge.load_synth(), Oracle()
ctx = pkg.Context(0)
ok = True


def check(name, cond):
    global ok
    print(("ok   " if cond else "FAIL ") + name, flush=True)
    ok = ok and bool(cond)


# front-end: interior (bulk copy + mbarrier) and edge tiles, both kernels, both modes
wav = synth.fbank_items(1, 3, 16000)
check("stft", np.abs(ctx.stft(wav) - o.stft(wav)).max() < 1e-4)
check("fbank", np.isfinite(ctx.fbank(wav, np.ones(3, np.float32))).all())
kst = ctx.stft(wav, params=ctx.stft_kaldi_params(snip_edges=False))
check("stft kaldi", np.abs(kst[1] - O.kaldi_stft(wav[1], snip_edges=False)).max() < 1e-4)
kfb = ctx.fbank(wav, params=ctx.fbank_kaldi_params(snip_edges=True))
check("fbank kaldi", np.abs(kfb[0] - O.kaldi_fbank(wav[0], 80, snip_edges=True))[O.kaldi_fbank(wav[0], 80, snip_edges=True) > -8].max() < 1e-3)

# the tile loop proper: more tiles than resident CTAs, so that the dynamic hand-out (global counter, next-tile
# coordinates passed through shared memory behind the end-of-tile barrier), the pair-lane exchange and the bulk-copy
# pipeline run for several rounds per CTA; every build of the kernel (SD_OPT_STFT_VARIANT)
big = synth.fbank_items(5, 12, 160000)
want = o.stft(big)
for variant in (0, 3, 5, 6, 2):
    ctx.set_option(3, variant)
    got = ctx.stft(big)
    check("stft 12 x 160000, variant %d" % variant, np.abs(got - want).max() < 1e-4)
    got = ctx.stft(big)  # second launch: the counter was reset by the last CTA of the first
    check("stft again, variant %d" % variant, np.abs(got - want).max() < 1e-4)
ctx.set_option(3, 0)
fb = ctx.fbank(big[:6], np.ones(6, np.float32))
check("fbank 6 x 160000", np.isfinite(fb).all())
if os.environ.get("SDB_SANITIZE_ONLY") == "frontend":
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)

# segmentation post-processing
seg = synth.segmentations(2, 24, 293, 3)
b = ctx.binarize_swf(seg)
check("binarize", np.array_equal(b, o.binarize(seg)))
check("speaker_count", np.array_equal(ctx.speaker_count(b)[0], o.speaker_count(b)[0]))

# clustering: every linkage kernel, the parallel fcluster, post-processing with small clusters, assignment
# option 4 = SD_OPT_LINKAGE_CLUSTER (0: one-CTA kernel), option 1 = SD_OPT_FORCE_EXACT_LINKAGE (heap kernel)
for n_chunks, opt in ((110, {}), (110, {4: 0}), (110, {1: 1}), (400, {})):
    for k, v in opt.items():
        ctx.set_option(k, v)
    emb, _ = synth.embeddings(3 + n_chunks, n_chunks, 3, 192, n_speakers=4, tiny=(3, 5))
    hard, kk = ctx.clustering(emb, None)[:2]
    rc, ho, _, ko = o.clustering_stage(emb)
    check("clustering C=%d opts=%s" % (n_chunks, opt), rc == 0 and np.array_equal(hard, ho))
    x = emb.reshape(-1, 192)
    x = o.normalize(x[~np.isnan(x[:, 0])])
    Z = ctx.linkage(x)
    check("linkage N=%d" % x.shape[0], np.array_equal(Z, o.linkage(x)))
    check("fcluster", np.array_equal(ctx.fcluster(Z, 0.7153814435005188), o.fcluster(Z, 0.7153814435005188)))
    for k in opt:
        ctx.set_option(k, 1 if k == 4 else 0)

# next rows: masking, reconstruct, to_annotation
masks = (synth.segmentations(5, 8, 293, 3)[:, :, 0] > 0.5).astype(np.float32)
w8 = synth.fbank_items(6, 8, 80000)
rc, s_o, l_o, t_o = o.mask_compact(w8, masks)
rc_g, sig, lens, ts = ctx.mask_compact(w8, masks)
check("mask_compact", rc_g == rc and np.array_equal(sig, s_o) and np.array_equal(lens, l_o))
count, cf = o.speaker_count(b)
hard24 = (np.arange(24 * 3).reshape(24, 3) % 3).astype(np.int32)
sf = (0.0, 0.5, 5.0, 16000 * 17)
dd, fr = ctx.reconstruct(seg, sf, hard24, count, cf)
dd_o, fr_o = o.reconstruct(seg, sf, hard24, count, cf)
check("reconstruct", np.array_equal(dd, dd_o))
frames = (fr.start, fr.step, fr.duration, 0)
t1 = ctx.to_annotation(dd, frames)
t2 = o.to_annotation(dd_o, frames)
check("to_annotation", len(t1[0]) == len(t2[0]) and np.array_equal(t1[0], t2[0]) and np.array_equal(t1[1], t2[1]))
ctx.close()
print("RACECHECK-RUN " + ("PASSED" if ok else "FAILED"))
sys.exit(0 if ok else 1)
