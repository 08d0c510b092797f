"""GPU: sd_batch_* -- a batch of files in flight (library-owned host threads, one sd_ctx per worker) must give, for
every file, exactly what the single-file entry points give (and therefore what the oracle gives), with host and with
device pointers, whatever the number of workers, and when the same files are submitted again (benchmark steps)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_inputs(synth, seed, Cn):
    F, S, L, D, Kd = 293, 3, 16000, 192, 4
    wav = synth.fbank_items(seed, Cn * S, L)
    seg = synth.segmentations(seed + 1, Cn, F, S)
    emb, _ = synth.embeddings(seed + 2, Cn, S, D, n_speakers=2 + seed % 3)
    diar = synth.segmentations(seed + 3, Cn, F, Kd).astype(np.float64)
    diar[::3, :, 1] = np.nan
    return dict(C=Cn, F=F, S=S, L=L, D=D, Kd=Kd, wav=wav, seg=seg, emb=emb, diar=diar,
                chunks=(0.0, 0.5, 5.0, int((Cn * 0.5 + 5.0) * 16000)))


def expected(oracle, x):
    b = oracle.binarize(x["seg"])
    count, _ = oracle.speaker_count(b)
    _, hard, _, k = oracle.clustering_stage(x["emb"], b)
    agg, _ = oracle.aggregate(x["diar"], x["chunks"], missing=0.0, skip_average=True)
    return dict(stft=oracle.stft(x["wav"]), binarized=b, count=count, hard=hard, k=k, diar=agg)


def host_file(pkg, x):
    NF = pkg.lib().sd_aggregate_num_frames(x["C"], C.byref(pkg._win(x["chunks"])), C.byref(pkg._win(pkg.FRAMES)))
    out = dict(stft=np.empty((x["C"] * x["S"], 1 + x["L"] // 160, 201, 2), np.float32),
               binarized=np.empty((x["C"], x["F"], x["S"]), np.float64), count=np.zeros(NF + 64, np.int32),
               hard=np.empty((x["C"], x["S"]), np.int32), diar=np.empty((NF, x["Kd"]), np.float64))
    f = pkg.make_file(x["C"], x["F"], x["S"], x["L"], x["D"], x["chunks"], Kd=x["Kd"], wav_items=x["wav"],
                      segmentations=x["seg"], embeddings=x["emb"], diar_scores=x["diar"], stft=out["stft"],
                      binarized=out["binarized"], count=out["count"], count_cap=out["count"].size, hard=out["hard"],
                      diar=out["diar"])
    return f, out


def check(f, out, want):
    assert f.status == 0
    assert np.abs(out["stft"] - want["stft"]).max() < 1e-4
    assert np.array_equal(out["binarized"], want["binarized"])
    assert f.n_count == want["count"].size and np.array_equal(out["count"][:f.n_count], want["count"])
    assert f.num_clusters == want["k"] and np.array_equal(out["hard"], want["hard"])
    assert f.n_diar == want["diar"].shape[0] and np.array_equal(out["diar"], want["diar"])


@pytest.mark.parametrize("workers", [1, 3, 8])
def test_batch_host_pointers_equal_oracle(pkg, synth, oracle, workers):
    xs = [make_inputs(synth, 10 * i, 20 + 7 * i) for i in range(5)]
    wants = [expected(oracle, x) for x in xs]
    pairs = [host_file(pkg, x) for x in xs]
    files = (pkg.SdFile * len(pairs))(*[p[0] for p in pairs])
    b = pkg.Batch(0, workers)
    try:
        for _ in range(2):  # the same files again: file i stays on worker i % workers, so no self-overlap
            for _, out in pairs:
                out["hard"][...] = -7
            b.run(files, pkg.SD_BATCH_HOST)
            for i, (_, out) in enumerate(pairs):
                check(files[i], out, wants[i])
    finally:
        b.close()


def test_batch_device_pointers_and_resubmission(pkg, synth, oracle):
    ctx = pkg.Context(0)
    xs = [make_inputs(synth, 100 + 10 * i, 30 + 5 * i) for i in range(4)]
    wants = [expected(oracle, x) for x in xs]
    dev, host = [], []
    for x in xs:
        f, out = host_file(pkg, x)
        d = dict(wav_items=ctx.to_device(x["wav"]), segmentations=ctx.to_device(x["seg"]),
                 embeddings=ctx.to_device(x["emb"]), diar_scores=ctx.to_device(x["diar"]))
        for k in ("stft", "binarized", "count", "hard", "diar"):
            d[k] = ctx.malloc(out[k].nbytes)
        dev.append(d)
        host.append(out)
    files = (pkg.SdFile * len(xs))(*[
        pkg.make_file(x["C"], x["F"], x["S"], x["L"], x["D"], x["chunks"], Kd=x["Kd"], count_cap=o["count"].size, **d)
        for x, d, o in zip(xs, dev, host)])
    b = pkg.Batch(0, 4)
    try:
        b.submit(files, pkg.SD_BATCH_DEVICE)   # two "steps" queued back to back before waiting
        b.submit(files, pkg.SD_BATCH_DEVICE)
        b.wait()
        for i, (d, out) in enumerate(zip(dev, host)):
            for k in ("stft", "binarized", "count", "hard", "diar"):
                ctx.d2h(out[k], d[k])
            check(files[i], out, wants[i])
    finally:
        b.close()
        for d in dev:
            for p in d.values():
                ctx.free(p)
        ctx.close()


def test_batch_reports_the_failing_file(pkg, synth):
    x = make_inputs(synth, 7, 12)
    x["emb"][0, 0, :] = 0.0  # a zero vector: the reference throws "Vectors have zero magnitude." (SD:493-495)
    good = make_inputs(synth, 8, 12)
    pairs = [host_file(pkg, x), host_file(pkg, good)]
    files = (pkg.SdFile * 2)(*[p[0] for p in pairs])
    b = pkg.Batch(0, 2)
    try:
        with pytest.raises(pkg.SdError) as e:
            b.run(files)
        assert e.value.code == pkg.SD_ERR_ZERO_MAGNITUDE
        assert files[0].status == pkg.SD_ERR_ZERO_MAGNITUDE and files[1].status == 0
        b.run((pkg.SdFile * 1)(files[1]))  # the batch stays usable
    finally:
        b.close()


def test_one_long_file_split_by_chunk_range(pkg, synth, oracle):
    """inside one long file the STFT items and binarize rows of a chunk range are independent, so the
    range can run as its own sd_file (here three ranges on three workers; on a multi-GPU box: on three GPUs), while
    count / clustering / aggregate run once over all chunks.  Results equal the unsplit file's."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("sdb200_shard", os.path.join(os.path.dirname(pkg.__file__), "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    x = make_inputs(synth, 55, 41)
    want = expected(oracle, x)
    f_all, out = host_file(pkg, x)
    ranges = shard.split_chunk_range(x["C"], 3)
    assert ranges == [(0, 14), (14, 28), (28, 41)]
    parts = []
    S, L, F = x["S"], x["L"], x["F"]
    for c0, c1 in ranges:
        parts.append(pkg.make_file(c1 - c0, F, S, L, x["D"], x["chunks"], wav_items=x["wav"][c0 * S:c1 * S],
                                   segmentations=x["seg"][c0:c1], stft=out["stft"][c0 * S:c1 * S],
                                   binarized=out["binarized"][c0:c1]))
    b = pkg.Batch(0, 3)
    try:
        b.run((pkg.SdFile * 3)(*parts))  # front-end stages, range by range
        assert np.abs(out["stft"] - want["stft"]).max() < 1e-4 and np.array_equal(out["binarized"], want["binarized"])
        rest = pkg.make_file(x["C"], F, S, L, x["D"], x["chunks"], Kd=x["Kd"], embeddings=x["emb"], diar_scores=x["diar"],
                             binarized=out["binarized"], count=out["count"], count_cap=out["count"].size,
                             hard=out["hard"], diar=out["diar"])
        arr = (pkg.SdFile * 1)(rest)
        b.run(arr)  # the stages that need every chunk
        assert arr[0].n_count == want["count"].size and np.array_equal(out["count"][:arr[0].n_count], want["count"])
        assert np.array_equal(out["hard"], want["hard"]) and np.array_equal(out["diar"], want["diar"])
    finally:
        b.close()


@pytest.mark.parametrize("cfg", [dict(stft_chain=0, linkage_cluster=1), dict(stft_chain=2, linkage_cluster=0),
                                 dict(narrow_sms=16), dict(narrow_sms=24, linkage_cluster=0, stft_chain=0)])
def test_batch_configurations_give_identical_results(pkg, synth, oracle, cfg):
    """sd_batch_config (STFT launch order, merge-loop kernel, SM partition through green contexts) changes how the
    files share the GPU, never what they compute: every configuration equals the oracle, with host and device
    pointers mixed over two submissions."""
    xs = [make_inputs(synth, 300 + 10 * i, 18 + 9 * i) for i in range(6)]
    wants = [expected(oracle, x) for x in xs]
    pairs = [host_file(pkg, x) for x in xs]
    files = (pkg.SdFile * len(pairs))(*[p[0] for p in pairs])
    b = pkg.Batch(0, 4, **cfg)
    try:
        got = b.config()
        for k, v in cfg.items():
            if k == "narrow_sms":
                assert got[k] in (0, v) or got[k] >= v  # granted size (rounded up), or 0 if the driver refused
            else:
                assert got[k] == v
        for _ in range(2):
            for _, out in pairs:
                out["hard"][...] = -7
                out["stft"][...] = 0
            b.run(files, pkg.SD_BATCH_HOST)
            for i, (_, out) in enumerate(pairs):
                check(files[i], out, wants[i])
    finally:
        b.close()


def test_batch_trace_writes_a_timeline(pkg, synth, tmp_path, monkeypatch):
    """SDB_BATCH_TRACE: one line per file with the seven %globaltimer stamps in order."""
    path = tmp_path / "trace.csv"
    monkeypatch.setenv("SDB_BATCH_TRACE", str(path))
    xs = [make_inputs(synth, 400 + i, 15) for i in range(3)]
    pairs = [host_file(pkg, x) for x in xs]
    b = pkg.Batch(0, 2)
    try:
        b.run((pkg.SdFile * 3)(*[p[0] for p in pairs]))
    finally:
        b.close()
    rows = path.read_text().strip().splitlines()
    assert rows[0].startswith("worker,start,stft_done") and len(rows) == 4
    for line in rows[1:]:
        t = [float(v) for v in line.split(",")[1:]]
        assert all(a <= b_ for a, b_ in zip(t, t[1:])), line
