"""GPU: the drop-in, proven on the reference's own entry point and its own test wav (BASELINE configs[0]).

tests/dropin/build/dropin_sdb200 is the reference program -- main(), speakerDiarization() (speakerDiarizer.cpp:
2937-3234), the ONNX wrappers, the final printout -- with the hot-path BODIES swapped for sdb200::* calls by
tests/dropin/patch_reference.py, built with -DWRITE_DATA, linked against libsdb200.so only
(no libtorch, no clustering.cpp).  The two ONNX forward passes are answered by deterministic, input-independent
stand-ins (the blobs are missing from the reference checkout).  The UNMODIFIED reference, built the same way, was run
once where /root/reference exists (tests/dropin/run_reference.py, ~10 CPU-minutes); what it wrote is frozen in
tests/golden/dropin_reference.json (+ dropin_ort_inputs.npz).  Here the patched program runs on the GPU and must give

  * every /tmp/cpp_<stage>.txt dump of pipeline/script/verifyEveryStepResult.py:6-17 BYTE-identical (SHA-256),
    in particular the ones written from inside the replaced bodies (SD:1271-1275, 1627-1636, 2074, 2186, 2206,
    2330-2331, ...) -- the shim re-emits them through the reference's own debugWrite* templates;
  * the tensors handed to emd4.onnx within 1e-4 abs (north_star bar; measured ~4e-6) and identical wav_lens;
  * identical printed speaker segments (speakerDiarizer.cpp:3437-3440).
"""
import glob
import gzip
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "dropin", "build", "dropin_sdb200")
WAV = os.path.join(ROOT, "oracle", "_ref", "multi-speaker_1min.wav")
KEEP = os.path.join(ROOT, "oracle", "_ref", "dropin")
GOLD = os.path.join(ROOT, "tests", "golden")


def _run_reference_module():
    spec = importlib.util.spec_from_file_location("dropin_run_reference",
                                                  os.path.join(ROOT, "tests", "dropin", "run_reference.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _first_difference(name):
    """human-readable location of the first differing line, when the reference's dump travelled with the snapshot"""
    ref = os.path.join(KEEP, name + ".gz")
    if not os.path.exists(ref):
        return "(reference dump not on this machine)"
    with gzip.open(ref, "rt") as a, open("/tmp/" + name) as b:
        for n, (la, lb) in enumerate(zip(a, b)):
            if la != lb:
                return "line %d:\n  reference: %s\n  sdb200   : %s" % (n + 1, la[:300], lb[:300])
    return "files differ in length"


@pytest.fixture(scope="module")
def dropin_run(tmp_path_factory):
    gold_path = os.path.join(GOLD, "dropin_reference.json")
    assert os.path.exists(gold_path), "tests/golden/dropin_reference.json missing: make -C tests/dropin reference-dumps"
    if not (os.path.exists(EXE) and os.path.exists(WAV)):
        pytest.skip("drop-in binary / reference wav not built (needs /root/reference at build time)")
    gold = json.load(open(gold_path))
    rr = _run_reference_module()
    assert rr.sha_file(WAV) == gold["wav_sha256"]
    for p in glob.glob("/tmp/cpp_*.txt"):
        os.remove(p)
    cap = str(tmp_path_factory.mktemp("ort_capture"))
    r = subprocess.run([EXE, WAV, cap], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    rec = {}
    caps = rr.collect(cap, r.stdout, rec)
    return gold, rec, caps, rr


def test_every_stage_dump_is_byte_identical(dropin_run):
    gold, rec, _, _ = dropin_run
    assert sorted(rec["dumps"]) == sorted(gold["dumps"]), "the two builds wrote different sets of stage dumps"
    stages = {n for n in gold["dumps"]}
    for must in ("cpp_aggregated_output.txt", "cpp_binary_ndarray.txt", "cpp_same_as.txt", "cpp_well_defined_idx.txt",
                 "cpp_filtered_embeddings.txt", "cpp_norm_embeddings.txt", "cpp_clusters.txt", "cpp_dist.txt",
                 "cpp_soft_clusters.txt", "cpp_hard_clusters.txt", "cpp_count.txt", "cpp_discrete_diarization.txt",
                 "cpp_clustered_segmentations.txt", "cpp_sorted_speakers.txt", "cpp_imasks0.txt", "cpp_wav_lens0.txt"):
        assert must in stages, must
    bad = [n for n in sorted(gold["dumps"]) if rec["dumps"][n] != gold["dumps"][n]]
    assert not bad, "\n".join("%s: %s" % (n, _first_difference(n)) for n in bad)


def test_printed_segments_are_identical(dropin_run):
    gold, rec, _, _ = dropin_run
    assert len(gold["segments"]) >= 8  # a real multi-speaker result, not an empty annotation
    assert rec["segments"] == gold["segments"]


def test_embedding_model_inputs_within_1e4(dropin_run):
    gold, rec, caps, rr = dropin_run
    assert sorted(rec["ort_inputs"]) == sorted(gold["ort_inputs"]) and len(caps) == 11  # 327 items -> 11 batches of 32
    g = np.load(os.path.join(GOLD, "dropin_ort_inputs.npz"))
    frames, sample, lens = g["frames"], g["sample"], g["wav_lens"]
    assert np.array_equal(frames, rr.sample_frames(501))  # T = 1 + 80000 / 160
    full_path = os.path.join(KEEP, "ort_inputs_full.npz")
    full = np.load(full_path)["x"] if os.path.exists(full_path) else None
    worst = 0.0
    for n, p in enumerate(caps):
        x = np.fromfile(p, np.float32).reshape(32, 501, 201, 2)
        wl = np.fromfile(p.replace("emb_input", "emb_lens"), np.float32)
        assert np.array_equal(wl, lens[n])  # wav_lens: bit-identical
        worst = max(worst, float(np.abs(x[:, frames] - sample[n]).max()))
        if full is not None:
            worst = max(worst, float(np.abs(x - full[n]).max()))
        assert np.abs(x).max() > 1.0  # a real spectrum
    print("max |STFT(sdb200) - STFT(reference, libtorch fp64)| = %.3g (%s)" %
          (worst, "all elements" if full is not None else "sampled frames"))
    assert worst < 1e-4


def _read_dump(name):
    rows = []
    with open("/tmp/" + name) as f:
        for line in f:
            line = line.strip().rstrip(",")
            if line:
                rows.append([float(v) for v in line.split(",")])
    return np.array(rows)


def test_select_masks_against_the_pipeline_own_choice(dropin_run, pkg):
    """sd_select_masks_dev (the clean-vs-raw mask choice, speakerDiarizer.cpp:3047-3082) is inline code of
    speakerDiarization() in the reference, so it can only be observed through the pipeline: the masks handed to
    getEmbedding are dumped as cpp_masks<n> (SD:2453), and those dumps are byte-identical to the unmodified
    reference's (test above).  Feed the dumped binarized segmentations to the kernel and compare."""
    gold, rec, _, _ = dropin_run
    assert all(rec["dumps"][n] == gold["dumps"][n] for n in rec["dumps"] if n.startswith("cpp_masks"))
    b = _read_dump("cpp_binarized_segmentations.txt").reshape(-1, 293, 3)
    n_items = b.shape[0] * 3
    masks = np.concatenate([_read_dump("cpp_masks%d.txt" % n) for n in range((n_items + 31) // 32)])
    assert masks.shape == (n_items, 293)
    min_num_frames = float(np.ceil(293 * 640 / (5.0 * 16000)))  # SD:3013
    ctx = pkg.Context(0)
    try:
        got = ctx.select_masks(b, min_num_frames)
    finally:
        ctx.close()
    assert np.array_equal(got, masks.astype(np.float32))
    clean = b * (b.sum(2, keepdims=True) < 2)
    assert (clean != b).any(), "the scenario has no overlapped speech: the choice would be vacuous"
