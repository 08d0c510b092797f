"""CPU: tests/dropin/patch_reference.py against the reference source (needs /root/reference; skipped on the GPU box).
Every hot-path body must have been replaced by its sdb200:: call, libtorch must be gone, and everything the patch does
not name must be untouched (the patched file differs from the reference only inside the replaced bodies)."""
import difflib
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pipeline/src/speakerDiarizer.cpp"


@pytest.fixture(scope="module")
def patched():
    if not os.path.exists(REF):
        pytest.skip("/root/reference not present")
    spec = importlib.util.spec_from_file_location("patch_reference", os.path.join(ROOT, "tests", "dropin", "patch_reference.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    src = open(REF).read()
    return src, m.patch(src)


def test_every_hot_path_body_calls_the_shim(patched):
    _, out = patched
    for call in ("sdb200::cleanSegmentations(", "sdb200::aggregate(", "sdb200::binarize_swf(", "sdb200::binarize_ndarray(",
                 "sdb200::crop(", "sdb200::speaker_count(", "sdb200::trim(", "sdb200::embedding_input(",
                 "sdb200::run_embedding_model<", "sdb200::Cluster().clustering(", "sdb200::Cluster().cluster(",
                 "sdb200::masked_signals(", "sdb200::reconstruct(", "sdb200::to_annotation<", "sdb200::read_wav("):
        assert out.count(call) == 1, call
    assert '#include "sdb200_host.hpp"' in out and "#define SDB200_DUMP2(data, ...) debugWrite2d(data, __VA_ARGS__)" in out


def test_libtorch_and_the_replaced_code_are_gone(patched):
    _, out = patched
    code = "\n".join(l for l in out.splitlines() if not l.lstrip().startswith("//"))
    for gone in ("torch::", "#include <torch", "Helper::interpolate( masks", "Helper::padSequence( dataChunks",
                 "Clustering::cluster( normalizedEmbeddings", "Helper::numpy_where( same_as", "wav::WavReader wav_reader( waveFile )"):
        assert gone not in code, gone


def test_nothing_else_changed(patched):
    src, out = patched
    a, b = src.splitlines(), out.splitlines()
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    removed = sum(i2 - i1 for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag in ("replace", "delete"))
    added = sum(j2 - j1 for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag in ("replace", "insert"))
    assert added < 120 and removed > 900  # ~1000 reference lines become ~80 lines of calls
    # callers and control flow stay: speakerDiarization() still slides, binarizes, embeds, clusters, reconstructs
    for kept in ("auto segmentations = mm.slide( input_wav, res_frames );", "auto binarized = mm.binarize_swf( segmentations, false );",
                 "cst.clustering( embeddings1, binarized, hard_clusters );", "auto embedding = getEmbedding( em, batchData, batchMasks );",
                 "hard_clusters[i][j] = -2;", "auto diarization = to_annotation( discrete_diarization,"):
        assert kept in out, kept
