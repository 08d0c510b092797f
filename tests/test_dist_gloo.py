"""CPU, world_size 2, gloo: the multi-GPU host logic (file sharding + result gather) without any GPU."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_shard():
    sys.path.insert(0, ROOT)
    import importlib.util
    p = os.path.join(ROOT, "pyannote-audio_speaker-diarization_cpp_b200", "shard.py")
    spec = importlib.util.spec_from_file_location("sdb200_shard", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _worker(rank, world, port, durations, out_dir):
    import torch.distributed as dist
    shard = _load_shard()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    bins = shard.assign_files(durations, world)
    mine = bins[rank]
    # stand-in for the per-file GPU work: labels derived from the file id
    labels = [np.arange(10 + fid, dtype=np.int32) % (fid + 2) for fid in mine]
    packed = shard.pack_results(mine, labels, max_len=64)
    res = shard.gather_results(packed, max(len(b) for b in bins))
    ok = sorted(res) == list(range(len(durations))) and all(
        np.array_equal(res[f], np.arange(10 + f, dtype=np.int32) % (f + 2)) for f in res)
    open(os.path.join(out_dir, "rank%d.txt" % rank), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


def test_assign_files_balanced():
    shard = _load_shard()
    d = [300.0] * 64
    bins = shard.assign_files(d, 8)
    assert sorted(sum(bins, [])) == list(range(64)) and all(len(b) == 8 for b in bins)
    d = [3600, 60, 60, 60, 600, 600, 30, 30]
    bins = shard.assign_files(d, 2)
    loads = [sum(d[i] for i in b) for b in bins]
    assert sorted(sum(bins, [])) == list(range(8)) and max(loads) == 3600  # the long file sits alone
    assert shard.assign_files([], 4) == [[], [], [], []]
    assert shard.assign_files([5.0], 3) == [[0], [], []]


def test_split_chunk_range():
    shard = _load_shard()
    assert shard.split_chunk_range(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard.split_chunk_range(3591, 8)[-1][1] == 3591 and len(shard.split_chunk_range(3591, 8)) == 8
    assert shard.split_chunk_range(2, 8) == [(0, 1), (1, 2)]
    r = shard.split_chunk_range(591, 4)
    assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and max(c1 - c0 for c0, c1 in r) - min(c1 - c0 for c0, c1 in r) <= 1


def test_gather_world_size_2(tmp_path):
    durations = [300.0, 120.0, 600.0, 60.0, 300.0]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, durations, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "rank0.txt").read() == "ok" and open(tmp_path / "rank1.txt").read() == "ok"
