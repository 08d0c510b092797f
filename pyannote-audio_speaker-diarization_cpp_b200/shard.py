"""File-level sharding of a batch across the GPUs of one box, and the gather of per-file results.

The hot path has no exchange step: files are independent units, clustering is per file on one GPU.
Ranks therefore take whole files (longest-processing-time-first bins so the longest file does not straggle) and
the only collective is one all_gather of the padded int32 label arrays at the end (NCCL on GPUs, gloo in the CPU
tests).  Nothing here touches audio data.
"""
import numpy as np


def assign_files(durations, world_size):
    """Longest-processing-time-first assignment.  Returns a list (per rank) of file indices, each sorted."""
    order = sorted(range(len(durations)), key=lambda i: (-float(durations[i]), i))
    load = [0.0] * world_size
    bins = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += float(durations[i])
    return [sorted(b) for b in bins]


def pack_results(file_ids, label_arrays, max_len):
    """[n_local, 2 + max_len] int32: (file id, length, labels..., -1 padding)."""
    out = np.full((len(file_ids), 2 + max_len), -1, np.int32)
    for row, (fid, lab) in enumerate(zip(file_ids, label_arrays)):
        lab = np.asarray(lab, np.int32).ravel()
        assert lab.size <= max_len
        out[row, 0] = fid
        out[row, 1] = lab.size
        out[row, 2:2 + lab.size] = lab
    return out


def gather_results(local_packed, max_files_per_rank, device="cpu"):
    """all_gather of every rank's packed results (padded to max_files_per_rank rows); returns {file id: labels}
    on every rank.  Requires an initialised torch.distributed process group (nccl or gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    width = local_packed.shape[1]
    buf = np.full((max_files_per_rank, width), -1, np.int32)
    buf[:local_packed.shape[0]] = local_packed
    t = torch.from_numpy(buf).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    res = {}
    for o in outs:
        a = o.cpu().numpy()
        for row in a:
            if row[0] >= 0:
                res[int(row[0])] = row[2:2 + int(row[1])].copy()
    return res


def split_chunk_range(num_chunks, parts):
    """Chunk-range split of ONE long file (the per-chunk loops of speakerDiarization(),
    speakerDiarizer.cpp:3047-3105, have no dependence between chunks): contiguous ranges [c0, c1) of nearly equal size.
    The STFT items and the binarize rows of a range are independent of the other ranges, so each range can be
    submitted as its own sd_file (STFT + binarize stages only) on any GPU; speaker_count / clustering / the
    diarization aggregate need every chunk and stay on one GPU (they are the cheap or the sequential stages)."""
    parts = max(1, min(int(parts), int(num_chunks)))
    base, extra = divmod(int(num_chunks), parts)
    out, c0 = [], 0
    for p in range(parts):
        c1 = c0 + base + (1 if p < extra else 0)
        out.append((c0, c1))
        c0 = c1
    return out
