"""Device-side timeline of a batch (SDB_BATCH_TRACE, csrc/batch.cu): where the step time of bench.py goes.

usage: python scripts/prof_batch_timeline.py trace.csv [files_in_flight]
Prints, for the steady-state part of the trace (device-pointer files only), per-file stage durations, how much of the
wall time at least one STFT / at least one merge loop was running, and the concurrency of both over time."""
import csv
import sys

import numpy as np


def union_length(iv):
    iv = sorted(iv)
    total, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                total += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        total += cur_b - cur_a
    return total


def main():
    path = sys.argv[1]
    rows = list(csv.DictReader(open(path)))
    t = {k: np.array([float(r[k]) for r in rows]) for k in rows[0] if k != "worker"}
    w = np.array([int(r["worker"]) for r in rows])
    # host-pointer (e2e) files have STFT phases of tens of ms: drop them
    dev = (t["stft_done"] - t["start"]) < 20.0
    for k in t:
        t[k] = t[k][dev]
    w = w[dev]
    n = len(w)
    t0, t1 = t["start"].min(), t["end"].max()
    print("files %d, workers %d, span %.2f ms -> %.3f ms per file" % (n, len(set(w)), t1 - t0, (t1 - t0) / n))
    stft = t["stft_done"] - t["start"]
    pre = t["merge_begin"] - t["count_done"]
    merge = t["merge_end"] - t["merge_begin"]
    post = t["cluster_done"] - t["merge_end"]
    tail = t["end"] - t["cluster_done"]
    for name, v in (("stft (incl. queueing for SMs)", stft), ("binarize+count", t["count_done"] - t["stft_done"]),
                    ("clustering before merge loop", pre), ("merge loop", merge), ("clustering after merge loop", post),
                    ("aggregate", tail), ("whole file", t["end"] - t["start"])):
        print("  %-32s mean %7.3f  p10 %7.3f  p50 %7.3f  p90 %7.3f  max %7.3f ms"
              % (name, v.mean(), np.percentile(v, 10), np.percentile(v, 50), np.percentile(v, 90), v.max()))
    # gaps between consecutive files of a worker (host turnaround)
    gaps = []
    for k in set(w):
        idx = np.where(w == k)[0]
        idx = idx[np.argsort(t["start"][idx])]
        gaps += list(t["start"][idx][1:] - t["end"][idx][:-1])
    gaps = np.array(gaps)
    print("  gap between files of a worker    mean %7.3f  p90 %7.3f  max %7.3f ms" % (gaps.mean(), np.percentile(gaps, 90), gaps.max()))
    span = t1 - t0
    u_stft = union_length(list(zip(t["start"], t["stft_done"])))
    u_merge = union_length(list(zip(t["merge_begin"], t["merge_end"])))
    print("  some STFT pending or running: %.1f %% of the span; some merge loop running: %.1f %%" % (100 * u_stft / span, 100 * u_merge / span))
    # concurrency histogram sampled every 50 us
    ts = np.arange(t0, t1, 0.05)
    c_stft = ((t["start"][None, :] <= ts[:, None]) & (ts[:, None] < t["stft_done"][None, :])).sum(1)
    c_merge = ((t["merge_begin"][None, :] <= ts[:, None]) & (ts[:, None] < t["merge_end"][None, :])).sum(1)
    print("  mean STFTs pending/running %.2f, mean merge loops running %.2f" % (c_stft.mean(), c_merge.mean()))
    print("  time with no STFT pending: %.1f %%; with no merge loop: %.1f %%; with neither: %.1f %%"
          % (100 * (c_stft == 0).mean(), 100 * (c_merge == 0).mean(), 100 * ((c_stft == 0) & (c_merge == 0)).mean()))
    for lo, hi in ((0, 0), (1, 4), (5, 8), (9, 12), (13, 64)):
        m = (c_merge >= lo) & (c_merge <= hi)
        if m.any():
            print("    merge loops %2d..%2d: %5.1f %% of time, STFTs pending there %.2f" % (lo, hi, 100 * m.mean(), c_stft[m].mean()))
    # STFT service rate: completions per ms while at least one is pending
    print("  STFT completions: %d in %.2f ms pending time -> %.3f ms per STFT while any is pending (alone: 0.93)"
          % (n, u_stft, u_stft / n))
    if len(sys.argv) > 2 and sys.argv[2] == "dump":
        order = np.argsort(t["start"])
        for i in order[:64]:
            print("   w%02d start %8.3f stft %6.3f | merge %8.3f..%8.3f (%.3f) | end %8.3f"
                  % (w[i], t["start"][i] - t0, stft[i], t["merge_begin"][i] - t0, t["merge_end"][i] - t0, merge[i], t["end"][i] - t0))


if __name__ == "__main__":
    main()
