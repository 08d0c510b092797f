"""Turn an .ncu-rep (brought back in gpurun_out/) into the small JSON summary that is committed under profiles/.
usage: python scripts/ncu_summarise.py gpurun_out/r02_fbank.ncu-rep profiles/r02_fbank_ncu_summary.json [note]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if k.endswith("bytes_read.sum") or k.endswith("bytes_write.sum"):
                    v *= UNIT_SCALE.get(u, 1.0)
                    u = "byte"
                if k == "gpu__time_duration.sum":
                    v *= UNIT_SCALE.get(u, 1.0)
                    u = "ms"
                d[k] = {"value": v, "unit": u}
        t = d.get("gpu__time_duration.sum", {}).get("value")
        rd, wr = d.get("dram__bytes_read.sum", {}).get("value", 0), d.get("dram__bytes_write.sum", {}).get("value", 0)
        if t:
            d["dram_gbs_under_ncu"] = (rd + wr) / (t / 1e3) / 1e9
        res.append(d)
    json.dump({"report": rep, "command": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1",
               "note": note, "launches": res}, open(out, "w"), indent=1)
    print("wrote", out, [x["kernel"][:40] for x in res])


if __name__ == "__main__":
    main()
