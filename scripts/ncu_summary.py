"""Turn an .ncu-rep (ncu --set full) into the small JSON summary committed under profiles/."""
import csv
import json
import subprocess
import sys

rep, out, capture = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
launches, traffic = [], []
for r in rows[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    rec = {"kernel": d.get("Kernel Name", "")[:60]}
    for w in want:
        if w in d:
            rec[w] = f"{d[w]} {u.get(w, '')}".strip()
    launches.append(rec)
    try:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d["dram__bytes_read.sum"]) * scale.get(u["dram__bytes_read.sum"], 1)
        wr = float(d["dram__bytes_write.sum"]) * scale.get(u["dram__bytes_write.sum"], 1)
        traffic.append(rd + wr)
    except Exception:
        pass
summary = {"capture": capture, "launches": launches}
if traffic:
    summary["dram_traffic_bytes_per_launch"] = sum(traffic) / len(traffic)
json.dump(summary, open(out, "w"), indent=1)
print(out, len(launches), "launches", summary.get("dram_traffic_bytes_per_launch"))
