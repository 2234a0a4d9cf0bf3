#!/usr/bin/env python
"""Summarise an ncu launch list (+ optional full capture) into tracked text files under profiles/.
    python profiles/summarize.py <tag>      # reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
lines = []

launch_csv = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(launch_csv):
    rows = list(csv.reader(open(launch_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    lines.append(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n")
    lines.append(f"| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        lines.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {sum(v)/tot*100:.1f}% |")
    lines.append("")

rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
            "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    ki = hdr.index("Kernel Name")
    seen = {}
    lines.append(f"# ncu --set full `{tag}` (one representative launch per kernel)\n")
    for r in rows[2:]:
        name = r[ki].split("(")[0][:70]
        key = name
        if name.startswith("slots_round"):
            key = name + f"#{seen.get(name, 0)}"
            seen[name] = seen.get(name, 0) + 1
            if seen[name] > 2:
                continue
        elif key in seen:
            continue
        seen.setdefault(key, 1)
        lines.append(f"## `{key}`\n")
        lines.append("| metric | value | unit |\n|---|---:|---|")
        for w, i in idx:
            lines.append(f"| {w} | {r[i]} | {units[i]} |")
        lines.append("")

path = os.path.join(out_dir, f"{tag}_summary.md")
open(path, "w").write("\n".join(lines) + "\n")
print(open(path).read())
