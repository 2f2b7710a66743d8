#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize.py <tag> [--rep gpurun_out/prof.ncu-rep] [--launches gpurun_out/launches.csv]
"""
import argparse
import csv
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    out.write("launch list (gpu__time_duration.sum, ns; cold-cache serialised: compare shares)\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.write(f"  {100 * sum(v) / tot:5.1f}%  n={len(v):4d}  mean={sum(v) / len(v) / 1e3:9.1f} us  {k[:110]}\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    out.write(f"ncu --set full: {len(rows) - 2} launches of {rows[2][name_i][:80]}\n")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.write(f"  {m:85s} {rows[1][i]:>14s}  {[r[i] for r in rows[2:]]}\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    with open(f"profiles/{a.tag}.txt", "w") as f:
        if a.note:
            f.write(a.note + "\n\n")
        if a.launches:
            launches(a.launches, f)
            f.write("\n")
        if a.rep:
            full(a.rep, f)
    sys.stdout.write(open(f"profiles/{a.tag}.txt").read())
