#!/usr/bin/env python
"""Per-CUDA-line hot spots from an ncu report (needs -lineinfo and --import-source on).
    python profiles/hotlines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows:
    if len(r) > ii and r[0].isdigit():
        try:
            data.append((int(r[si]), int(r[ii]), int(r[0]), r[1].strip()))
        except ValueError:
            pass
ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print(f"total stall samples {ts}, warp instructions {ti}")
print("  %samples  %instr  line  source")
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"  {100 * d[0] / ts:7.1f}  {100 * d[1] / ti:6.1f}  {d[2]:4d}  {d[3][:100]}")
