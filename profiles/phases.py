#!/usr/bin/env python
"""Per-phase share of warp instructions and stall samples of agar_step_kernel, from an ncu
report captured with --import-source on (kernel built with -lineinfo).  Phases are delimited
by the marker comments in agarai_b200/csrc/agar_kernels.cuh.
    python profiles/phases.py gpurun_out/prof.ncu-rep <arenas per launch>"""
import csv
import subprocess
import sys

rep, arenas = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 4096.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = {}
cur_file = None
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1]
    if len(r) > ii and r[0].isdigit() and cur_file and cur_file.endswith("agar_kernels.cuh"):
        try:
            s0, i0 = data.get(int(r[0]), (0, 0))
            data[int(r[0])] = (s0 + int(r[si]), i0 + int(r[ii]))
        except ValueError:
            pass
src = open("agarai_b200/csrc/agar_kernels.cuh").read().split("\n")
MARKS = [
    ("device helpers (philox, sums, TMA)", "namespace agar {"),
    ("reset_arena (out of line)", "__device__ __noinline__ void reset_arena"),
    ("build_hash (out of line)", "__device__ __noinline__ void build_hash"),
    ("build_list", "__device__ __forceinline__ void build_list"),
    ("raster helpers (bins, ordered sums)", "__device__ __noinline__ float div_rn"),
    ("raster: centroids, setup", "__device__ void raster_arena"),
    ("raster: gather others", "// ---- gather: other players' cells in view"),
    ("raster: base image stores", "// ---- base image of every channel"),
    ("raster: mass patches (warp 0)", "// ---- warp 0: own cells and others"),
    ("raster: count channels", "// ---- count channels ----"),
    ("virus pops (out of line)", "__device__ __noinline__ void virus_pops"),
    ("kernel: prelude + stage record", "template <int kThreads, int kMinBlocks, class CV>"),
    ("step start: hash check, list", "// ---- pellet hash (persisted across steps"),
    ("step start: players", "// ---- players: bot respawn, centroid"),
    ("step start: multi-cell players", "// ---- players with several cells"),
    ("step start: agent actions", "// ---- agents: action -> target point"),
    ("step start: bots", "// ---- bots: nearest pellet in the square"),
    ("step start: split", "// ---- split: i-th eligible"),
    ("step start: feed", "// ---- feed: cells in gid order"),
    ("tick: motion + pellet query", "for (int tk = 0; tk < TICKS; ++tk)"),
    ("tick: resolve pellets, foods", "// ---- resolve the claimed pellets"),
    ("tick: eat test lambda + apply gains", "// cell-eats-cell test (different players"),
    ("tick: eat loop, eat apply", "// ---- cell eats cell (one code site)"),
    ("tick: merge", "// ---- merge own cells whose timers ran out"),
    ("rewards / dones, auto reset", "// ---- rewards / dones ----"),
    ("persist hash", "// ---- persist the pellet hash"),
    ("write back", "// ---- write back:"),
    ("observations call", "// ---- observations ----"),
    ("gae kernel", "// rl.gae / rl.make_returns (rl/__init__.py:94-159"),
]
pos = []
start = 0
for name, key in MARKS:
    line = next((i + 1 for i in range(start, len(src)) if key in src[i]), None)
    if line is None:
        continue
    pos.append((name, line))
    start = line
pos.append(("end", 10 ** 9))
ts = sum(v[0] for v in data.values()) or 1
ti = sum(v[1] for v in data.values()) or 1
print(f"warp instructions per arena-step: {ti / arenas:.0f}   (stall samples {ts})")
print(f"{'phase':42s} {'%samples':>9s} {'%instr':>8s} {'instr/arena':>12s}")
for (name, lo), (_, hi) in zip(pos, pos[1:]):
    s_ = sum(v[0] for l, v in data.items() if lo <= l < hi)
    i_ = sum(v[1] for l, v in data.items() if lo <= l < hi)
    print(f"{name:42s} {100 * s_ / ts:9.1f} {100 * i_ / ti:8.1f} {i_ / arenas:12.0f}")
