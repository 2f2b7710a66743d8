#!/usr/bin/env python
"""Cycles per phase of agar_step_kernel from a -DAGAR_PHASE_CLOCKS build (tools/build_variant.sh WORK clocks -DAGAR_PHASE_CLOCKS):
    AGAR_B200_LIB=$PWD/agarai_b200/_ab/clocks.so python tools/phase_clocks.py [arenas] [world] [FIELD=VALUE ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agarai_b200 as ab
import bench
from agarai_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
W = dict(bench.WORLDS[sys.argv[2] if len(sys.argv) > 2 else "cfg2"])
AGE = 0
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    if k == "AGE":
        AGE = int(v); continue
    W[k] = type(W[k])(float(v))
env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=n, **W), device="cuda:0", seed=0, action_config=ab.ActionConfig(8, 1, True, True))
obs = env.reset()
idx = torch.randint(0, env.num_actions, (120, n, W["num_agents"]), device="cuda", dtype=torch.int32)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for t in range(AGE):
    env.step_discrete(torch.randint(0, env.num_actions, (n, W["num_agents"]), generator=g, device="cuda", dtype=torch.int32), out_obs=obs)
for i in range(20):
    env.step_discrete(idx[i], out_obs=obs)
L = _lib.lib()
buf = (C.c_ulonglong * 32)()
L.agar_debug_phase_clocks(buf)
steps = 100
for i in range(steps):
    env.step_discrete(idx[20 + i], out_obs=obs)
L.agar_debug_phase_clocks(buf)
names = ["staging", "hash check/build + list", "players, actions", "bots", "split, feed", "in-tick rebuild (+merge of earlier ticks)",
         "tick: motion + query", "tick: resolve, foods", "tick: apply gains", "tick: eat test", "tick: eat apply", "merge (last tick)",
         "rewards, auto reset", "persist hash", "write back", "raster: count channels", "bulk store wait",
         "raster: setup + centroids", "raster: barrier + gather others", "raster: base image stores", "raster: barrier after stores",
         "raster: mass patches (warp 0)"]
tot = sum(buf)
print(f"{n} arenas, {W['ticks_per_step']} ticks/step: {tot / steps / n:.0f} cycles per arena-step (thread 0's clock)")
for i, nm in enumerate(names):
    print(f"  {nm:45s} {buf[i] / steps / n:9.0f}  {100 * buf[i] / tot:5.1f}%")
