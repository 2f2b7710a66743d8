#!/usr/bin/env python
"""What a cfg2 world looks like at a given episode age (tuning aid): alive cells, foods, masses, episode ages.
    python tools/aged_stats.py [age_steps] [arenas]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agarai_b200 as ab
import bench

age = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
W = bench.WORLDS["cfg2"]
env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=n, **W), device="cuda:0", seed=0, action_config=ab.ActionConfig(8, 1, True, True))
obs = env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1000)
L = env.layout
for t in range(age + 1):
    if t in (0, 200, 500, 1000, 1500, 2000, 3000, age):
        st = env.get_state().cpu().numpy()
        m = st[:, L.cell_m:L.cell_m + L.cells_total]
        alive = (m > 0).sum(1)
        per_player = (m.reshape(n, -1, 16) > 0).sum(2)
        foods = (st[:, L.food_x:L.food_x + W["max_foods"]] >= 0).sum(1)
        ep = st[:, L.episode_steps].view(np.int32)
        timers = st[:, L.cell_t:L.cell_t + L.cells_total]
        print(f"t={t}: alive mean {alive.mean():.1f} p50 {np.median(alive):.0f} p90 {np.percentile(alive, 90):.0f} p99 {np.percentile(alive, 99):.0f} max {alive.max()} | "
              f">32: {(alive > 32).mean() * 100:.1f}% >64: {(alive > 64).mean() * 100:.1f}% | multi-cell players/arena {(per_player > 1).sum(1).mean():.2f} | "
              f"foods mean {foods.mean():.1f} any {(foods > 0).mean() * 100:.0f}% | max mass p50 {np.median(m.max(1)):.0f} p99 {np.percentile(m.max(1), 99):.0f} | "
              f"episode_steps mean {ep.mean():.0f} p10 {np.percentile(ep, 10):.0f} | cells with timer>0 {((timers > 0) & (m > 0)).sum(1).mean():.1f}", flush=True)
    idx = torch.randint(0, env.num_actions, (n, 1), generator=g, device="cuda", dtype=torch.int32)
    env.step_discrete(idx, out_obs=obs)
