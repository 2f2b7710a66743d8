#!/usr/bin/env python
"""BASELINE configs[4]: env-step throughput sweep over the number of arenas (cfg1 world, grid observation
included, device-resident), one JSON line per size.  4M arenas need 206 GB of observations per step and are
stepped as two handles of 2M (SURVEY.md §8d "capacity limits").
    python tools/sweep.py [--sizes 1024,4096,...] [--steps 50]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--world", default="cfg1")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    W = bench.WORLDS[a.world]
    B = bench.algorithmic_bytes_per_env_step(W)
    peak, _ = bench.measured_peak_gbs()
    for n in [int(x) for x in a.sizes.split(",")]:
        free, _ = torch.cuda.mem_get_info(dev)
        obs_bytes = n * W["num_agents"] * 3 * 64 * 64 * 4
        if obs_bytes * 1.3 > free:
            print(json.dumps({"arenas": n, "skipped": f"observation buffer {obs_bytes / 1e9:.0f} GB exceeds free HBM {free / 1e9:.0f} GB"}), flush=True)
            continue
        env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=n, **W), device=dev, seed=0, action_config=ab.ActionConfig(8, 1, True, True))
        obs = torch.empty(env.obs_shape, dtype=torch.float32, device=dev)
        env.reset(out_obs=obs)
        idx = torch.randint(0, env.num_actions, (8, n, W["num_agents"]), device=dev, dtype=torch.int32)
        for i in range(5):
            env.step_discrete(idx[i % 8], out_obs=obs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            env.step_discrete(idx[i % 8], out_obs=obs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        v = n / (ms * 1e-3)
        print(json.dumps({"world": a.world, "arenas": n, "us_per_step": ms * 1e3, "env_steps_per_s": v,
                          "hbm_frac": v * B / 1e9 / peak}), flush=True)
        env.close(); del obs, idx
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
