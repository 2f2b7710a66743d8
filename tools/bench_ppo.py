#!/usr/bin/env python
"""BASELINE configs[2]: PPO training loop, N envs x T-step rollout with GAE on the GPU, 1 B200.
Times one episode of agarai_b200.ppo.PPOTrainer (policy forward in PyTorch, everything else in the CUDA
library): rollout collection, advantages (one agar_gae launch), the SGD phase (state snapshots re-rasterised
per minibatch).  One JSON line.
    python tools/bench_ppo.py --envs 65536 --steps 128 [--policy-chunk 8192]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402
from agarai_b200.ppo import PPOConfig, PPOTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--policy-chunk", type=int, default=8192, help="policy forward batch (bounds activation memory)")
    ap.add_argument("--episodes", type=int, default=2)
    ap.add_argument("--amp", action="store_true", help="policy under bf16 autocast + channels_last")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="let cuDNN pick the convolution algorithms by timing")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = bool(a.cudnn_benchmark)
    W = bench.WORLDS["cfg1"]
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=a.envs, **W), device=dev, seed=0, action_config=ab.ActionConfig(8, 1, False, False))
    need = a.envs * a.steps * env.layout.words * 4 + a.envs * 3 * 64 * 64 * 4 * 3
    free, total = torch.cuda.mem_get_info(dev)
    if need > 0.8 * free:
        raise SystemExit(f"needs ~{need / 1e9:.0f} GB of {free / 1e9:.0f} GB free: lower --envs")
    tr = PPOTrainer(env, PPOConfig(episode_length=a.steps, amp_bf16=a.amp), seed=0)
    net = tr.net

    def chunked(x):  # same model, evaluated in slices so that conv activations stay bounded
        outs = [net(x[i:i + a.policy_chunk]) for i in range(0, x.shape[0], a.policy_chunk)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
    tr.net = chunked
    res = []
    for ep in range(a.episodes):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tr.collect()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        adv, ret = tr.advantages()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        tr.net = net
        stats = tr.update(adv, ret)
        tr.net = chunked
        torch.cuda.synchronize(); t3 = time.perf_counter()
        res.append((t1 - t0, t2 - t1, t3 - t2, stats))
    c, g, u, stats = res[-1]
    n = a.envs * a.steps
    print(json.dumps({"workload": f"configs[2]: PPO, {a.envs} envs x {a.steps}-step rollout, cfg1 world, 9 actions",
                      "collect_s": c, "gae_s": g, "update_s": u, "env_steps_per_s_collect": n / c,
                      "env_steps_per_s_episode": n / (c + g + u), "gae_us": g * 1e6,
                      "policy": "bf16 autocast" if a.amp else "fp32", "snapshot_gb": a.envs * a.steps * env.layout.words * 4 / 1e9,
                      "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "stats": stats}))


if __name__ == "__main__":
    main()
