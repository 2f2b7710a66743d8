#!/usr/bin/env python
"""The one collective of the path (BASELINE north star): PPO gradient all-reduce over NCCL.  One process per GPU, each
with its own arena shard (global arena ids keyed by rank), PPOTrainer(ddp=True); after every episode the ranks'
parameters must be identical.  One JSON line from rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/ppo_ddp.py --envs 4096 --steps 32 --episodes 3"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402
from agarai_b200 import sharding  # noqa: E402
from agarai_b200.ppo import PPOConfig, PPOTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="arenas per GPU")
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--world", default="cfg1")
    ap.add_argument("--amp", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard = sharding.shard_from_env(a.envs)
    W = bench.WORLDS[a.world]
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=a.envs, arena_id_base=shard.arena_id_base, **W), device=dev, seed=0,
                              action_config=ab.ActionConfig(8, 1, False, False))
    tr = PPOTrainer(env, PPOConfig(episode_length=a.steps, amp_bf16=a.amp), seed=0, ddp=world > 1)
    times, stats = [], {}
    for ep in range(a.episodes):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        stats = tr.train_episode()
        torch.cuda.synchronize(dev)
        times.append(sharding.max_over_ranks(time.perf_counter() - t0, dev))
        # identical parameters on every rank: compare each rank's flat parameter vector with rank 0's
        flat = torch.cat([p.detach().reshape(-1) for p in tr.model.parameters()])
        if world > 1:
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            same = torch.tensor([int(torch.equal(ref, flat))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            assert int(same.item()) == 1, f"parameters diverged across ranks after episode {ep}"
    if rank == 0:
        n = world * a.envs * a.steps
        print(json.dumps({"workload": f"PPO with DDP gradient all-reduce (NCCL), {world} rank(s) x {a.envs} arenas x {a.steps}-step rollout, {a.world} world",
                          "ranks": world, "episode_s": times[-1], "env_steps_per_s": n / times[-1], "sgd_steps": tr.sgd_steps,
                          "params_identical_across_ranks": True, "grad_bytes_per_sgd_step": int(sum(p.numel() for p in tr.model.parameters()) * 4),
                          "stats": stats}), flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
