#!/usr/bin/env python
"""Where the host-buffer step (`e2e`) spends its time when several GPUs of one box step at once.
One process per GPU (torchrun).  Per rank: CPU affinity and the GPU's NUMA node, device-to-host bandwidth of a dense
observation buffer alone (all ranks copying at once), the step kernel alone, agar_step_host with the dense image copy
(AGAR_HOST_DENSE=1) and with the compact copy (occupied-bin lists + host threads).  Rank 0 prints one JSON line.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_diag.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402
from agarai_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, W = 4096, bench.WORLDS["cfg2"]
    info = {"rank": rank, "cpus": len(os.sched_getaffinity(0))}
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        info["gpu_numa_node"] = int(open(path).read()) if os.path.exists(path) else None
    except Exception:
        info["gpu_numa_node"] = None

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def mx(x):
        return sharding.max_over_ranks(x, dev)

    def run(mode):
        os.environ.pop("AGAR_HOST_DENSE", None)
        if mode == "dense":
            os.environ["AGAR_HOST_DENSE"] = "1"
        shard = sharding.shard_from_env(N)
        env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=N, arena_id_base=shard.arena_id_base, **W), device=dev, seed=0,
                                  action_config=ab.ActionConfig(8, 1, True, True))
        obs = env.reset()
        g = torch.Generator(device=dev); g.manual_seed(rank)
        for _ in range(300):
            env.step_discrete(torch.randint(0, env.num_actions, (N, 1), generator=g, device=dev, dtype=torch.int32), out_obs=obs)
        xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
        idx = np.random.default_rng(rank).integers(0, len(act_tab), (12, N, 1))
        h_xy = torch.from_numpy(np.ascontiguousarray(xy_tab[idx])).pin_memory(); h_act = torch.from_numpy(np.ascontiguousarray(act_tab[idx])).pin_memory()
        h_obs = torch.empty(env.obs_shape, dtype=torch.float32).pin_memory()
        h_r = torch.empty((N, 1), dtype=torch.float32).pin_memory(); h_d = torch.empty((N, 1), dtype=torch.uint8).pin_memory()
        out = {}
        if mode == "dense":  # the copy alone and the kernel alone, all ranks at once
            sync_all(); t0 = time.perf_counter()
            for _ in range(10):
                h_obs.copy_(obs, non_blocking=True)
            torch.cuda.synchronize(dev); dt = time.perf_counter() - t0
            out["d2h_only_gbs_per_gpu"] = 10 * h_obs.numel() * 4 / dt / 1e9
            out["d2h_only_gbs_all_gpus"] = world * 10 * h_obs.numel() * 4 / mx(dt) / 1e9
            sync_all(); t0 = time.perf_counter()
            for i in range(10):
                env.step_discrete(torch.from_numpy(idx[i].astype(np.int32)).to(dev), out_obs=obs)
            torch.cuda.synchronize(dev)
            out["kernel_only_ms"] = mx(time.perf_counter() - t0) / 10 * 1e3
        env.step_host(h_xy[0].numpy(), h_act[0].numpy(), h_obs.numpy(), h_r.numpy(), h_d.numpy())
        sync_all(); t0 = time.perf_counter()
        for i in range(1, 11):
            env.step_host(h_xy[i].numpy(), h_act[i].numpy(), h_obs.numpy(), h_r.numpy(), h_d.numpy())
        dt = time.perf_counter() - t0
        out["e2e_env_steps_per_s_this_gpu"] = 10 * N / dt
        out["e2e_env_steps_per_s_all_gpus"] = world * 10 * N / mx(dt)
        out["host_delivery_gbs_all_gpus"] = world * 10 * h_obs.numel() * 4 / mx(dt) / 1e9
        out["pcie_d2h_bytes_per_step"] = env.host_copy_bytes()[1]
        env.close()
        return out

    res = {"dense_copy": run("dense"), "compact_copy": run("compact")}
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"gpus": world, "arenas_per_gpu": N, "host_threads_per_rank": max(1, min(32, info["cpus"] // world)), "ranks": infos, **res}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
