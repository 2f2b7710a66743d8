#!/usr/bin/env python
"""Where agar_step_host spends its wall time on one GPU (tuning aid): cfg2 world, 4096 arenas aged by 300 steps, the
default delivery and the persistent-buffer delivery; the library prints its own breakdown (AGAR_HOST_PROFILE) on close.
    AGAR_HOST_THREADS=16 python tools/e2e_profile.py [arenas] [steps]"""
import os
import sys
import time

os.environ["AGAR_HOST_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402

torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))  # (under torchrun: one copy per GPU, all stepping at once)
WORLD = int(os.environ.get("WORLD_SIZE", 1))
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for persistent in (False, True):
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=N, **bench.WORLDS["cfg2"]), seed=0, action_config=ab.ActionConfig(8, 1, True, True))
    obs = env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for _ in range(int(os.environ.get("AGE", 300))):
        env.step_discrete(torch.randint(0, env.num_actions, (N, 1), generator=g, device="cuda", dtype=torch.int32), out_obs=obs)
    xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
    idx = np.random.default_rng(0).integers(0, len(act_tab), (K + 1, N, 1))
    h_xy = torch.from_numpy(np.ascontiguousarray(xy_tab[idx])).pin_memory(); h_act = torch.from_numpy(np.ascontiguousarray(act_tab[idx])).pin_memory()
    h_obs = torch.empty(env.obs_shape, dtype=torch.float32).pin_memory()
    h_r = torch.empty((N, 1), dtype=torch.float32).pin_memory(); h_d = torch.empty((N, 1), dtype=torch.uint8).pin_memory()
    env.host_obs_persistent(persistent)
    env.step_host(h_xy[0].numpy(), h_act[0].numpy(), h_obs.numpy(), h_r.numpy(), h_d.numpy())
    if WORLD > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(1, K + 1):
        env.step_host(h_xy[i].numpy(), h_act[i].numpy(), h_obs.numpy(), h_r.numpy(), h_d.numpy())
    dt = time.perf_counter() - t0
    if WORLD > 1:
        dist.barrier()
    print(f"rank {os.environ.get('RANK', 0)} persistent={persistent} threads={os.environ.get('AGAR_HOST_THREADS', 'default')}: {K * N / dt / 1e6:.2f} M env-steps/s, {dt / K * 1e3:.3f} ms per step", flush=True)
    sys.stdout.flush()
    env.close()
