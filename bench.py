#!/usr/bin/env python
"""bench.py — env-steps/s including the grid observation (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA, C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[1]): 4096 arenas per GPU, 1 agent + 25 bots + 25 viruses, 1000
pellets, arena 500, ticks_per_step 4, observation 3x64x64 float32, random policy over the PPO
action set with split and feed.  One "step" = one env.step of every arena on every GPU = one
kernel launch per GPU.  Arenas never interact, so N GPUs = N independent shards (weak scaling,
no collective on the data path); RNG is keyed by global arena id.

`value`   device-resident: action indices already in HBM, observations written to HBM.  The steps follow a reset,
          and an arena-step gets dearer as episodes age (cells grow, viruses pop them into many pieces): the default
          200 steps give ~30 M env-steps/s on one B200, `--steps 2000` ~18.5 M.
`e2e`     the same step through agar_step_host: actions from pinned host memory, observations,
          rewards and dones copied back to pinned host memory inside the timed region.
`roofline` algorithmic bytes per env-step x arenas per launch / launch duration.
`cpu_baseline` the CPU oracle (a port: the reference's simulator is an absent external package)
          on this box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARENAS_PER_GPU = 4096
WORLDS = {
    # BASELINE.json configs[1] (the headline): viruses + bot opponents (a2c/train.py:19-27 values)
    "cfg2": dict(num_agents=1, num_bots=25, num_viruses=25, num_pellets=1000, arena_size=500.0, ticks_per_step=4,
                 grid_size=64, obs_flags=7, pellet_regen=1, auto_reset=1, max_foods=64),
    # configs[0]/[2]/[4] world: single agent, pellets only (ppo.textproto:6-12 with num_agents 1)
    "cfg1": dict(num_agents=1, num_bots=0, num_viruses=0, num_pellets=1000, arena_size=1000.0, ticks_per_step=4,
                 grid_size=64, obs_flags=7, pellet_regen=1, auto_reset=1, max_foods=64),
    # configs[3] world: 16 agents per arena with split/merge (ppo.textproto:4-23)
    "cfg4": dict(num_agents=16, num_bots=0, num_viruses=0, num_pellets=1000, arena_size=1000.0, ticks_per_step=4,
                 grid_size=64, obs_flags=7, pellet_regen=1, auto_reset=1, max_foods=64),
    # the reference's own default worlds (a2c/hyperparameters.py:89-107, a2c/train.py:17-34, dqn/hyperparameters.py:23-28)
    "a2c": dict(num_agents=32, num_bots=0, num_viruses=25, num_pellets=1000, arena_size=500.0, ticks_per_step=4,
                grid_size=32, obs_flags=15, pellet_regen=1, auto_reset=1, max_foods=64),
    "a2c_test": dict(num_agents=8, num_bots=25, num_viruses=25, num_pellets=1000, arena_size=500.0, ticks_per_step=4,
                     grid_size=32, obs_flags=15, pellet_regen=1, auto_reset=1, max_foods=64),
    "dqn": dict(num_agents=1, num_bots=0, num_viruses=0, num_pellets=30, arena_size=45.0, ticks_per_step=4,
                grid_size=128, obs_flags=15, pellet_regen=1, auto_reset=1, max_foods=64, view_size=30.0),
}
WORLD = WORLDS["cfg2"]
METRIC = "env_steps_per_sec_incl_grid_obs"
UNIT = "env-steps/s"


def algorithmic_bytes_per_env_step(w=None):
    """compulsory HBM traffic of one env step (state touched once, obs written)."""
    w = w or WORLD
    P, V, F, A = w["num_pellets"], w["num_viruses"], w["max_foods"], w["num_agents"]
    K = A + w["num_bots"]
    C = bin(w["obs_flags"]).count("1")
    G = w["grid_size"]
    return 8 * P + 12 * V + 2 * 24 * K * 16 + 2 * 16 * F + A * (12 + 5) + 32 + A * C * G * G * 4


def ncu_traffic_bytes(world, arenas):
    """DRAM bytes of one step-kernel launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[world]
        if int(t["arenas"]) != int(arenas):
            return None
        return int(t["dram_read_bytes"]) + int(t["dram_write_bytes"])
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms from before the warm-up; the samples whose time
    stamp falls inside the timed region (or, for a very short region, the nearest ones) are reported."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [(t, r) for t, r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        inside = [r for t, r in rows if self.t0 is not None and self.t0 - 0.02 <= t <= self.t1 + 0.03]
        window = "timed region"
        if not inside and rows and self.t0 is not None:  # region shorter than the sampling period: nearest samples
            mid = 0.5 * (self.t0 + self.t1)
            inside = [r for _, r in sorted(rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            window = "nearest to the timed region"
        sm = sorted(float(r[2]) for r in inside)
        mx = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for name, val in zip(names, r[6:10]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_arm(arenas, steps, warmup, nthreads, seed=0):
    """Times the CPU oracle (oracle/agar_oracle.c) on `arenas` arenas of the bench world."""
    import numpy as np
    from oracle import oracle as O
    import agarai_b200.actions as act_mod
    cfg = O.default_config(num_arenas=arenas, **WORLD)
    env = O.OracleEnv(cfg, seed=seed, nthreads=nthreads)
    obs = env.reset()
    xy_tab, act_tab = act_mod.ppo_action_table(act_mod.ActionConfig(8, 1, True, True))
    rng = np.random.default_rng(seed)
    times = []
    for i in range(warmup + steps):
        idx = rng.integers(0, len(act_tab), size=(arenas, WORLD["num_agents"]))
        t0 = time.perf_counter()
        env.step(xy_tab[idx], act_tab[idx], out=obs)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    env.close()
    return arenas * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference's own
    simulator (gym_agario) is an external package that is not available, so this arm runs the
    oracle port, multi-threaded over arenas (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 512
    value, sec = cpu_arm(sample, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.world}: {WORLD['num_agents']} agent(s) + {WORLD['num_bots']} bots + "
                               f"{WORLD['num_viruses']} viruses, {WORLD['num_pellets']} pellets, arena "
                               f"{WORLD['arena_size']:g}, obs {bin(WORLD['obs_flags']).count('1')}x{WORLD['grid_size']}x{WORLD['grid_size']} f32 per agent",
                   "arenas_per_step": sample, **WORLD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} arenas x {args.steps} steps (of the {ARENAS_PER_GPU}-arena workload), "
                                   "oracle/agar_oracle.c with one pthread per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import agarai_b200 as ab
    from agarai_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N = args.arenas
    ac = ab.ActionConfig(8, 1, True, True)
    shard = sharding.shard_from_env(N)  # rank r owns global arenas [r*N, (r+1)*N): no data-path collective
    cfg = ab.default_config(num_arenas=N, arena_id_base=shard.arena_id_base, **WORLD)
    env = ab.BatchedAgarioEnv(cfg, device=dev, seed=args.seed, action_config=ac)
    K, W = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device=dev); gen.manual_seed(args.seed + rank)
    A = WORLD["num_agents"]
    idx = torch.randint(0, env.num_actions, (K + W, N, A), generator=gen, device=dev, dtype=torch.int32)
    obs = [torch.empty(env.obs_shape, dtype=torch.float32, device=dev) for _ in range(2)]
    env.reset(out_obs=obs[0])
    if args.age > 0:  # fast-forward every arena to a fixed episode age with the same random policy (untimed)
        ga = torch.Generator(device=dev); ga.manual_seed(args.seed + 1000 + rank)
        for t in range(args.age):
            ia = torch.randint(0, env.num_actions, (N, A), generator=ga, device=dev, dtype=torch.int32)
            env.step_discrete(ia, out_obs=obs[t & 1], want_obs=not args.no_obs)
        torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(W):
        env.step_discrete(idx[i], out_obs=obs[i & 1], want_obs=not args.no_obs)
    barrier()
    sampler.mark_begin()
    l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        _, rew, done, _ = env.step_discrete(idx[W + i], out_obs=obs[i & 1], want_obs=not args.no_obs)
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_max = sharding.max_over_ranks(ms, dev)
    value = world * N * K / (ms_max * 1e-3)

    # ---- e2e: host buffers through agar_step_host (H2D actions, D2H obs/reward/done inside) ----
    xy_tab, act_tab = ab.ppo_action_table(ac)
    Ke = max(2, min(K, args.e2e_steps))
    h_idx = idx[W:W + Ke].cpu().numpy()
    h_xy = torch.from_numpy(np.ascontiguousarray(xy_tab[h_idx])).pin_memory()   # [Ke,N,A,2]
    h_act = torch.from_numpy(np.ascontiguousarray(act_tab[h_idx])).pin_memory()  # [Ke,N,A]
    h_obs = torch.empty(env.obs_shape, dtype=torch.float32).pin_memory()
    h_rew = torch.empty((N, A), dtype=torch.float32).pin_memory()
    h_done = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    env.step_host(h_xy[0].numpy(), h_act[0].numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())  # warm-up
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        env.step_host(h_xy[i].numpy(), h_act[i].numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())
    torch.cuda.synchronize(dev)
    e2e_value = world * N * Ke / sharding.max_over_ranks(time.perf_counter() - t0, dev)
    h2d = N * A * (8 + 4)
    d2h = h_obs.numel() * 4 + N * A * (4 + 1)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        B = algorithmic_bytes_per_env_step()
        launch_s = (ms_max * 1e-3) / max(launches, 1)
        achieved = B * N / launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.world}: {N} arenas/GPU, {A} agent(s) + {WORLD['num_bots']} bots + "
                                   f"{WORLD['num_viruses']} viruses, {WORLD['num_pellets']} pellets, arena "
                                   f"{WORLD['arena_size']:g}, obs {bin(WORLD['obs_flags']).count('1')}x{WORLD['grid_size']}x{WORLD['grid_size']} f32 per agent, random policy over 25 actions "
                                   "(8 dirs, split, feed)",
                       "arenas_per_gpu": N,
                       "l2": f"no flush: each step streams {env.layout.words * 4 * N / 1e6:.0f} MB of state and writes "
                             f"a {obs[0].numel() * 4 / 1e6:.0f} MB observation buffer (2 alternate); L2 is 126 MB",
                       **WORLD},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": "agar_step_host (pinned host buffers; obs D2H every step, 4 chunks: copy of one overlaps the kernel of the next)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(args.world, N) if not args.set else None, "kernel": "agar_step_kernel", "bytes_per_env_step": B,
                         "env_steps_per_launch": N, "launch_us": launch_s * 1e6, "peak_source": peak_src},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = 512
            _, sec = cpu_arm(sample, 1, 1, cores)        # size the sample for ~10-20 s of CPU work
            steps_cpu = int(max(3, min(200, 12.0 / max(sec, 1e-3))))
            v, sec = cpu_arm(sample, steps_cpu, 1, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} arenas x {steps_cpu} steps of the same world, "
                                              "oracle/agar_oracle.c, one pthread per core"}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--arenas", type=int, default=ARENAS_PER_GPU, help="arenas per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--age", type=int, default=0, help="untimed steps run before the warm-up (episode age of the timed region)")
    ap.add_argument("--no-obs", action="store_true", help="tuning only: step without the observation raster")
    ap.add_argument("--world", default="cfg2", choices=sorted(WORLDS), help="BASELINE.json world (default: the headline cfg2)")
    ap.add_argument("--set", action="append", default=[], metavar="FIELD=VALUE",
                    help="tuning experiments only: override a world field (e.g. ticks_per_step=8); the line says so in config")
    args = ap.parse_args()
    global WORLD
    WORLD = dict(WORLDS[args.world])
    for kv in args.set:
        k, v = kv.split("=")
        WORLD[k] = type(WORLDS[args.world].get(k, 0))(float(v))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
