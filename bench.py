#!/usr/bin/env python
"""bench.py — env-steps/s including the grid observation (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA, C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[1]): 4096 arenas per GPU, 1 agent + 25 bots + 25 viruses, 1000
pellets, arena 500, ticks_per_step 4, observation 3x64x64 float32, random policy over the PPO
action set with split and feed.  One "step" = one env.step of every arena on every GPU = one
kernel launch per GPU.  Arenas never interact, so N GPUs = N independent shards (weak scaling,
no collective on the data path); RNG is keyed by global arena id.

Episode age.  An arena-step gets dearer as episodes age (cells grow, viruses pop them into many pieces), and a
training run lives in the steady state, not in the first steps after a reset.  Both arms therefore fast-forward
every arena by `--age` untimed steps of the same random policy (default 1000: the alive-cell and mass
distributions are stationary from ~1000 steps on, profiles/r02_notes.txt) before the warm-up;
`config.episode_age_steps` says so, and the line also carries the post-reset figure as `young_world`.

`value`   device-resident: action indices already in HBM, observations written to HBM.
`e2e`     the same step through agar_step_host: actions from pinned host memory, observations,
          rewards and dones copied back to pinned host memory inside the timed region.
`roofline` algorithmic bytes per env-step x arenas per launch / launch duration.
`extras`  the other BASELINE configs on the same box and in the same run: cfg1-world and cfg4-world steps
          (configs[0]/[2]/[4] and configs[3]; cfg4 at 32,768 arenas per GPU = 262,144 arenas on 8 GPUs), `agar_gae` at
          [128, 65,536] (configs[2]) and the observe-only raster, each with its launch time and fraction of the HBM peak.
`cpu_baseline` the CPU oracle (a port: the reference's simulator is an absent external package)
          on this box's host cores, on a bounded sample of the same workload at the same episode age.
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
DEFAULT_AGE = 1000
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
    # the headline world with the optional mass-decay rule on (AgarConfig.mass_decay; 2e-4 per tick = 2 % per second at 25 steps/s)
    "cfg2_decay": dict(num_agents=1, num_bots=25, num_viruses=25, num_pellets=1000, arena_size=500.0, ticks_per_step=4,
                       grid_size=64, obs_flags=7, pellet_regen=1, auto_reset=1, max_foods=64, mass_decay=2e-4),
}
WORLD = WORLDS["cfg2"]
METRIC = "env_steps_per_sec_incl_grid_obs"
UNIT = "env-steps/s"
CFG4_ARENAS_PER_GPU = 32768  # BASELINE configs[3]: 262,144 arenas over 8 GPUs


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


def describe(w):
    return (f"{w['num_agents']} agent(s) + {w['num_bots']} bots + {w['num_viruses']} viruses, {w['num_pellets']} pellets, "
            f"arena {w['arena_size']:g}, obs {bin(w['obs_flags']).count('1')}x{w['grid_size']}x{w['grid_size']} f32 per agent"
            + (f", mass_decay {w['mass_decay']:g} per tick" if w.get("mass_decay") else ""))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms for the whole run.  A region is reported from the
    samples whose time stamp falls inside it; the caller makes every region at least ~100 ms of the same load (the
    untimed ageing / repeat steps count: same kernel, same arenas), so a region is never empty."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t_end = time.time() + 5.0  # nvidia-smi needs a moment before its first line
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def region(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        rows = [(t, r) for t, r in list(self.rows) if len(r) >= 10 and r[2].replace(".", "").isdigit()]
        inside = [r for t, r in rows if t0 - 0.01 <= t <= t1 + 0.01]
        window = "region under load"
        if not inside and rows:
            mid = 0.5 * (t0 + t1)
            inside = [r for _, r in sorted(rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            window = "nearest samples"
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

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def cpu_arm(arenas, steps, warmup, nthreads, age=0, seed=0):
    """Times the CPU oracle (oracle/agar_oracle.c) on `arenas` arenas of the bench world, aged by `age` untimed steps
    of the same random policy.  Returns (env-steps/s, seconds per step, seconds spent ageing)."""
    import numpy as np
    from oracle import oracle as O
    import agarai_b200.actions as act_mod
    cfg = O.default_config(num_arenas=arenas, **WORLD)
    env = O.OracleEnv(cfg, seed=seed, nthreads=nthreads)
    obs = env.reset()
    xy_tab, act_tab = act_mod.ppo_action_table(act_mod.ActionConfig(8, 1, True, True))
    rng = np.random.default_rng(seed)
    A = WORLD["num_agents"]
    t_age = time.perf_counter()
    for _ in range(age):
        idx = rng.integers(0, len(act_tab), size=(arenas, A))
        env.step(xy_tab[idx], act_tab[idx], want_obs=False)
    t_age = time.perf_counter() - t_age
    times = []
    for i in range(warmup + steps):
        idx = rng.integers(0, len(act_tab), size=(arenas, A))
        t0 = time.perf_counter()
        env.step(xy_tab[idx], act_tab[idx], out=obs)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    env.close()
    return arenas * len(times) / sum(times), sum(times) / len(times), t_age


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores, on the arm's own config: all
    4096 arenas, aged by the same number of untimed steps.  The reference's own simulator (gym_agario) is an
    external package that is not available, so this arm runs the oracle port, multi-threaded over arenas
    (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    N = args.arenas
    value, sec, t_age = cpu_arm(N, args.steps, args.warmup, cores, age=args.age)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.world}: {N} arenas, {describe(WORLD)}, random policy over 25 actions (8 dirs, split, feed)",
                   "arenas_per_step": N, "episode_age_steps": args.age, **WORLD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"all {N} arenas x {args.steps} steps after {args.age} untimed ageing steps ({t_age:.0f} s), "
                                   "oracle/agar_oracle.c with one pthread per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class _Ctx:
    """Per-process CUDA / distributed context of the CUDA arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, reps):
        """device time of `reps` calls of fn(i) between two barriers, max over ranks; returns (ms, wall t0, wall t1)"""
        from agarai_b200 import sharding
        torch = self.torch
        self.barrier()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        self.barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1), self.dev), t0, time.time()


def time_world(ctx, world_name, arenas, steps, warmup, age, seed, no_obs=False, want_young=False):
    """Steps one BASELINE world device-resident.  Returns a dict with ms per launch (max over ranks), env-steps/s of the
    whole job, the roofline fraction and the wall-clock window of the loaded region (for the clock sampler)."""
    import agarai_b200 as ab
    from agarai_b200 import sharding
    torch = ctx.torch
    w = WORLD if world_name is None else WORLDS[world_name]
    A = w["num_agents"]
    ac = ab.ActionConfig(8, 1, True, True)
    shard = sharding.shard_from_env(arenas)  # rank r owns global arenas [r*N, (r+1)*N): no data-path collective
    cfg = ab.default_config(num_arenas=arenas, arena_id_base=shard.arena_id_base, **w)
    env = ab.BatchedAgarioEnv(cfg, device=ctx.dev, seed=seed, action_config=ac)
    gen = torch.Generator(device=ctx.dev); gen.manual_seed(seed + ctx.rank)
    idx = torch.randint(0, env.num_actions, (steps + warmup, arenas, A), generator=gen, device=ctx.dev, dtype=torch.int32)
    obs = [torch.empty(env.obs_shape, dtype=torch.float32, device=ctx.dev) for _ in range(2)]
    env.reset(out_obs=obs[0])
    out = {"env": env, "idx": idx, "obs": obs}

    def step(i, base=0):
        env.step_discrete(idx[base + i], out_obs=obs[i & 1], want_obs=not no_obs)

    if want_young:  # the post-reset regime (what round 1 reported), before the ageing below
        for i in range(warmup):
            step(i)
        ms, _, _ = ctx.timed(lambda i: step(i, warmup), steps)
        out["young_ms_per_launch"] = ms / steps
    t_load0 = time.time()
    if age > 0:  # fast-forward every arena with the same random policy (untimed)
        ga = torch.Generator(device=ctx.dev); ga.manual_seed(seed + 1000 + ctx.rank)
        for t in range(age):
            ia = torch.randint(0, env.num_actions, (arenas, A), generator=ga, device=ctx.dev, dtype=torch.int32)
            env.step_discrete(ia, out_obs=obs[t & 1], want_obs=not no_obs)
    for i in range(warmup):
        step(i)
    l0 = env.launch_count
    ms, _, t1 = ctx.timed(lambda i: step(i, warmup), steps)
    out["launches"] = env.launch_count - l0
    # keep the same load running until the region holds a handful of clock samples (untimed)
    while time.time() - t_load0 < 0.15:
        for i in range(min(steps, 20)):
            step(i, warmup)
        torch.cuda.synchronize(ctx.dev)
        t1 = time.time()
    out.update(ms_per_launch=ms / steps, window=(t_load0, t1), world=w, arenas=arenas, steps=steps)
    return out


def roofline_of(w, arenas, ms_per_launch, peak):
    B = algorithmic_bytes_per_env_step(w)
    gbs = B * arenas / (ms_per_launch * 1e-3) / 1e9
    return B, gbs, gbs / peak


def run_extras(ctx, args, sampler, peak):
    """The other BASELINE configs, each timed like the headline (CUDA events, max over ranks)."""
    import agarai_b200 as ab
    torch = ctx.torch
    extras = {}
    K, W = max(args.steps, 10), max(args.warmup, 3)

    def world_entry(name, arenas, age, label):
        r = time_world(ctx, name, arenas, min(K, 50), W, age, args.seed)
        B, gbs, frac = roofline_of(r["world"], arenas, r["ms_per_launch"], peak)
        A = r["world"]["num_agents"]
        e = {"config": label, "arenas_per_gpu": arenas, "episode_age_steps": age, "us": r["ms_per_launch"] * 1e3,
             "env_steps_per_s": ctx.world * arenas / (r["ms_per_launch"] * 1e-3),
             "agent_steps_per_s": ctx.world * arenas * A / (r["ms_per_launch"] * 1e-3),
             "bytes_per_env_step": B, "achieved_gbs": gbs, "frac": frac, "kernel": "agar_step_kernel (" + r["env"].kernel_variant + ")",
             "clocks": sampler.region(*r["window"]) if ctx.rank == 0 else None}
        r["env"].close()
        del r
        torch.cuda.empty_cache()
        return e

    extras["cfg1_world"] = world_entry("cfg1", args.arenas, args.age, "BASELINE configs[0]/[2]/[4] world: " + describe(WORLDS["cfg1"]))
    extras["cfg2_mass_decay"] = world_entry("cfg2_decay", args.arenas, args.age, "headline world with the optional mass-decay rule: " + describe(WORLDS["cfg2_decay"]))
    # configs[3]: 262,144 arenas over 8 GPUs = 32,768 per GPU (25.8 GB of observations per step and GPU)
    free, _ = torch.cuda.mem_get_info(ctx.dev)
    n4 = CFG4_ARENAS_PER_GPU if free > 80e9 else 4096
    extras["cfg4_world"] = world_entry("cfg4", n4, min(args.age, 200), "BASELINE configs[3] world: " + describe(WORLDS["cfg4"]))

    # configs[4]: env-step throughput sweep over the arena count (cfg1 world; per GPU 1 Ki .. 512 Ki arenas, i.e. 1 K .. 4 M arenas
    # over 1 .. 8 GPUs), device-resident with the grid observation, each size aged by 200 untimed steps
    sweep = []
    for n_sw in (1024, 4096, 16384, 65536, 262144, 524288):
        free, _ = torch.cuda.mem_get_info(ctx.dev)
        if free < 2.5 * n_sw * 49152 + 2e9:  # two observation buffers + state
            continue
        r = time_world(ctx, "cfg1", n_sw, 20, W, 200, args.seed)
        B, gbs, frac = roofline_of(r["world"], n_sw, r["ms_per_launch"], peak)
        sweep.append({"arenas_per_gpu": n_sw, "arenas_total": ctx.world * n_sw, "us": r["ms_per_launch"] * 1e3,
                      "env_steps_per_s": ctx.world * n_sw / (r["ms_per_launch"] * 1e-3), "frac": frac})
        r["env"].close()
        del r
        torch.cuda.empty_cache()
    extras["sweep_cfg1_world"] = {"config": "BASELINE configs[4]: " + describe(WORLDS["cfg1"]) + ", episode age 200, 20 timed steps per size",
                                  "points": sweep}

    # rl.gae + rl.make_returns at configs[2]: T=128, N=65,536, gamma .75, lambda .1 (17 B per (t, n): SURVEY.md §8d).
    # Timed two ways.  (1) `us`: 60 launches back to back (a CUDA graph of three calls replayed 20 times) over three
    # rotating sets of inputs and outputs, 429 MB in all and 143 MB per call against 126 MB of L2, one event pair around
    # them: what a launch costs in a training loop, write-back of its own outputs included.  (2) `flushed_us`: one launch
    # per event pair after a 256 MB write that evicts L2 - this charges the launch with the write-back of that buffer's
    # dirty lines too (ncu, profiles/r02e_gae_final.txt: the kernel alone takes 24.4 us).
    from agarai_b200 import sharding
    T, N = 128, 65536
    dev = ctx.dev
    sets = []
    for _ in range(3):
        r = torch.randn(T, N, device=dev); v = torch.randn(T + 1, N, device=dev)
        d = (torch.rand(T, N, device=dev) < 0.01).to(torch.uint8)
        sets.append((r, v, d, v[T].clone()))
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MB > L2: evicts it between repetitions
    call = lambda q: ab.gae_and_returns(q[0], q[1], 0.75, 0.1, dones=q[2], end_value=q[3])
    for q in sets:
        call(q)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        outs = [call(q) for q in sets]   # three calls, three distinct output pairs kept alive
    graph.replay(); torch.cuda.synchronize(dev)
    t0 = time.time()
    rounds = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_b2b = sharding.max_over_ranks(e0.elapsed_time(e1) / (rounds * len(sets)), dev)
    ms, reps = 0.0, 200  # ~70 ms of wall clock: a few clock samples fall inside
    for i in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(sets[0]); e1.record()
        torch.cuda.synchronize(dev)
        ms += e0.elapsed_time(e1)
    ms = sharding.max_over_ranks(ms / reps, dev)
    nbytes = T * N * 17 + N * 8
    extras["gae"] = {"config": "BASELINE configs[2]: rl.gae + rl.make_returns over [128, 65536] per GPU, done cut, bootstrap",
                     "us": ms_b2b * 1e3, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (ms_b2b * 1e-3) / 1e9,
                     "frac": nbytes / (ms_b2b * 1e-3) / 1e9 / peak, "kernel": "agar_gae_tma_kernel<8,5,128>",
                     "l2": "60 launches back to back over three rotating input / output sets (429 MB; 143 MB per call, L2 is 126 MB)",
                     "flushed_us": ms * 1e3, "flushed_frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                     "flushed_l2": "one launch per event pair after a 256 MB write (its dirty lines are written back during the launch)",
                     "clocks": sampler.region(t0, time.time()) if ctx.rank == 0 else None}
    del sets, outs, graph

    # observe-only raster of the headline world (GridFeatureExtractor.extract for every agent: state read, obs written).
    # `us`: 60 launches back to back into three alternating 201 MB images (no flush needed: a launch streams 280 MB);
    # `flushed_us`: one launch per event pair after the 256 MB write.
    w = WORLDS["cfg2"]
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=args.arenas, **w), device=dev, seed=args.seed, action_config=ab.ActionConfig(8, 1, True, True))
    obs3 = [env.reset() for _ in range(3)]
    obs = obs3[0]
    for _ in range(3):
        env.observe(out_obs=obs)
    torch.cuda.synchronize(dev)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        env.observe(out_obs=obs3[i % 3])
    e1.record()
    torch.cuda.synchronize(dev)
    ms_b2b = sharding.max_over_ranks(e0.elapsed_time(e1) / 60, dev)
    ms = 0.0
    for _ in range(reps):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.observe(out_obs=obs); e1.record()
        torch.cuda.synchronize(dev)
        ms += e0.elapsed_time(e1)
    ms = sharding.max_over_ranks(ms / reps, dev)
    nbytes = args.arenas * (env.layout.words * 4 + obs[0].numel() * 4)
    extras["observe_only"] = {"config": f"grid observation of {args.arenas} cfg2 arenas from resident state (record read + image written)",
                              "us": ms_b2b * 1e3, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (ms_b2b * 1e-3) / 1e9,
                              "frac": nbytes / (ms_b2b * 1e-3) / 1e9 / peak, "kernel": "agar_step_kernel (F_OBS|F_READONLY)",
                              "l2": "60 launches back to back into three alternating images (a launch streams 280 MB; L2 is 126 MB)",
                              "flushed_us": ms * 1e3, "flushed_frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                              "clocks": sampler.region(t0, time.time()) if ctx.rank == 0 else None}
    env.close()
    return extras


def run_cuda(args):
    import numpy as np
    import agarai_b200 as ab
    from agarai_b200 import sharding

    ctx = _Ctx()
    torch = ctx.torch
    N, A = args.arenas, WORLD["num_agents"]
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()

    r = time_world(ctx, None, N, K, W, args.age, args.seed, no_obs=args.no_obs, want_young=args.age > 0)
    env, idx = r["env"], r["idx"]
    ms_launch, launches, young_ms = r["ms_per_launch"], r["launches"], r.get("young_ms_per_launch")
    value = ctx.world * N / (ms_launch * 1e-3)
    clocks = sampler.region(*r["window"]) if ctx.rank == 0 else None

    # ---- e2e: host buffers through agar_step_host (H2D actions, D2H obs/reward/done inside), same aged arenas ----
    xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
    Ke = max(2, min(K, args.e2e_steps))
    h_idx = idx[W:W + Ke].cpu().numpy()
    h_xy = torch.from_numpy(np.ascontiguousarray(xy_tab[h_idx])).pin_memory()   # [Ke,N,A,2]
    h_act = torch.from_numpy(np.ascontiguousarray(act_tab[h_idx])).pin_memory()  # [Ke,N,A]
    h_obs = torch.empty(env.obs_shape, dtype=torch.float32).pin_memory()
    h_rew = torch.empty((N, A), dtype=torch.float32).pin_memory()
    h_done = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    def e2e_run():
        for j in range(W):  # the same W untimed warm-up steps as the device-timed region (W >= 3): a training run lives in the steady state
            env.step_host(h_xy[j % Ke].numpy(), h_act[j % Ke].numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            env.step_host(h_xy[i].numpy(), h_act[i].numpy(), h_obs.numpy(), h_rew.numpy(), h_done.numpy())
        torch.cuda.synchronize(ctx.dev)
        dt = time.perf_counter() - t0
        return dt, ctx.world * N * Ke / sharding.max_over_ranks(dt, ctx.dev)

    # (1) every grid streamed out in full each step (any h_obs, the library's default); (2) the env-owned observation
    # buffer of a vector env: the same pinned buffer every step, declared with agar_host_obs_persistent, updated in place
    _, e2e_full = e2e_run()
    env.host_obs_persistent(True)
    e2e_local, e2e_value = e2e_run()
    h2d, d2h = env.host_copy_bytes()            # what the last call moved over PCIe (compact copy: occupied bins, not grids)
    delivered = h_obs.numel() * 4 + N * A * (4 + 1)   # what the caller received in its host buffers
    e2e_gbs_local = delivered * Ke / e2e_local / 1e9
    obs_mb = h_obs.numel() * 4 / 1e6
    state_mb = env.layout.words * 4 * N / 1e6
    env.close()
    del r, env, h_obs
    torch.cuda.empty_cache()

    peak, peak_src = measured_peak_gbs()
    extras = None if (args.no_extras or args.set or args.no_obs) else run_extras(ctx, args, sampler, peak)

    if ctx.rank == 0:
        B, achieved, frac = roofline_of(WORLD, N, ms_launch, peak)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": K, "warmup": W,
            "ms_per_step": ms_launch, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.world}: {N} arenas/GPU, {describe(WORLD)}, random policy over 25 actions (8 dirs, split, feed)",
                       "arenas_per_gpu": N, "episode_age_steps": args.age,
                       "l2": f"no flush: each step streams {state_mb:.0f} MB of state and writes a {obs_mb:.0f} MB observation "
                             "buffer (2 alternate); L2 is 126 MB",
                       **WORLD},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "host_bytes_delivered_per_step": delivered, "host_delivery_gbs_rank0": e2e_gbs_local,
                    "full_rewrite_value": e2e_full,
                    "api": "agar_step_host: actions from pinned host memory, dense float32 grids + rewards + dones in pinned host "
                           "buffers after every step (4 chunks; grids cross PCIe as per-agent occupied-bin lists and the library's host "
                           "threads bring the caller's grids up to date while the next chunk runs).  `value`: the env-owned observation "
                           "buffer of a vector env (same buffer every step, agar_host_obs_persistent: only the lines that changed are "
                           "rewritten, grids bit-identical); `full_rewrite_value`: the default, every grid streamed out in full"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": frac,
                         "traffic": ncu_traffic_bytes(args.world, N) if not args.set else None, "kernel": "agar_step_kernel",
                         "bytes_per_env_step": B, "env_steps_per_launch": N, "launch_us": ms_launch * 1e3, "peak_source": peak_src},
        }
        if young_ms is not None:  # the regime round 1 reported: the first steps after a reset
            _, yg, yf = roofline_of(WORLD, N, young_ms, peak)
            line["young_world"] = {"episode_age_steps": 0, "value": ctx.world * N / (young_ms * 1e-3), "unit": UNIT,
                                   "launch_us": young_ms * 1e3, "frac": yf}
        if extras is not None:
            line["extras"] = extras
        if ctx.world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = 512
            _, sec, _ = cpu_arm(sample, 1, 1, cores)        # size the sample for ~10 s of timed CPU work
            steps_cpu = int(max(3, min(100, 8.0 / max(sec, 1e-3))))
            v, sec, t_age = cpu_arm(sample, steps_cpu, 1, cores, age=args.age)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} arenas x {steps_cpu} steps of the same world after {args.age} untimed ageing "
                                              f"steps ({t_age:.0f} s), oracle/agar_oracle.c, one pthread per core"}
        print(json.dumps(line), flush=True)
    sampler.stop()
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


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
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg1 / cfg4 / GAE / observe-only extras")
    ap.add_argument("--age", type=int, default=DEFAULT_AGE,
                    help="untimed steps run before the warm-up in both arms (episode age of the timed region); 0 = right after reset")
    ap.add_argument("--no-obs", action="store_true", help="tuning only: step without the observation raster")
    ap.add_argument("--world", default="cfg2", choices=sorted(WORLDS), help="BASELINE.json world (default: the headline cfg2)")
    ap.add_argument("--set", action="append", default=[], metavar="FIELD=VALUE",
                    help="tuning experiments only: override a world field (e.g. ticks_per_step=8); the line says so in config")
    args = ap.parse_args()
    global WORLD
    WORLD = dict(WORLDS[args.world])
    for kv in args.set:
        k, v = kv.split("=")
        WORLD[k] = type(WORLDS[args.world][k])(float(v)) if k in WORLDS[args.world] else (float(v) if any(ch in v for ch in ".e") else int(v))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
