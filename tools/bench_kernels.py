#!/usr/bin/env python
"""Per-kernel timings of the secondary kernels against the HBM roofline (CUDA events, inputs resident):
GAE/returns at BASELINE configs[2] size, screen->gray at the DQN frame size, RAM-style features and the
observation-only raster on the configs[1] world.  One JSON line per kernel.
    python tools/bench_kernels.py [--reps 50]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
import bench  # noqa: E402


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        flush.add_(1.0)  # 512 MB write: evicts L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--only", default="", help="gae: time the GAE kernel only")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, src = bench.measured_peak_gbs()
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)

    def report(name, ms, nbytes, extra):
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "us": ms * 1e3, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "peak_gbs": peak,
                          "frac": gbs / peak, "peak_source": src, "l2": "flushed by a 512 MB write between repetitions", **extra}), flush=True)

    # rl.gae + rl.make_returns, configs[2]: T=128, N=65,536 (17 B per (t, n): SURVEY.md §8d)
    T, N = 128, 65536
    r = torch.randn(T, N, device=dev); v = torch.randn(T + 1, N, device=dev)
    d = (torch.rand(T, N, device=dev) < 0.01).to(torch.uint8)
    ev = v[T].clone()
    ms = timed(lambda: ab.gae_and_returns(r, v, 0.75, 0.1, dones=d, end_value=ev), a.reps, flush)
    # the same launch back to back over three rotating input / output sets (a CUDA graph of three calls replayed 20 times):
    # no flush buffer whose dirty lines the launch would have to write back (bench.py's `extras.gae.us`)
    sets = [(r, v, d, ev)] + [(torch.randn_like(r), torch.randn_like(v), d.clone(), ev.clone()) for _ in range(2)]
    call = lambda q: ab.gae_and_returns(q[0], q[1], 0.75, 0.1, dones=q[2], end_value=q[3])
    for q in sets:
        call(q)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=dev)):
        outs = [call(q) for q in sets]
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 60
    report("agar_gae_tma_kernel", ms, T * N * 17 + N * 8, {"T": T, "N": N, "us_back_to_back": b2b * 1e3,
                                                           "frac_back_to_back": (T * N * 17 + N * 8) / (b2b * 1e-3) / 1e9 / peak})
    del sets, outs, graph
    if a.only == "gae":
        return

    # ScreenFeatureExtractor, dqn frame size: [1024, 4, 128, 128, 3] uint8
    x = torch.randint(0, 256, (1024, 4, 128, 128, 3), dtype=torch.uint8, device=dev)
    sx = ab.ScreenFeatureExtractor(4, 128)
    npx = x.numel() // 3
    out32 = torch.empty(x.shape[:-1], dtype=torch.float32, device=dev)
    import ctypes as C
    from agarai_b200 import _lib
    L = _lib.lib()
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ms = timed(lambda: L.agar_screen_gray(C.c_void_p(x.data_ptr()), npx, C.c_void_p(out32.data_ptr()), 0, st()), a.reps, flush)
    report("screen_gray_kernel<float>", ms, npx * 7, {"pixels": npx})
    out64 = torch.empty(x.shape[:-1], dtype=torch.float64, device=dev)
    ms = timed(lambda: L.agar_screen_gray(C.c_void_p(x.data_ptr()), npx, C.c_void_p(out64.data_ptr()), 1, st()), a.reps, flush)
    report("screen_gray_kernel<double>", ms, npx * 11, {"pixels": npx})
    del x, out32, out64

    # configs[1] world, 4096 arenas, a few steps in so that the state is a typical one
    W = bench.WORLDS["cfg2"]
    n = 4096
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=n, **W), device=dev, seed=0, action_config=ab.ActionConfig(8, 1, True, True))
    env.reset()
    idx = torch.randint(0, env.num_actions, (20, n, 1), device=dev, dtype=torch.int32)
    for i in range(20):
        env.step_discrete(idx[i], want_obs=False)
    fx = ab.FeatureExtractor()
    out = torch.empty((n, 1, fx.size), dtype=torch.float32, device=dev)
    ms = timed(lambda: fx.extract(env, out=out), a.reps, flush)
    rec_bytes = env.layout.words * 4
    report("ram_features_kernel", ms, n * (rec_bytes + fx.size * 4), {"arenas": n, "note": "sort-bound (bitonic sort of 1024 keys per agent), not HBM-bound"})
    obs = torch.empty(env.obs_shape, dtype=torch.float32, device=dev)
    ms = timed(lambda: env.observe(out_obs=obs), a.reps, flush)
    report("agar_step_kernel (observe only)", ms, n * (rec_bytes + obs[0].numel() * 4), {"arenas": n})
    env.close()


if __name__ == "__main__":
    main()
