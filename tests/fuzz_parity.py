#!/usr/bin/env python
"""Differential fuzzing of the CUDA engine against the CPU oracle (test infrastructure): random configs, random
actions, bit-exact comparison of state, observations, rewards, dones and (when logged) the event multiset.
    python tests/fuzz_parity.py --configs 200 --steps 40 [--seed 0]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
from oracle import oracle as O  # noqa: E402


def random_config(rng):
    A = int(rng.integers(1, 7)); B = int(rng.integers(0, 8))
    arena = float(rng.choice([40.0, 77.5, 120.0, 300.0, 1000.0]))
    kw = dict(
        num_arenas=int(rng.integers(1, 6)), num_agents=A, num_bots=B, num_pellets=int(rng.choice([0, 1, 17, 100, 333, 1000, 2500])),
        num_viruses=int(rng.choice([0, 0, 1, 5, 25])), max_foods=int(rng.choice([0, 1, 8, 64])), ticks_per_step=int(rng.choice([1, 2, 4, 7])),
        grid_size=int(rng.choice([4, 8, 12, 16, 32, 64, 128])), obs_flags=int(rng.integers(1, 32)), pellet_regen=int(rng.integers(0, 2)),
        auto_reset=int(rng.integers(0, 2)), event_capacity=int(rng.choice([0, 0, 4096])), arena_size=arena,
        view_size=float(rng.choice([16.0, 30.0, 64.0, 100.0, 128.0])), start_mass=float(rng.choice([10.0, 25.0, 60.0, 400.0])),
        merge_delay=float(rng.choice([0.0, 3.0, 12.0, 400.0])), virus_mass=float(rng.choice([30.0, 100.0])),
        virus_pop_mass=float(rng.choice([35.0, 60.0, 110.0])), speed_k=float(rng.choice([1.0, 3.0, 8.0])),
        pop_pieces=int(rng.choice([0, 3, 7, 15])), split_min_mass=float(rng.choice([12.0, 20.0])), bot_sense_radius=float(rng.choice([8.0, 64.0, 500.0])),
        arena_id_base=int(rng.choice([0, 7, 2 ** 33])), mass_decay=float(rng.choice([0.0, 0.0, 0.0005, 0.02])))
    if kw["max_foods"] == 0:
        kw["max_foods"] = 1
    return kw


def run_one(kw, seed, steps, rng):
    ocfg = O.default_config(**kw)
    gcfg = ab.default_config(**{n: getattr(ocfg, n) for n, _ in type(ocfg)._fields_})
    genv = ab.BatchedAgarioEnv(gcfg, seed=seed)
    oenv = O.OracleEnv(ocfg, seed=seed, nthreads=4)
    N, A, cap = kw["num_arenas"], kw["num_agents"], kw["event_capacity"]
    assert np.array_equal(genv.reset().cpu().numpy().view(np.int32), oenv.reset().view(np.int32)), "reset obs"
    for t in range(steps):
        xy = rng.uniform(-1.2, 1.2, (N, A, 2)).astype(np.float32)
        act = rng.integers(0, 3, (N, A)).astype(np.int32)
        g_o, g_r, g_d, _ = genv.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        o_o, o_r, o_d = oenv.step(xy, act)
        assert np.array_equal(genv.get_state().cpu().numpy().view(np.int32), oenv.get_state().view(np.int32)), f"state step {t}"
        assert np.array_equal(g_o.cpu().numpy().view(np.int32), o_o.view(np.int32)), f"obs step {t}"
        assert np.array_equal(g_r.cpu().numpy().view(np.int32), o_r.view(np.int32)), f"reward step {t}"
        assert np.array_equal(g_d.cpu().numpy(), o_d), f"done step {t}"
        if cap:
            ge, gc = genv.events(); oe, oc = oenv.events()
            ge, gc = ge.cpu().numpy(), gc.cpu().numpy()
            assert np.array_equal(gc, oc), f"event counts step {t}"
            for n in range(N):
                if oc[n] <= cap:
                    a_ = ge[n, :gc[n]]; b_ = oe[n, :oc[n]]
                    assert np.array_equal(a_[np.lexsort(a_.T[::-1])], b_[np.lexsort(b_.T[::-1])]), f"events step {t}"
    fx = ab.FeatureExtractor(int(rng.integers(1, 60)), int(rng.integers(0, 6)), int(rng.integers(0, 12)), int(rng.integers(0, 7)), int(rng.integers(1, 17)))
    want = oenv.observe_ram(fx.num_pellet, fx.num_virus, fx.num_food, fx.num_other, fx.num_cell)
    assert np.array_equal(fx.extract(genv).cpu().numpy().view(np.int32), want.view(np.int32)), "ram features"
    genv.close(); oenv.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, default=100)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    for i in range(a.configs):
        kw = random_config(rng)
        try:
            run_one(kw, int(rng.integers(0, 2 ** 31)), a.steps, rng)
        except AssertionError as e:
            print(f"MISMATCH in config {i}: {e}\n  {kw}")
            raise SystemExit(1)
    print(f"fuzz ok: {a.configs} random configs x {a.steps} steps bit-equal to the oracle")


if __name__ == "__main__":
    main()
