#!/usr/bin/env python
"""Debug aid for kernel variants: lockstep against the oracle on a scenario, and on the first differing state word
print the cell's fields on both sides (this step and the previous one) and the oracle's events for that arena.
    python tools/parity_debug.py [dense|fuzz:<seed>:<index>] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import agarai_b200 as ab  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_gpu_parity import random_actions  # noqa: E402

SCEN = {
    "dense": (dict(num_arenas=8, num_agents=4, num_bots=6, num_viruses=6, num_pellets=400, max_foods=16, arena_size=120.0,
                   view_size=100.0, grid_size=32, obs_flags=31, event_capacity=1024, auto_reset=1, start_mass=40.0,
                   merge_delay=12.0, virus_mass=30.0, virus_pop_mass=60.0, speed_k=6.0, pop_pieces=5), 3, 2, 0.5),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "dense"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    if which.startswith("fuzz:"):
        import fuzz_parity as FP
        _, seed, index = which.split(":")
        rng = np.random.default_rng(int(seed))
        for i in range(int(index) + 1):
            kw = FP.random_config(rng)
            sd = int(rng.integers(0, 2 ** 31))
            if i < int(index):
                # consume the same draws run_one makes
                N, A = kw["num_arenas"], kw["num_agents"]
                for t in range(30):
                    rng.uniform(-1.2, 1.2, (N, A, 2)); rng.integers(0, 3, (N, A))
                for _ in range(5):
                    rng.integers(1, 60)
        kw = dict(kw); kw["event_capacity"] = max(kw["event_capacity"], 4096)
        seed_env, arng, p_act = sd, rng, None
    else:
        kw, seed_env, aseed, p_act = SCEN[which]
        arng = np.random.default_rng(aseed)
    ocfg = O.default_config(**kw)
    gcfg = ab.default_config(**{n: getattr(ocfg, n) for n, _ in type(ocfg)._fields_})
    genv = ab.BatchedAgarioEnv(gcfg, seed=seed_env)
    oenv = O.OracleEnv(ocfg, seed=seed_env, nthreads=4)
    L = genv.layout
    N, A = kw["num_arenas"], kw["num_agents"]
    genv.reset(); oenv.reset()
    prev = oenv.get_state().copy()
    names = sorted(((getattr(L, n), n) for n, _ in type(L)._fields_ if n not in ("words", "dyn_begin", "dyn_end", "num_players", "cells_total", "P_pad", "V_pad", "F_pad", "A_pad")))
    for t in range(steps):
        if p_act is None:
            xy = arng.uniform(-1.2, 1.2, (N, A, 2)).astype(np.float32); act = arng.integers(0, 3, (N, A)).astype(np.int32)
        else:
            xy, act = random_actions(arng, N, A, (0, 1, 2), p_act)
        g_o, g_r, g_d, _ = genv.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        o_o, o_r, o_d = oenv.step(xy, act)
        gs = genv.get_state().cpu().numpy(); os_ = oenv.get_state()
        bad = np.argwhere(gs.view(np.int32) != os_.view(np.int32))
        ok_obs = np.array_equal(g_o.cpu().numpy().view(np.int32), o_o.view(np.int32))
        ok_r = np.array_equal(g_r.cpu().numpy().view(np.int32), o_r.view(np.int32)) and np.array_equal(g_d.cpu().numpy(), o_d)
        if len(bad) or not ok_obs or not ok_r:
            print(f"step {t}: {len(bad)} state words differ, obs ok {ok_obs}, reward/done ok {ok_r}")
            for arena, word in bad[:12]:
                field = [n for o, n in names if o <= word][-1]
                print(f"  arena {arena} word {word} ({field}+{word - getattr(L, field)}): gpu={gs[arena, word]!r} oracle={os_[arena, word]!r} prev={prev[arena, word]!r}")
            if len(bad):
                arena, word = bad[0]
                field = [n for o, n in names if o <= word][-1]
                if field.startswith("cell_"):
                    g = int(word - getattr(L, field))
                    k = g // 16
                    print(f"  player {k} (agent: {k < A}), action xy={xy[arena, k] if k < A else None} act={act[arena, k] if k < A else None}")
                    for sl in range(16):
                        gg = k * 16 + sl
                        f = lambda S: [float(S[arena, getattr(L, n) + gg]) for n in ("cell_x", "cell_y", "cell_ix", "cell_iy", "cell_m", "cell_t")]
                        if any(f(gs)) or any(f(os_)) or any(f(prev)):
                            print(f"    slot {sl}: gpu {f(gs)}\n             ora {f(os_)}\n             prv {f(prev)}")
                ev, cnt = oenv.events()
                print("  oracle events of the arena (tick, type, a, b):")
                for e in ev[arena, :cnt[arena]]:
                    if e[1] != 1:
                        print("    ", e.tolist())
            if not ok_obs:
                d = np.argwhere(g_o.cpu().numpy().view(np.int32) != o_o.view(np.int32))
                print("  obs diffs (first 8):", d[:8].tolist(), "of", len(d))
            return 1
        prev = os_.copy()
    print("ok", steps, "steps")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
