#!/usr/bin/env python
"""Per-arena CTA cycles of one late step against the arena state (needs a -DAGAR_PHASE_CLOCKS build of the library):
    tools/build_variant.sh WORK clocks -DAGAR_PHASE_CLOCKS; AGAR_B200_LIB=$PWD/agarai_b200/_ab/clocks.so python tools/arena_cost.py"""
import ctypes as C, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import agarai_b200 as ab, bench
from agarai_b200 import _lib
W = bench.WORLDS["cfg2"]; n = 4096
env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=n, **W), device="cuda:0", seed=0, action_config=ab.ActionConfig(8, 1, True, True))
obs = env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
L = env.layout; lib = _lib.lib()
for t in range(1501):
    idx = torch.randint(0, env.num_actions, (n, 1), generator=g, device="cuda", dtype=torch.int32)
    if t == 1500:
        st = env.get_state().cpu().numpy()
    env.step_discrete(idx, out_obs=obs)
buf = (C.c_uint * 8192)(); lib.agar_debug_arena_cycles(buf)
cyc = np.array(buf[:n], dtype=np.float64)
m = st[:, L.cell_m:L.cell_m + L.cells_total]
alive = (m > 0).sum(1); mmax = m.max(1)
print("cycles per arena: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max()))
for lo, hi in ((0, 28), (28, 33), (33, 41), (41, 65), (65, 999)):
    sel = (alive >= lo) & (alive < hi)
    if sel.any(): print("alive %3d..%3d: %5d arenas, mean cycles %.0f" % (lo, hi - 1, sel.sum(), cyc[sel].mean()))
for lo, hi in ((0, 150), (150, 400), (400, 1000), (1000, 1e9)):
    sel = (mmax >= lo) & (mmax < hi)
    if sel.any(): print("max mass %5.0f..%5.0f: %5d arenas, mean cycles %.0f, mean alive %.1f" % (lo, min(hi, 99999), sel.sum(), cyc[sel].mean(), alive[sel].mean()))
print("corr(cycles, alive) %.2f  corr(cycles, maxmass) %.2f" % (np.corrcoef(cyc, alive)[0, 1], np.corrcoef(cyc, mmax)[0, 1]))
A = np.stack([np.ones(n), alive, alive.astype(float) ** 2, np.sqrt(mmax)], 1)
coef, *_ = np.linalg.lstsq(A, cyc, rcond=None)
print("fit cycles ~ %.0f + %.0f*n + %.1f*n^2 + %.0f*sqrt(maxmass)" % tuple(coef))
