"""Long lockstep runs of the specialised step kernels against the CPU oracle (test infrastructure; run on the GPU box):
    python tests/long_parity.py"""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import agarai_b200 as ab, bench
from oracle import oracle as O
# long lockstep runs of the BASELINE worlds (specialised kernels) against the oracle: 600 steps, random policy over the PPO action set
for world, n in (("cfg2", 12), ("cfg4", 4), ("cfg1", 16), ("a2c_test", 6)):
    kw = dict(bench.WORLDS[world], num_arenas=n)
    ocfg = O.default_config(**kw)
    g = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=77, action_config=ab.ActionConfig(8, 1, True, True))
    o = O.OracleEnv(ocfg, seed=77, nthreads=8)
    xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
    assert np.array_equal(g.reset().cpu().numpy(), o.reset())
    rng = np.random.default_rng(5)
    A = kw["num_agents"]
    for t in range(600):
        idx = rng.integers(0, len(act_tab), (n, A))
        go, gr, gd, _ = g.step_discrete(torch.from_numpy(idx.astype(np.int32)).cuda())
        oo, orr, od = o.step(xy_tab[idx], act_tab[idx])
        assert np.array_equal(go.cpu().numpy().view(np.int32), oo.view(np.int32)), (world, t, "obs")
        assert np.array_equal(gr.cpu().numpy().view(np.int32), orr.view(np.int32)) and np.array_equal(gd.cpu().numpy(), od), (world, t)
        if t % 50 == 0:
            assert np.array_equal(g.get_state().cpu().numpy().view(np.int32), o.get_state().view(np.int32)), (world, t, "state")
    assert np.array_equal(g.get_state().cpu().numpy().view(np.int32), o.get_state().view(np.int32))
    print(world, g.kernel_variant, "600 steps ok; max mass", float(g.get_state()[:, g.layout.cell_m:g.layout.cell_m + g.layout.cells_total].max()))
    g.close(); o.close()
