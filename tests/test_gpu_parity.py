"""GPU parity: the CUDA engine, called through the C-ABI, against the CPU oracle on the same
seeds and actions.  Bit-exact on everything: state words, observations, rewards, dones and the
event multiset (the CUDA build uses -fmad=false and the oracle -ffp-contract=off, so float
results are identical IEEE single operations; the 1e-5 relative bar of the north star is met
with zero difference)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
f32 = np.float32


def _fields(cfg):
    return {name: getattr(cfg, name) for name, _ in type(cfg)._fields_}


def make_pair(oracle, seed=0, **kw):
    import agarai_b200 as ab
    ocfg = oracle.default_config(**kw)
    gcfg = ab.default_config(**_fields(ocfg))
    genv = ab.BatchedAgarioEnv(gcfg, seed=seed)
    oenv = oracle.OracleEnv(ocfg, seed=seed, nthreads=8)
    L = genv.layout
    for name, _ in type(L)._fields_:
        assert getattr(L, name) == getattr(oenv.L, name), name
    return genv, oenv


def random_actions(rng, N, A, acts=(0, 1, 2), p_act=0.3):
    k = rng.integers(0, 9, size=(N, A))
    th = 2 * np.pi * ((k - 1) % 8) / 8
    mag = rng.choice([1.0, 0.5, 1.7], size=(N, A))  # 1.7 exercises the unit-disc clamp
    xy = np.where(k[..., None] == 0, 0, np.stack([np.cos(th), np.sin(th)], -1) * mag[..., None]).astype(f32)
    act = np.where(rng.random((N, A)) < p_act, rng.choice(acts, size=(N, A)), 0).astype(np.int32)
    return xy, act


def sorted_events(ev, cnt, cap):
    out = []
    for n in range(ev.shape[0]):
        assert cnt[n] <= cap, "event log overflow: raise event_capacity in the test"
        e = ev[n, :cnt[n]]
        out.append(e[np.lexsort(e.T[::-1])] if len(e) else e)
    return out


def assert_same_state(genv, oenv, where=""):
    gs = genv.get_state().cpu().numpy().view(np.int32)
    os_ = oenv.get_state().view(np.int32)
    if not np.array_equal(gs, os_):
        bad = np.argwhere(gs != os_)
        L = genv.layout
        names = sorted(((getattr(L, n), n) for n, _ in type(L)._fields_ if n not in ("words", "dyn_begin", "dyn_end", "num_players", "cells_total", "P_pad", "V_pad", "F_pad", "A_pad")))
        arena, word = bad[0]
        field = [n for o, n in names if o <= word][-1]
        raise AssertionError(f"{where}: state differs at {len(bad)} words; first arena {arena} word {word} ({field}+{word - getattr(L, field)}): gpu={gs[arena, word].view(f32) if False else gs[arena, word]} oracle={os_[arena, word]}")


def lockstep(genv, oenv, steps, rng, acts=(0, 1, 2), p_act=0.3, check_events=True):
    N, A = genv.num_envs, genv.num_agents
    cap = genv.cfg.event_capacity
    g_obs = genv.reset().cpu().numpy()
    o_obs = oenv.reset()
    assert np.array_equal(g_obs, o_obs), "reset observation"
    assert_same_state(genv, oenv, "after reset")
    counts = np.zeros(16, np.int64)
    for t in range(steps):
        xy, act = random_actions(rng, N, A, acts, p_act)
        g_o, g_r, g_d, _ = genv.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        o_o, o_r, o_d = oenv.step(xy, act)
        assert_same_state(genv, oenv, f"step {t}")
        assert np.array_equal(g_r.cpu().numpy().view(np.int32), o_r.view(np.int32)), f"rewards step {t}"
        assert np.array_equal(g_d.cpu().numpy(), o_d), f"dones step {t}"
        assert np.array_equal(g_o.cpu().numpy().view(np.int32), o_o.view(np.int32)), f"observations step {t}"
        if cap > 0 and check_events:
            ge, gc = genv.events(); oe, oc = oenv.events()
            ge, gc = ge.cpu().numpy(), gc.cpu().numpy()
            assert np.array_equal(gc, oc), f"event counts step {t}"
            for a_, b_ in zip(sorted_events(ge, gc, cap), sorted_events(oe, oc, cap)):
                assert np.array_equal(a_, b_), f"event trace step {t}"
                for e in a_:
                    counts[e[1]] += 1
    return counts


def test_cfg1_single_agent_pellets_only_1000_steps(oracle):
    """BASELINE configs[0]: single-agent pellets-only arena, 4 envs, random policy, 1000-step
    seeded rollout; full trace compare."""
    genv, oenv = make_pair(oracle, seed=1, num_arenas=4, event_capacity=64,
                           obs_flags=7)  # pellets, cells, others: 3x64x64
    counts = lockstep(genv, oenv, 1000, np.random.default_rng(0), acts=(0,), p_act=0.0)
    assert counts[1] > 20  # pellets were eaten
    genv.close()


def test_cfg2_bots_viruses_split_feed(oracle):
    """BASELINE configs[1] world (bots + viruses, arena 500, split/feed) at a test-sized N."""
    genv, oenv = make_pair(oracle, seed=2, num_arenas=6, num_bots=25, num_viruses=25, arena_size=500.0,
                           event_capacity=512, obs_flags=7)
    counts = lockstep(genv, oenv, 300, np.random.default_rng(1))
    assert counts[1] > 1000 and counts[4] > 0  # pellets and cell-eats happened
    genv.close()


def test_dense_small_arena_all_event_types(oracle):
    """Small crowded arena with low thresholds so that every rule fires often: split, feed, food
    eaten, virus pop, cell eaten, merge, death, bot respawn, auto-reset."""
    genv, oenv = make_pair(oracle, seed=3, num_arenas=8, num_agents=4, num_bots=6, num_viruses=6, num_pellets=400,
                           max_foods=16, arena_size=120.0, view_size=100.0, grid_size=32, obs_flags=31,
                           event_capacity=1024, auto_reset=1, start_mass=40.0, merge_delay=12.0,
                           virus_mass=30.0, virus_pop_mass=60.0, speed_k=6.0, pop_pieces=5)
    counts = lockstep(genv, oenv, 400, np.random.default_rng(2), p_act=0.5)
    for ev in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
        assert counts[ev] > 0, f"event type {ev} never fired: {counts}"
    genv.close()


def test_more_pellets_eaten_in_a_tick_than_the_hit_list_holds(oracle):
    """A huge cell over a carpet of pellets eats hundreds of them per tick: the kernel's per-tick hit list
    (128 entries) overflows and the resolver falls back to scanning every pellet's winner slot."""
    genv, oenv = make_pair(oracle, seed=11, num_arenas=4, num_agents=2, num_bots=1, num_pellets=1500, arena_size=60.0,
                           view_size=64.0, grid_size=32, obs_flags=7, event_capacity=8192, start_mass=2500.0,
                           split_min_mass=1e9, virus_pop_mass=1e9)
    counts = lockstep(genv, oenv, 12, np.random.default_rng(5), acts=(0,), p_act=0.0)
    assert counts[1] > 4 * 12 * 4 * 128, counts  # well over 128 pellets per arena-tick on average
    genv.close()


def test_cfg4_sixteen_agents_split_merge(oracle):
    """BASELINE configs[3] world: 16 agents per arena with split/merge (ppo.textproto:4-23)."""
    genv, oenv = make_pair(oracle, seed=4, num_arenas=3, num_agents=16, event_capacity=512, obs_flags=7,
                           start_mass=30.0, merge_delay=20.0)
    counts = lockstep(genv, oenv, 150, np.random.default_rng(3), acts=(0, 1), p_act=0.4)
    assert counts[6] > 0 and counts[5] > 0
    genv.close()


def test_no_regen_pellets_run_out(oracle):
    genv, oenv = make_pair(oracle, seed=5, num_arenas=2, num_pellets=60, arena_size=40.0, pellet_regen=0,
                           view_size=30.0, grid_size=16, event_capacity=128, speed_k=5.0)
    lockstep(genv, oenv, 200, np.random.default_rng(4), acts=(0,), p_act=0.0)
    L = genv.layout
    px = genv.get_state().cpu().numpy()[:, L.pellet_x:L.pellet_x + 60]
    assert (px < 0).sum() > 10
    genv.close()


def test_ragged_sizes_and_masked_reset(oracle):
    """P, V, F not multiples of 4, odd player count, reset of a subset of arenas."""
    genv, oenv = make_pair(oracle, seed=6, num_arenas=5, num_agents=3, num_bots=2, num_pellets=333, num_viruses=7,
                           max_foods=9, arena_size=90.0, view_size=64.0, grid_size=8, obs_flags=27, event_capacity=256,
                           start_mass=35.0, ticks_per_step=3)
    rng = np.random.default_rng(5)
    lockstep(genv, oenv, 40, rng)
    mask = np.array([1, 0, 0, 1, 0], np.uint8)
    g_obs = genv.reset(mask=torch.from_numpy(mask)).cpu().numpy()
    o_obs = oenv.reset(mask=mask)
    assert np.array_equal(g_obs, o_obs)
    assert_same_state(genv, oenv, "masked reset")
    xy, act = random_actions(rng, 5, 3)
    genv.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda()); oenv.step(xy, act)
    assert_same_state(genv, oenv, "step after masked reset")
    genv.close()


def test_sharding_invariance_and_state_roundtrip(oracle):
    """arena n of an 8-arena handle == arena 0 of a handle created with arena_id_base=n, and
    get_state -> set_state into a fresh handle continues identically (checkpoint/resume)."""
    import agarai_b200 as ab
    kw = dict(num_agents=2, num_bots=2, num_viruses=3, arena_size=150.0, obs_flags=15)
    big = ab.BatchedAgarioEnv(ab.default_config(num_arenas=8, **kw), seed=11)
    part = ab.BatchedAgarioEnv(ab.default_config(num_arenas=3, arena_id_base=5, **kw), seed=11)
    big.reset(); part.reset()
    rng = np.random.default_rng(6)
    for t in range(30):
        xy, act = random_actions(rng, 8, 2)
        txy, tact = torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda()
        ob, rb, db, _ = big.step(txy, tact)
        op, rp, dp, _ = part.step(txy[5:], tact[5:])
        assert torch.equal(ob[5:], op) and torch.equal(rb[5:], rp) and torch.equal(db[5:], dp)
    resumed = ab.BatchedAgarioEnv(ab.default_config(num_arenas=8, **kw), seed=11)
    resumed.set_state(big.get_state())
    xy, act = random_actions(rng, 8, 2)
    txy, tact = torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda()
    o1, r1, d1, _ = big.step(txy, tact); o2, r2, d2, _ = resumed.step(txy, tact)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(big.get_state(), resumed.get_state())
    for e in (big, part, resumed):
        e.close()


def test_step_discrete_equals_table_lookup(oracle):
    """agar_step_indexed == agar_step on the decoded table rows (ppo/action_conversion.py)."""
    import agarai_b200 as ab
    ac = ab.ActionConfig(8, 2, True, True)
    cfgkw = dict(num_arenas=4, num_agents=2, start_mass=40.0)
    a = ab.BatchedAgarioEnv(ab.default_config(**cfgkw), seed=1, action_config=ac)
    b = ab.BatchedAgarioEnv(ab.default_config(**cfgkw), seed=1, action_config=ac)
    a.reset(); b.reset()
    xy, act = ab.ppo_action_table(ac)
    assert len(act) == ab.action_size(ac) == 1 + 16 + 16
    rng = np.random.default_rng(7)
    for t in range(40):
        idx = rng.integers(0, len(act), size=(4, 2))
        oa, ra, da, _ = a.step_discrete(torch.from_numpy(idx).cuda())
        ob, rb, db, _ = b.step(torch.from_numpy(xy[idx]).cuda(), torch.from_numpy(act[idx]).cuda())
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    a.close(); b.close()


def test_raster_matches_reference_golden(oracle):
    """agar_observe_records on hand-built records vs the outputs recorded from the reference's
    GridFeatureExtractor (features/extractors.py:39-96): integer counts and sentinels bit-exact."""
    import agarai_b200 as ab
    from test_oracle_golden import check_obs_against_reference, load_grid_cases
    for c in load_grid_cases():
        K = int(c["K"]); P, V, F = len(c["pellets"]), len(c["viruses"]), len(c["foods"])
        cfg = ab.default_config(num_arenas=1, num_agents=1, num_bots=K - 1, num_pellets=P, num_viruses=V, max_foods=F,
                                grid_size=int(c["G"]), view_size=float(c["view"]), arena_size=float(c["arena"]),
                                obs_flags=int(c["flags"]))
        env = ab.BatchedAgarioEnv(cfg)
        L = env.layout
        rec = np.zeros((1, L.words), f32)
        rec[0, L.pellet_x:L.pellet_x + L.P_pad] = -1; rec[0, L.food_x:L.food_x + L.F_pad] = -1
        rec[0, L.pellet_x:L.pellet_x + P] = c["pellets"][:, 0]; rec[0, L.pellet_y:L.pellet_y + P] = c["pellets"][:, 1]
        rec[0, L.virus_x:L.virus_x + V] = c["viruses"][:, 0]; rec[0, L.virus_y:L.virus_y + V] = c["viruses"][:, 1]
        rec[0, L.food_x:L.food_x + F] = c["foods"][:, 0]; rec[0, L.food_y:L.food_y + F] = c["foods"][:, 1]
        rec[0, L.cell_x:L.cell_x + K * 16] = c["cell_x"]; rec[0, L.cell_y:L.cell_y + K * 16] = c["cell_y"]
        rec[0, L.cell_m:L.cell_m + K * 16] = c["cell_m"]
        got = env.observe_records(torch.from_numpy(rec).cuda()).cpu().numpy()[0, 0]
        check_obs_against_reference(got, c)
        ogot, _ = oracle.grid_extract(int(c["G"]), float(c["view"]), float(c["arena"]), int(c["flags"]), 0, K,
                                      c["pellets"], c["viruses"], c["foods"], c["cell_x"], c["cell_y"], c["cell_m"])
        assert np.array_equal(got.view(np.int32), ogot.view(np.int32))  # and bit-equal to the oracle, masses too
        env.close()


def test_gae_matches_oracle_and_reference_golden(oracle):
    import agarai_b200 as ab
    z = np.load(os.path.join(GOLD, "gae_golden.npz"))
    r, v = torch.from_numpy(z["r"]).cuda(), torch.from_numpy(z["v"]).cuda()
    for n in range(r.shape[1]):  # per-column hyper-parameters, as recorded from rl/test.py's numpy statements
        g, l = float(z["gamma"][n]), float(z["lam"][n])
        adv, ret = ab.gae_and_returns(r[:, n:n + 1].contiguous(), v[:, n:n + 1].contiguous(), g, l)
        assert np.array_equal(adv.cpu().numpy()[:, 0], z["adv"][:, n])
        assert np.array_equal(ret.cpu().numpy()[:, 0], z["ret0"][:, n])
    rng = np.random.default_rng(8)
    # N % 16 == 0 takes the TMA-fed kernel (producer warp, ring of five 8-row stages: T below, at and across the stage
    # and ring sizes, strips narrower than a CTA and a ragged last strip); other widths the register-ring kernel
    for T, N in ((1, 1), (3, 5), (128, 1000), (513, 257), (7, 4099), (1, 128), (8, 256), (9, 128), (24, 384), (25, 128),
                 (128, 65536), (131, 1280), (41, 48), (40, 176), (81, 2064), (200, 112), (39, 75776 + 16)):
        rr = rng.normal(size=(T, N)).astype(f32); vv = rng.normal(size=(T + 1, N)).astype(f32)
        dd = (rng.random((T, N)) < 0.05).astype(np.uint8); ev = rng.normal(size=N).astype(f32)
        for done, end in ((None, None), (dd, ev)):
            adv, ret = ab.gae_and_returns(torch.from_numpy(rr).cuda(), torch.from_numpy(vv).cuda(), 0.75, 0.1,
                                          dones=None if done is None else torch.from_numpy(done).cuda(),
                                          end_value=None if end is None else torch.from_numpy(end).cuda())
            oadv, oret = oracle.gae(rr, vv, 0.75, 0.1, done=done, end_value=end)
            assert np.array_equal(adv.cpu().numpy().view(np.int32), oadv.view(np.int32)), (T, N)
            assert np.array_equal(ret.cpu().numpy().view(np.int32), oret.view(np.int32)), (T, N)
        only_ret = ab.make_returns(torch.from_numpy(rr).cuda(), 0.75, end_value=torch.from_numpy(ev).cuda())  # no value buffer
        assert np.array_equal(only_ret.cpu().numpy().view(np.int32), oracle.gae(rr, vv, 0.75, 0.0, end_value=ev, want_adv=False)[1].view(np.int32)), (T, N)
    # rl/test.py:59-81 known answers through the CUDA path
    r3 = torch.tensor([[0.0], [3.0], [0.0]], device="cuda")
    for gamma in (1.0, 0.5, 0.0):
        a = ab.gae(r3, torch.zeros(4, 1, device="cuda"), gamma, 0.0)
        assert a[:, 0].tolist() == [0.0, 3.0, 0.0]
        a = ab.gae(r3, torch.tensor([[0.0], [0.0], [0.0], [7.0]], device="cuda"), gamma, 0.0)
        assert a[:, 0].tolist() == [0.0, 3.0, 7 * gamma]
    # test/returns_test.py:15-19: empty rollout keeps its length
    assert ab.make_returns(torch.zeros(0, 3, device="cuda"), 1.0).shape == (0, 3)


def test_gym_compat_view_matches_reference_call_sites(oracle):
    """The loop shape of ppo/train.py:156-197 on the compat view: list-of-(xy, act) actions in,
    (obs list, rewards, dones mask, info) out; single-agent (x, y, act) form of dqn/training.py."""
    import agarai_b200 as ab
    env = ab.make("agario-grid-v0", num_agents=3, multi_agent=True, difficulty="empty", ticks_per_step=4,
                  arena_size=200, num_pellets=200, num_viruses=0, num_bots=0, pellet_regen=True, grid_size=32,
                  num_frames=1, observe_cells=True, observe_others=True, observe_pellets=True,
                  observe_viruses=False, seed=5)
    assert env.observation_space.shape == (3, 32, 32)
    conv = ab.ActionConverter.ppo(ab.ActionConfig(8, 1))
    obs = env.reset()
    assert len(obs) == 3 and obs[0].shape == (3, 32, 32)
    dones = [False] * 3
    kept = []
    for i in range(20):
        if all(dones): break
        actions = [conv(1 + (i + a) % 8) for a in range(3)]
        next_obs, rewards, next_dones, _ = env.step(actions)
        kept.append(obs)
        assert rewards.shape == (3,) and next_dones.dtype == bool
        values = np.ones(3); values[next_dones] = 0.0  # boolean-mask use, ppo/train.py:192
        obs, dones = next_obs, next_dones
    assert not np.shares_memory(kept[0][0], kept[1][0])
    env.close()
    env1 = ab.make("agario-grid-v0", multi_agent=False, difficulty="empty", frames_per_step=4, arena_size=45,
                   num_pellets=30, num_viruses=0, num_bots=0, pellet_regen=True, grid_size=128, view_size=30.0,
                   observe_viruses=False)
    o = env1.reset()
    o, r, d, _ = env1.step((0.5, -0.5, 0))
    assert o.shape == (3, 128, 128) and isinstance(r, float) and isinstance(d, bool)
    env1.render(); env1.close()


def test_large_batch_properties(oracle):
    """Full-size property checks (BASELINE configs[1] N=4096): every reward equals the change of
    the agent's total mass, dones are consistent with empty cell rows, pellets stay in the arena,
    and a spot-check of 16 arenas against the oracle stepping from the same state."""
    import agarai_b200 as ab
    kw = dict(num_bots=25, num_viruses=25, arena_size=500.0, obs_flags=7, auto_reset=0)
    N = 4096
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=N, **kw), seed=9)
    env.reset()
    L = env.layout
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(12):
        idx = torch.randint(0, env.num_actions, (N, 1), generator=g, device="cuda", dtype=torch.int32)
        m0 = env.get_state()[:, L.cell_m:L.cell_m + 16].sum(1)
        s_before = env.get_state() if t == 11 else None
        obs, rew, done, _ = env.step_discrete(idx)
        st = env.get_state()
        m1 = st[:, L.cell_m:L.cell_m + 16].sum(1)
        assert torch.allclose(rew[:, 0], m1 - m0, rtol=1e-5, atol=1e-3)
        assert torch.equal(done[:, 0] != 0, m1 == 0)
        px = st[:, L.pellet_x:L.pellet_x + 1000]
        assert (px >= 0).all() and (px < 500.0).all()
        assert (obs[:, 0, 0] >= -1).all() and obs.shape == (N, 1, 3, 64, 64)
    # spot check: oracle continues 16 arenas from the GPU's state before the last step
    sel = np.arange(0, N, N // 16)[:16]
    xy_tab, act_tab = ab.ppo_action_table(env.action_config)
    for n in sel:
        ocfg = oracle.default_config(num_arenas=1, arena_id_base=int(n), **kw)
        oenv = oracle.OracleEnv(ocfg, seed=9)
        oenv.set_state(s_before[n:n + 1].cpu().numpy())
        i = int(idx[n, 0])
        o_o, o_r, o_d = oenv.step(xy_tab[i][None, None], act_tab[i][None, None])
        assert np.array_equal(oenv.get_state().view(np.int32), st[n:n + 1].cpu().numpy().view(np.int32))
        assert np.array_equal(o_o[0], obs[n].cpu().numpy()) and o_r[0, 0] == float(rew[n, 0])
    env.close()


def test_costly_arenas_first_order_does_not_change_results(monkeypatch):
    """Full-batch step launches of >= 2048 arenas map blocks to arenas through the cost-class lists the previous launch
    filed (three rotating order buffers).  The order is scheduling only: a twin env with the order switched off
    (AGAR_NO_ORDER, read at create) must stay bit-identical in state, observations, rewards and dones, including across a
    masked reset and a host-buffer step in between (which do not use the lists)."""
    import agarai_b200 as ab
    kw = dict(num_arenas=2304, num_agents=1, num_bots=6, num_viruses=4, num_pellets=200, arena_size=150.0, grid_size=16,
              obs_flags=7, auto_reset=1, start_mass=30.0, virus_pop_mass=60.0, virus_mass=30.0)
    a = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=5, action_config=ab.ActionConfig(8, 1, True, True))
    monkeypatch.setenv("AGAR_NO_ORDER", "1")
    b = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=5, action_config=ab.ActionConfig(8, 1, True, True))
    monkeypatch.delenv("AGAR_NO_ORDER")
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    N = kw["num_arenas"]
    for t in range(40):
        idx = torch.randint(0, a.num_actions, (N, 1), generator=g, device="cuda", dtype=torch.int32)
        oa, ra, da, _ = a.step_discrete(idx)
        ob, rb, db, _ = b.step_discrete(idx)
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)), t
        assert torch.equal(ra.view(torch.int32), rb.view(torch.int32)) and torch.equal(da, db), t
        if t == 15:  # a masked reset between ordered launches
            mask = (torch.arange(N, device="cuda") % 3 == 0).to(torch.uint8)
            assert torch.equal(a.reset(mask), b.reset(mask))
        if t == 25:  # a chunked host-buffer step (sub-batch launches: no lists) between ordered launches
            xy = np.zeros((N, 1, 2), f32); xy[:, 0, 0] = 0.7
            act = np.zeros((N, 1), np.int32)
            outs = []
            for e in (a, b):
                ho = np.empty(tuple(e.obs_shape), f32); hr = np.empty((N, 1), f32); hd = np.empty((N, 1), np.uint8)
                e.step_host(xy, act, ho, hr, hd)
                outs.append((ho, hr, hd))
            assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(*outs))
    assert torch.equal(a.get_state().view(torch.int32), b.get_state().view(torch.int32))
    a.close(); b.close()


def test_step_captured_in_a_cuda_graph_replays_like_plain_steps():
    """A step launch captured into a CUDA graph and replayed (launch-bound inner loops) gives the results of plain
    launches: the launch-order buffers rotate on the host side from launch to launch, so a captured launch must not use
    them (it would replay one fixed triple of buffers)."""
    import agarai_b200 as ab
    kw = dict(num_arenas=2304, num_agents=1, num_bots=5, num_viruses=3, num_pellets=200, arena_size=150.0, grid_size=16,
              obs_flags=7, auto_reset=1, start_mass=30.0)
    a = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=8, action_config=ab.ActionConfig(8, 1, True, True))
    b = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=8, action_config=ab.ActionConfig(8, 1, True, True))
    assert torch.equal(a.reset(), b.reset())
    N = kw["num_arenas"]
    gen = torch.Generator(device="cuda"); gen.manual_seed(4)
    idx_all = torch.randint(0, a.num_actions, (12, N, 1), generator=gen, device="cuda", dtype=torch.int32)
    for t in range(3):  # a few plain (ordered) launches first: an order is pending when the capture starts
        a.step_discrete(idx_all[t]); b.step_discrete(idx_all[t])
    idx = idx_all[3].clone()
    obs = torch.empty(a.obs_shape, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        _, rew, done, _ = a.step_discrete(idx, out_obs=obs)
    for t in range(3, 12):
        idx.copy_(idx_all[t])
        graph.replay()
        ob, rb, db, _ = b.step_discrete(idx_all[t])
        torch.cuda.synchronize()
        assert torch.equal(obs.view(torch.int32), ob.view(torch.int32)), t
        assert torch.equal(rew.view(torch.int32), rb.view(torch.int32)) and torch.equal(done, db), t
    assert torch.equal(a.get_state().view(torch.int32), b.get_state().view(torch.int32))
    a.close(); b.close()


def test_ppo_loop_runs_and_snapshots_rerasterise_exactly(oracle):
    """configs[2] shape at test size: rollout with GAE on the GPU, SGD on minibatches whose
    observations are re-rasterised from state snapshots (keep="state"); the re-rasterised observations equal the
    ones the step produced."""
    import agarai_b200 as ab
    from agarai_b200 import ppo
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=32, num_agents=1, obs_flags=7, auto_reset=1,
                                                arena_size=200.0, num_pellets=300, grid_size=32), seed=3,
                              action_config=ab.ActionConfig(8, 1))
    tr = ppo.PPOTrainer(env, ppo.PPOConfig(episode_length=16, num_sgd_steps=2, batch_size=64, num_features=64, keep="state"), seed=0)
    s0 = env.get_state()
    assert torch.equal(env.observe_records(s0), tr.obs)          # snapshot -> same observation as reset gave
    stats = tr.train_episode()
    assert np.isfinite(stats["loss"]) and tr.sgd_steps == 2
    assert tr.roll["rewards"].shape == (16, 32) and tr.roll["values"].shape == (17, 32)
    adv, ret = tr.advantages()
    oadv, oret = oracle.gae(tr.roll["rewards"].cpu().numpy(), tr.roll["values"].cpu().numpy(), 0.75, 0.1,
                            done=tr.roll["dones"].cpu().numpy(), end_value=tr.roll["values"][-1].cpu().numpy())
    assert np.array_equal(adv.cpu().numpy(), oadv) and np.array_equal(ret.cpu().numpy(), oret)
    stats2 = tr.train_episode()
    assert np.isfinite(stats2["loss"]) and tr.sgd_steps == 4
    env.close()


def test_ppo_presampled_minibatch_rows_hold_the_rollout_observations():
    """keep="presampled" (the default): the minibatch rows are drawn before the rollout and only their observations are
    kept; they must be exactly the observations the policy saw at those (t, b)."""
    import agarai_b200 as ab
    from agarai_b200 import ppo
    env = ab.BatchedAgarioEnv(ab.default_config(num_arenas=24, num_agents=2, obs_flags=7, auto_reset=1,
                                                arena_size=200.0, num_pellets=300, grid_size=16), seed=5,
                              action_config=ab.ActionConfig(8, 1))
    tr = ppo.PPOTrainer(env, ppo.PPOConfig(episode_length=12, num_sgd_steps=3, batch_size=32, num_features=32), seed=1)
    seen = []
    step = env.step_discrete
    def spy(actions, **kw):
        seen.append(tr._flat(tr.obs).clone())   # what the policy was evaluated on at this t
        return step(actions, **kw)
    env.step_discrete = spy
    tr.collect()
    hist = torch.stack(seen)                     # [T, B, C, G, G]
    rows = 0
    for k in range(3):
        t_idx, b_idx, obs = tr._minibatch(k)
        assert len(t_idx) == 32 and torch.equal(obs, hist[t_idx, b_idx])
        rows += len(t_idx)
    assert rows == 96 and "state" not in tr.roll and "observations" not in tr.roll
    adv, ret = tr.advantages()
    assert np.isfinite(tr.update(adv, ret)["loss"])
    env.close()


def test_ppo_reference_episode_regime_done_semantics(oracle):
    """auto_reset = 0: the reference's episode handling (ppo/train.py:164-197, 268-290): reset per episode, PRE-step
    dones recorded, bootstrap of finished agents zeroed, no cut inside GAE, finished agents' rows never sampled."""
    import agarai_b200 as ab
    from agarai_b200 import ppo
    kw = dict(num_arenas=16, num_agents=4, num_bots=6, num_viruses=6, num_pellets=400, max_foods=16, arena_size=120.0,
              view_size=100.0, grid_size=16, obs_flags=7, auto_reset=0, start_mass=40.0, merge_delay=12.0, virus_mass=30.0,
              virus_pop_mass=60.0, speed_k=6.0, pop_pieces=5)
    env = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=3, action_config=ab.ActionConfig(8, 1, True, True))
    for keep in ("presampled", "state"):
        tr = ppo.PPOTrainer(env, ppo.PPOConfig(episode_length=96, num_sgd_steps=2, batch_size=48, num_features=32, keep=keep), seed=0)
        tr.collect()
        d = tr.roll["dones"].cpu().numpy()                       # [T, B] pre-step
        assert not d[0].any() and d.any(), "no agent finished: the test world is too gentle"
        assert (np.diff(d.astype(np.int8), axis=0) >= 0).all()   # sticky
        final = np.zeros(tr.B, bool)
        dead_rows = d[-1] != 0
        boot = tr.roll["values"][-1].cpu().numpy()
        assert (boot[dead_rows] == 0).all()                      # ppo/train.py:189-195
        adv, ret = tr.advantages()
        oadv, oret = oracle.gae(tr.roll["rewards"].cpu().numpy(), tr.roll["values"].cpu().numpy(), 0.75, 0.1,
                                end_value=tr.roll["values"][-1].cpu().numpy())   # no done cut (rl/__init__.py:94-137)
        assert np.array_equal(adv.cpu().numpy(), oadv) and np.array_equal(ret.cpu().numpy(), oret)
        for k in range(2):
            t_idx, b_idx, obs = tr._minibatch(k)
            assert len(t_idx) == 48 and (tr.roll["dones"][t_idx, b_idx] == 0).all()   # ppo/train.py:284-290
        assert np.isfinite(tr.update(adv, ret)["loss"])
        ep0 = env.get_state()[:, env.layout.episode_steps].view(torch.int32)
        tr.collect()                                             # next episode starts from a reset
        assert not tr.roll["dones"][0].any()
    env.close()


def test_ppo_trainer_save_resume_reproduces_the_next_episode(tmp_path):
    """Trainer-level checkpoint (model, optimiser, schedule, env state + seed, sampling generator, torch RNG): a trainer
    resumed from the file produces the next episode's losses and parameters bit for bit."""
    import agarai_b200 as ab
    from agarai_b200 import ppo
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    kw = dict(num_arenas=16, num_agents=2, num_bots=2, obs_flags=7, auto_reset=1, arena_size=150.0, num_pellets=300, grid_size=16)
    cfg = ppo.PPOConfig(episode_length=10, num_sgd_steps=3, batch_size=32, num_features=32)

    def fresh(seed):
        env = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=11, action_config=ab.ActionConfig(8, 1, True, True))
        return ppo.PPOTrainer(env, cfg, seed=seed)

    a = fresh(0)
    a.train_episode()
    path = str(tmp_path / "ppo.ckpt")
    a.save(path)
    want = a.train_episode()
    want_params = [p.detach().clone() for p in a.model.parameters()]
    want_state = a.env.get_state().clone()
    b = fresh(123).resume(path)   # different construction seed: everything must come from the file
    got = b.train_episode()
    assert got == want, (got, want)
    assert all(torch.equal(x, y) for x, y in zip(want_params, b.model.parameters()))
    assert torch.equal(want_state.view(torch.int32), b.env.get_state().view(torch.int32))
    assert b.sgd_steps == a.sgd_steps == 6 and b.episodes == 2
    a.env.close(); b.env.close()


def test_raster_crowded_view_more_than_32_other_cells(oracle):
    """More than 32 other cells inside one agent's view (sequential patch path) and several cells
    per bin: bit-equal to the oracle, which is pinned to the reference extractor."""
    import agarai_b200 as ab
    rng = np.random.default_rng(11)
    K, P = 6, 200
    cfg = ab.default_config(num_arenas=3, num_agents=2, num_bots=K - 2, num_pellets=P, num_viruses=4, max_foods=8,
                            grid_size=32, view_size=96.0, arena_size=300.0, obs_flags=31)
    env = ab.BatchedAgarioEnv(cfg)
    L = env.layout
    rec = np.zeros((3, L.words), f32)
    for n in range(3):
        rec[n, L.pellet_x:L.pellet_x + L.P_pad] = -1; rec[n, L.food_x:L.food_x + L.F_pad] = -1
        rec[n, L.pellet_x:L.pellet_x + P] = rng.uniform(0, 300, P); rec[n, L.pellet_y:L.pellet_y + P] = rng.uniform(0, 300, P)
        rec[n, L.virus_x:L.virus_x + 4] = rng.uniform(100, 200, 4); rec[n, L.virus_y:L.virus_y + 4] = rng.uniform(100, 200, 4)
        rec[n, L.food_x:L.food_x + 5] = rng.uniform(120, 180, 5); rec[n, L.food_y:L.food_y + 5] = rng.uniform(120, 180, 5)
        centre = np.array([20.0, 150.0, 280.0])[n]  # corner / middle / far edge: sentinels in play
        rec[n, L.cell_x:L.cell_x + K * 16] = np.clip(centre + rng.uniform(-30, 30, K * 16), 0, 300)
        rec[n, L.cell_y:L.cell_y + K * 16] = np.clip(centre + rng.uniform(-30, 30, K * 16), 0, 300)
        m = rng.uniform(5, 300, K * 16); m[rng.random(K * 16) < 0.1] = 0
        rec[n, L.cell_m:L.cell_m + K * 16] = m
        # clustered cells: many in the same bin
        rec[n, L.cell_x + 40:L.cell_x + 60] = centre + 1.0; rec[n, L.cell_y + 40:L.cell_y + 60] = centre - 2.0
    got = env.observe_records(torch.from_numpy(rec).cuda()).cpu().numpy()
    ocfg = oracle.default_config(**_fields(cfg))
    exp = oracle.OracleEnv(ocfg).observe_records(rec)
    assert np.array_equal(got.view(np.int32), exp.view(np.int32))
    assert (exp[:, :, 2] > 0).sum(axis=(2, 3)).max() > 20  # the others channel really is crowded
    env.close()


def test_ram_features_match_oracle_and_reference_golden(oracle):
    """FeatureExtractor (features/extractors.py:99-149) on the device: (1) the reference's own outputs on the
    golden entity sets, loaded as engine state; (2) the oracle on states reached by stepping dense worlds,
    including agents without cells, fresh splits (equal keys) and fewer players than slots."""
    import agarai_b200 as ab
    z = np.load(os.path.join(GOLD, "ram_golden.npz"))
    for i in range(int(z["n"])):
        g = {k: z[f"{k}_{i}"] for k in ("counts", "K", "px", "py", "vx", "vy", "fx", "fy", "cx", "cy", "cix", "ciy", "cm", "expect")}
        K = int(g["K"])
        cfg = ab.default_config(num_arenas=1, num_agents=1, num_bots=K - 1, num_pellets=len(g["px"]),
                                num_viruses=len(g["vx"]), max_foods=len(g["fx"]))
        env = ab.BatchedAgarioEnv(cfg, device="cuda:0", seed=0)
        L = env.layout
        st = np.zeros((1, L.words), np.float32)
        st[0, L.pellet_x:L.pellet_x + L.P_pad] = -1.0
        st[0, L.food_x:L.food_x + L.F_pad] = -1.0
        for name, off in (("px", L.pellet_x), ("py", L.pellet_y), ("vx", L.virus_x), ("vy", L.virus_y), ("fx", L.food_x),
                          ("fy", L.food_y), ("cx", L.cell_x), ("cy", L.cell_y), ("cix", L.cell_ix), ("ciy", L.cell_iy),
                          ("cm", L.cell_m)):
            st[0, off:off + len(g[name])] = g[name]
        env.set_state(torch.from_numpy(st))
        ex = ab.FeatureExtractor(*[int(v) for v in g["counts"]])
        got = ex.extract(env).cpu().numpy()[0, 0]
        assert np.array_equal(got.astype(np.float64), g["expect"]), f"golden case {i}"
        env.close()
    # oracle on evolving states
    genv, oenv = make_pair(oracle, seed=21, num_arenas=6, num_agents=3, num_bots=4, num_viruses=5, num_pellets=300,
                           max_foods=16, arena_size=100.0, start_mass=40.0, merge_delay=12.0, auto_reset=0,
                           virus_mass=30.0, virus_pop_mass=60.0, speed_k=6.0)
    genv.reset(); oenv.reset()
    rng = np.random.default_rng(3)
    for counts in ((50, 5, 10, 5, 15), (7, 3, 2, 9, 16), (400, 8, 32, 2, 1)):
        ex = ab.FeatureExtractor(*counts)
        for t in range(25):
            xy, act = random_actions(rng, 6, 3, (0, 1, 2), 0.6)
            genv.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda(), want_obs=False)
            oenv.step(xy, act, want_obs=False)
            got = ex.extract(genv).cpu().numpy()
            want = oenv.observe_ram(*counts)
            assert np.array_equal(got.view(np.int32), want.view(np.int32)), f"{counts} step {t}"
    genv.close()


def test_screen_gray_matches_reference_golden(oracle):
    """ScreenFeatureExtractor.extract (features/extractors.py:152-164): float64 bit-exact against the
    reference's outputs; float32 = that value rounded once; ragged tail; full-size property check."""
    import agarai_b200 as ab
    z = np.load(os.path.join(GOLD, "screen_golden.npz"))
    ex = ab.ScreenFeatureExtractor(2, 40)
    for key, want in (("frames", "gray"), ("odd", "odd_gray")):
        x = torch.from_numpy(z[key]).cuda()
        assert np.array_equal(ex.extract(x).cpu().numpy(), z[want])
        assert np.array_equal(ex.extract(x, dtype=torch.float32).cpu().numpy(), z[want].astype(np.float32))
    big = torch.randint(0, 256, (64, 4, 128, 128, 3), dtype=torch.uint8, device="cuda")   # dqn: 4 frames of 128x128
    got = ex.extract(big)
    assert np.array_equal(got.cpu().numpy(), big.cpu().numpy().mean(axis=-1))  # numpy's own statement, full size
    assert np.array_equal(got[:2].cpu().numpy(), oracle.screen_gray(big[:2].cpu().numpy()))
    assert np.array_equal(ex.extract(big, dtype=torch.float32).cpu().numpy(), big.cpu().numpy().mean(axis=-1).astype(np.float32))
    ragged = big.reshape(-1, 3)[:3 * 4096 + 100].reshape(1, -1, 3).contiguous()   # three whole tiles and a partial one
    for dt, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        assert np.array_equal(ex.extract(ragged, dtype=dt).cpu().numpy(), ragged.cpu().numpy().mean(axis=-1).astype(npdt))
    assert ex.extract(big[:0]).shape == (0, 4, 128, 128)


def test_coordinator_matches_stand_alone_envs_and_reference_semantics(oracle):
    """a2c/coordinator.py:17-100 on the batched engine: env i of the Coordinator is bit-identical to a
    stand-alone gym-style env built with arena_id_base=i; finished envs return None and stay finished until
    reset() (coordinator.py:43-66)."""
    import agarai_b200 as ab
    kw = dict(num_agents=1, multi_agent=False, difficulty="normal", num_pellets=200, num_viruses=3, num_bots=10, arena_size=50.0,
              grid_size=32, observe_viruses=True, start_mass=12.0)
    W = 5
    singles = [ab.make("agario-grid-v0", seed=9, arena_id_base=i, **kw) for i in range(W)]
    rng = np.random.default_rng(17)
    with ab.Coordinator(lambda: kw, W, seed=9) as coord:
        assert coord.observation_space().shape == singles[0].observation_space.shape
        obs = coord.reset()
        for i, e in enumerate(singles):
            assert np.array_equal(obs[i], e.reset())
        alive = [True] * W
        for t in range(300):
            acts = [(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), int(rng.integers(0, 3))) for _ in range(W)]
            obs, rs, dones, infos = coord.step(acts)
            for i, e in enumerate(singles):
                if not alive[i]:
                    assert obs[i] is None and rs[i] is None and infos[i] is None and dones[i]
                    continue
                o, r, d, _ = e.step(acts[i])
                assert np.array_equal(obs[i], o) and rs[i] == r and dones[i] == d, (t, i)
                alive[i] = not d
        assert not all(alive), "no env finished in 300 steps: the sticky-done path was not exercised"
        obs = coord.reset()
        assert all(o is not None for o in obs) and coord.dones == [False] * W
    for e in singles:
        e.close()
    # multi-agent env behind RemoteEnvironment (a2c/remote_environment.py:20-53)
    mkw = dict(num_agents=3, difficulty="empty", num_pellets=100, arena_size=60.0, grid_size=16)
    with ab.RemoteEnvironment(mkw, seed=2) as renv:
        ref = ab.make("agario-grid-v0", seed=2, **mkw)
        o1, o2 = renv.reset(), ref.reset()
        assert all(np.array_equal(a, b) for a, b in zip(o1, o2))
        acts = [(np.array([0.5, -0.5], np.float32), 0), (np.array([0.0, 1.0], np.float32), 1), (np.array([-1.0, 0.0], np.float32), 2)]
        a1, a2 = renv.step(acts), ref.step(acts)
        assert all(np.array_equal(a, b) for a, b in zip(a1[0], a2[0])) and np.array_equal(a1[1], a2[1]) and np.array_equal(a1[2], a2[2])
        ref.close()


def test_ram_env_id_and_full_observation_view(oracle):
    """gym id agario-ram-v0 (configuration/__init__.py:38-42, dqn/train.py:89-96): observations are the
    FeatureExtractor vectors; the FullObservation host view lists exactly what the extractor consumed."""
    import agarai_b200 as ab
    kw = dict(num_agents=2, difficulty="normal", num_pellets=120, num_viruses=4, num_bots=3, arena_size=90.0)
    env = ab.make("agario-ram-v0", seed=4, **kw)          # the reference-shaped per-arena view
    obs = env.reset()
    assert len(obs) == 2 and obs[0].shape == (580,) == env.observation_space.shape
    o, r, d, _ = env.step([(np.array([1.0, 0.0], np.float32), 0), (np.array([0.0, -1.0], np.float32), 1)])
    want = ab.FeatureExtractor().extract(env.env).cpu().numpy()[0]
    assert np.array_equal(o[0], want[0]) and np.array_equal(o[1], want[1])
    full = ab.features.full_observation(env.env, arena=0, agent=1)
    L = env.env.layout
    st = env.env.get_state()[0].cpu().numpy()
    got, ok = oracle.ram_extract((50, 5, 10, 5, 15), 1, L.num_players,
                                 np.stack([st[L.pellet_x:L.pellet_x + 120], st[L.pellet_y:L.pellet_y + 120]], 1),
                                 np.stack([st[L.virus_x:L.virus_x + 4], st[L.virus_y:L.virus_y + 4]], 1),
                                 np.stack([st[L.food_x:L.food_x + 64], st[L.food_y:L.food_y + 64]], 1),
                                 *[st[o_:o_ + L.cells_total] for o_ in (L.cell_x, L.cell_y, L.cell_ix, L.cell_iy, L.cell_m)])
    assert ok and np.array_equal(got, o[1])
    assert full.pellets.shape == (120, 2) and full.viruses.shape == (4, 2) and full.agent.shape[1] == 5
    assert len(full.others) == 4 and all(p.shape[1] == 5 for p in full.others)
    env.close()
    benv = ab.make("agario-ram-v0", num_envs=16, seed=4, **kw)  # batched: torch tensors
    assert tuple(benv.reset().shape) == (16, 2, 580)
    benv.close()


@pytest.mark.parametrize("world", ["cfg1", "cfg2", "cfg4", "a2c", "a2c_test", "dqn"])
def test_compile_time_specialised_kernels_match_oracle_and_generic_kernel(oracle, world):
    """The BASELINE worlds run step kernels whose sizes and record offsets are compile-time constants
    (agar_kernels.cuh "config views").  Same source as the generic kernel: states, observations, rewards and
    dones must equal the oracle's, and the generic kernel's, bit for bit."""
    import agarai_b200 as ab
    import bench
    kw = dict(bench.WORLDS[world], num_arenas=24, auto_reset=1)
    genv, oenv = make_pair(oracle, seed=31, **kw)                   # event_capacity 0 -> specialised instantiation
    os.environ["AGAR_GENERIC_KERNEL"] = "1"
    try:
        renv = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=31)  # same world on the runtime-config kernel
    finally:
        del os.environ["AGAR_GENERIC_KERNEL"]
    assert genv.kernel_variant == world and renv.kernel_variant == "generic"
    rng = np.random.default_rng(8)
    N, A = 24, kw["num_agents"]
    g0 = genv.reset(); r0 = renv.reset()
    assert np.array_equal(g0.cpu().numpy(), oenv.reset()) and torch.equal(g0, r0)
    for t in range(80):
        xy, act = random_actions(rng, N, A, (0, 1, 2), 0.4)
        txy, tact = torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda()
        g_o, g_r, g_d, _ = genv.step(txy, tact)
        r_o, r_r, r_d, _ = renv.step(txy, tact)
        o_o, o_r, o_d = oenv.step(xy, act)
        assert_same_state(genv, oenv, f"{world} step {t}")
        assert np.array_equal(g_o.cpu().numpy().view(np.int32), o_o.view(np.int32)), f"obs step {t}"
        assert np.array_equal(g_r.cpu().numpy().view(np.int32), o_r.view(np.int32)) and np.array_equal(g_d.cpu().numpy(), o_d)
        assert torch.equal(g_o, r_o) and torch.equal(g_r, r_r) and torch.equal(g_d, r_d)
        assert torch.equal(genv.get_state(), renv.get_state())
    genv.close(); renv.close()


def test_step_host_chunked_path_equals_device_step(oracle):
    """agar_step_host steps batches of >= 1024 arenas in four chunks (observation copy of one chunk overlapping
    the kernel of the next): host results equal the device-buffer step of a twin env, ragged chunk sizes."""
    import agarai_b200 as ab
    kw = dict(num_arenas=1501, num_agents=2, num_bots=2, num_viruses=3, num_pellets=150, arena_size=90.0, grid_size=16,
              obs_flags=15, auto_reset=1)
    e1 = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=7)
    e2 = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=7)
    e1.reset(); e2.reset()
    rng = np.random.default_rng(1)
    obs = np.empty(e1.obs_shape, np.float32); rew = np.empty((1501, 2), np.float32); done = np.empty((1501, 2), np.uint8)
    for t in range(6):
        xy, act = random_actions(rng, 1501, 2, (0, 1, 2), 0.5)
        e1.step_host(xy, act, obs, rew, done)
        o, r, d, _ = e2.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        assert np.array_equal(obs, o.cpu().numpy()) and np.array_equal(rew, r.cpu().numpy()) and np.array_equal(done, d.cpu().numpy())
    assert torch.equal(e1.get_state(), e2.get_state())
    e1.close(); e2.close()


def test_random_configs_fuzz(oracle):
    """Differential fuzzing (tests/fuzz_parity.py): random sizes, grids, channel sets, thresholds, tick counts, arena
    id bases; state, observations, rewards, dones, event multisets and RAM features bit-equal to the oracle."""
    import fuzz_parity
    rng = np.random.default_rng(2024)
    for i in range(40):
        kw = fuzz_parity.random_config(rng)
        fuzz_parity.run_one(kw, int(rng.integers(0, 2 ** 31)), 25, rng)


def test_per_config_specialised_library_build(oracle):
    """build.build_specialised / make(..., specialise=True): a copy of the library whose step kernel has an
    arbitrary config's sizes as compile-time constants (nvcc at run time, cached in-tree) equals the oracle."""
    import shutil
    import agarai_b200 as ab
    if not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("nvcc is not on this machine")
    kw = dict(num_agents=2, num_bots=3, num_pellets=500, num_viruses=5, arena_size=200.0, grid_size=32,
              observe_viruses=True, difficulty="normal")
    env = ab.make("agario-grid-v0", num_envs=9, seed=12, specialise=True, gym_compat=False, **kw)
    assert env.kernel_variant == "extra"
    ocfg = oracle.default_config(**{n: getattr(env.cfg, n) for n, _ in type(env.cfg)._fields_})
    oenv = oracle.OracleEnv(ocfg, seed=12, nthreads=4)
    assert np.array_equal(env.reset().cpu().numpy(), oenv.reset())
    rng = np.random.default_rng(4)
    for t in range(40):
        xy, act = random_actions(rng, 9, 2, (0, 1, 2), 0.5)
        o, r, d, _ = env.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        oo, orr, od = oenv.step(xy, act)
        assert np.array_equal(o.cpu().numpy().view(np.int32), oo.view(np.int32)) and np.array_equal(r.cpu().numpy(), orr)
        assert np.array_equal(d.cpu().numpy(), od)
    assert np.array_equal(env.get_state().cpu().numpy().view(np.int32), oenv.get_state().view(np.int32))
    env.close()


def test_crowded_arena_eat_test(oracle):
    """More than 32 alive cells in an arena: the motion / pellet-query and cell-eats-cell phases (8 cells per warp pass,
    four warps) take a second pass.  Many split cells of several players in a small arena, eats every few ticks."""
    genv, oenv = make_pair(oracle, seed=17, num_arenas=6, num_agents=6, num_bots=6, num_viruses=2, num_pellets=300,
                           arena_size=260.0, grid_size=16, obs_flags=7, event_capacity=4096, auto_reset=1, start_mass=300.0,
                           split_min_mass=20.0, merge_delay=40.0, speed_k=5.0, virus_pop_mass=1e9)
    counts = lockstep(genv, oenv, 160, np.random.default_rng(6), acts=(0, 1, 1, 2), p_act=0.7)
    st = genv.get_state().cpu().numpy()
    L = genv.layout
    assert counts[4] > 20 and counts[6] > 100, counts          # cell-eaten and split events
    assert ((st[:, L.cell_m:L.cell_m + L.cells_total] > 0).sum(1) > 32).any(), "no arena stayed crowded"
    genv.close()


def test_mass_decay_rule_matches_oracle(oracle):
    """AgarConfig.mass_decay (own-spec rule, off by default): every alive cell loses that fraction of its mass at the
    start of a tick's motion unless the rest would fall below start_mass.  cfg2 world (the compile-time specialised
    kernel) and a dense world where every rule fires, both with a strong decay so that it shapes the trajectory."""
    import agarai_b200 as ab
    import bench
    n, seed = 6, 21
    kw = dict(bench.WORLDS["cfg2"], num_arenas=n, mass_decay=0.002)
    genv = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=seed, action_config=ab.ActionConfig(8, 1, True, True))
    assert genv.kernel_variant == "cfg2"
    oenv = oracle.OracleEnv(oracle.default_config(**kw), seed=seed, nthreads=8)
    xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
    assert np.array_equal(genv.reset().cpu().numpy(), oenv.reset())
    rng = np.random.default_rng(seed)
    L = genv.layout
    for t in range(400):
        idx = rng.integers(0, len(act_tab), (n, 1))
        go, gr, gd, _ = genv.step_discrete(torch.from_numpy(idx.astype(np.int32)).cuda())
        oo, orr, od = oenv.step(xy_tab[idx], act_tab[idx])
        assert np.array_equal(genv.get_state().cpu().numpy().view(np.int32), oenv.get_state().view(np.int32)), f"state, step {t}"
        assert np.array_equal(go.cpu().numpy().view(np.int32), oo.view(np.int32)), f"obs, step {t}"
        assert np.array_equal(gr.cpu().numpy().view(np.int32), orr.view(np.int32)) and np.array_equal(gd.cpu().numpy(), od), f"step {t}"
    m = genv.get_state().cpu().numpy()[:, L.cell_m:L.cell_m + L.cells_total]
    assert (m[m > 0] >= 10.0).all()  # never below start_mass
    genv.close(); oenv.close()
    genv, oenv = make_pair(oracle, seed=3, num_arenas=8, num_agents=4, num_bots=6, num_viruses=6, num_pellets=400,
                           max_foods=16, arena_size=120.0, view_size=100.0, grid_size=32, obs_flags=31,
                           event_capacity=1024, auto_reset=1, start_mass=40.0, merge_delay=12.0,
                           virus_mass=30.0, virus_pop_mass=60.0, speed_k=6.0, pop_pieces=5, mass_decay=0.01)
    counts = lockstep(genv, oenv, 300, np.random.default_rng(2), p_act=0.5)
    assert counts[4] > 0 and counts[6] > 0, counts
    genv.close()


def test_aged_cfg2_world_lockstep(oracle):
    """The regime a training run lives in (bench.py `episode_age_steps`): the BASELINE configs[1] world stepped for 1,500
    steps with the bench's random policy, state / observations / rewards / dones compared with the oracle at every step.
    The run must reach the states that dominate the steady-state cost: an arena with more than 40 alive cells (virus
    pops) and a cell heavier than 1,000."""
    import agarai_b200 as ab
    import bench
    n, seed = 8, 50
    kw = dict(bench.WORLDS["cfg2"], num_arenas=n)
    genv = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=seed, action_config=ab.ActionConfig(8, 1, True, True))
    assert genv.kernel_variant == "cfg2"
    oenv = oracle.OracleEnv(oracle.default_config(**kw), seed=seed, nthreads=8)
    xy_tab, act_tab = ab.ppo_action_table(ab.ActionConfig(8, 1, True, True))
    assert np.array_equal(genv.reset().cpu().numpy(), oenv.reset())
    rng = np.random.default_rng(seed)
    L = genv.layout
    max_alive, max_mass = 0, 0.0
    for t in range(1500):
        idx = rng.integers(0, len(act_tab), (n, 1))
        go, gr, gd, _ = genv.step_discrete(torch.from_numpy(idx.astype(np.int32)).cuda())
        oo, orr, od = oenv.step(xy_tab[idx], act_tab[idx])
        gs = genv.get_state().cpu().numpy()
        assert np.array_equal(gs.view(np.int32), oenv.get_state().view(np.int32)), f"state, step {t}"
        assert np.array_equal(go.cpu().numpy().view(np.int32), oo.view(np.int32)), f"obs, step {t}"
        assert np.array_equal(gr.cpu().numpy().view(np.int32), orr.view(np.int32)) and np.array_equal(gd.cpu().numpy(), od), f"step {t}"
        m = gs[:, L.cell_m:L.cell_m + L.cells_total]
        max_alive = max(max_alive, int((m > 0).sum(1).max()))
        max_mass = max(max_mass, float(m.max()))
    assert max_alive > 40 and max_mass > 1000.0, (max_alive, max_mass)
    genv.close(); oenv.close()
