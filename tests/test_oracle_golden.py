"""Pins the CPU oracle against the reference's own code (golden fixtures made by
tests/golden/make_golden.py from /root/reference) and known-answer vectors.  CPU only."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
f32 = np.float32


def load_grid_cases():
    z = np.load(os.path.join(GOLD, "grid_golden.npz"))
    keys = ["G", "view", "arena", "flags", "K", "pellets", "viruses", "foods", "cell_x", "cell_y", "cell_m", "expect", "loc"]
    return [{k: z[f"{k}_{i}"] for k in keys} for i in range(int(z["n"]))]


def count_mask(case):
    """channels holding integer counts (pellets, viruses, foods) vs masses (cells, others)"""
    out = []
    for bit in range(5):
        if int(case["flags"]) & (1 << bit):
            out.append(bit in (0, 3, 4))
    return out


def check_obs_against_reference(got, case):
    exp = case["expect"]
    assert got.shape == exp.shape
    for ch, is_count in enumerate(count_mask(case)):
        if is_count:  # bit-exact integer counts and sentinels
            assert np.array_equal(got[ch].astype(np.float64), exp[ch]), f"count channel {ch}"
        else:  # masses: same cells hit, same sentinels, float32 vs float64 accumulation
            assert np.array_equal(got[ch] == -1, exp[ch] == -1)
            assert np.array_equal(got[ch] != 0, exp[ch] != 0)
            np.testing.assert_allclose(got[ch], exp[ch], rtol=1e-6, atol=0)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_position_matches_reference_bit_exact(oracle):
    """features/extractors.py:207-212 on float32 cells: numpy's pairwise order reproduced."""
    z = np.load(os.path.join(GOLD, "position_golden.npz"))
    for x, y, m, loc in zip(z["x"], z["y"], z["m"], z["loc"]):
        got = oracle.centroid(x, y, m)
        assert got is not None and got[0] == loc[0] and got[1] == loc[1]
    assert oracle.centroid(np.zeros(16, f32), np.zeros(16, f32), np.zeros(16, f32)) is None


def test_grid_extract_matches_reference(oracle):
    """GridFeatureExtractor.extract (features/extractors.py:39-96), 42 recorded cases."""
    cases = load_grid_cases()
    assert len(cases) >= 40
    for c in cases:
        got, ok = oracle.grid_extract(int(c["G"]), float(c["view"]), float(c["arena"]), int(c["flags"]), 0,
                                      int(c["K"]), c["pellets"], c["viruses"], c["foods"],
                                      c["cell_x"], c["cell_y"], c["cell_m"])
        assert ok
        check_obs_against_reference(got, c)


def test_grid_no_other_players_leaves_channel_unmarked(oracle):
    """features/extractors.py:53-56: with no other player the `others` channel gets no -1 pass."""
    c = load_grid_cases()[1]  # corner case, K == 1
    assert int(c["K"]) == 1
    got, _ = oracle.grid_extract(int(c["G"]), float(c["view"]), float(c["arena"]), 7, 0, 1, c["pellets"],
                                 c["viruses"], c["foods"], c["cell_x"], c["cell_y"], c["cell_m"])
    assert (got[0] == -1).any() and (got[1] == -1).any() and not (got[2] != 0).any()


def test_grid_agent_without_cells_is_all_zero(oracle):
    got, ok = oracle.grid_extract(64, 128.0, 1000.0, 7, 0, 1, np.full((10, 2), 5, f32), np.zeros((0, 2), f32),
                                  np.zeros((0, 2), f32), np.zeros(16, f32), np.zeros(16, f32), np.zeros(16, f32))
    assert not ok and not got.any()


# ---- GAE / returns ---------------------------------------------------------------------------
def test_gae_known_answers(oracle):
    """rl/test.py:59-81: test_gae_lam0 and test_gae_values."""
    r = np.array([[0.0], [3.0], [0.0]], f32)
    for gamma in (1.0, 0.5, 0.0):
        adv, _ = oracle.gae(r, np.zeros((4, 1), f32), gamma, 0.0)
        np.testing.assert_allclose(adv[:, 0], [0.0, 3.0, 0.0])
        v = np.array([[0.0], [0.0], [0.0], [7.0]], f32)
        adv, _ = oracle.gae(r, v, gamma, 0.0)
        np.testing.assert_allclose(adv[:, 0], [0.0, 3.0, 7 * gamma])


def test_gae_matches_reference_numpy_statement(oracle):
    """rl/test.py:12-30 (`normal_gae`, `make_returns`) run on seeded float32 data, per column
    with its own (gamma, lam); end_value=0 and a2c.losses.make_returns are bit-exact, end_value=2
    differs only in how the first product is rounded (double in numpy, float32 in jax)."""
    z = np.load(os.path.join(GOLD, "gae_golden.npz"))
    r, v = z["r"], z["v"]
    for n in range(r.shape[1]):
        g, l = float(z["gamma"][n]), float(z["lam"][n])
        adv, ret = oracle.gae(r[:, n:n + 1], v[:, n:n + 1], g, l)
        assert np.array_equal(adv[:, 0], z["adv"][:, n])
        assert np.array_equal(ret[:, 0], z["ret0"][:, n])
        assert np.array_equal(ret[:, 0], z["reta2c"][:, n])
        _, ret2 = oracle.gae(r[:, n:n + 1], v[:, n:n + 1], g, l, end_value=np.array([2.0], f32), want_adv=False)
        np.testing.assert_allclose(ret2[:, 0], z["ret2"][:, n], rtol=1e-6, atol=1e-6)


def test_returns_properties(oracle):
    """test/returns_test.py:15-46."""
    for length in (1, 10):  # T == 0 is handled by the C-ABI wrapper (nothing to launch)
        _, ret = oracle.gae(np.zeros((length, 1), f32), np.zeros((length + 1, 1), f32), 1.0, 0.0, want_adv=False)
        assert ret.shape == (length, 1) and ret.sum() == 0
    rewards = (4 + np.arange(10)).astype(f32)[:, None]
    _, ret = oracle.gae(rewards, np.zeros((11, 1), f32), 0.0, 0.0, want_adv=False)
    assert np.array_equal(ret, rewards)
    z = np.load(os.path.join(GOLD, "gae_golden.npz"))
    _, ret = oracle.gae(z["rt"].astype(f32)[:, None], np.zeros((31, 1), f32), 0.75, 0.0, want_adv=False)
    np.testing.assert_allclose(ret[:, 0], z["rt_ret"], rtol=1e-5)  # float32 engine vs the float64 original


def test_gae_done_cuts_recursion(oracle):
    rng = np.random.default_rng(3)
    T, N = 12, 5
    r = rng.normal(size=(T, N)).astype(f32); v = rng.normal(size=(T + 1, N)).astype(f32)
    done = np.zeros((T, N), np.uint8); done[5] = 1
    adv, ret = oracle.gae(r, v, 0.9, 0.8, done=done)
    v_cut = v[:7].copy(); v_cut[6] = 0
    adv_a, ret_a = oracle.gae(r[:6], v_cut, 0.9, 0.8)
    adv_b, ret_b = oracle.gae(r[6:], v[6:], 0.9, 0.8)
    assert np.array_equal(adv, np.concatenate([adv_a, adv_b])) and np.array_equal(ret, np.concatenate([ret_a, ret_b]))


# ---- env rules: internal consistency of the (unpinned) spec -------------------------------------
def random_actions(rng, N, A, with_acts=True):
    k = rng.integers(0, 9, size=(N, A))
    th = 2 * np.pi * ((k - 1) % 8) / 8
    xy = np.where(k[..., None] == 0, 0, np.stack([np.cos(th), np.sin(th)], -1)).astype(f32)
    act = rng.integers(0, 3, size=(N, A)).astype(np.int32) if with_acts else np.zeros((N, A), np.int32)
    return xy, act


def test_env_reward_is_mass_delta_and_sharding_invariant(oracle):
    """reward = change of total mass (a2c/training.py:257); arena n of a 4-arena env equals arena 0
    of a 1-arena env created with arena_id_base = n (RNG keyed by global arena id)."""
    cfg = oracle.default_config(num_arenas=4, num_agents=2, num_bots=3, num_viruses=5, arena_size=200.0)
    env = oracle.OracleEnv(cfg, seed=7)
    env.reset()
    single = []
    for n in range(4):
        c1 = oracle.default_config(num_arenas=1, num_agents=2, num_bots=3, num_viruses=5, arena_size=200.0, arena_id_base=n)
        e1 = oracle.OracleEnv(c1, seed=7); e1.reset(); single.append(e1)
    rng = np.random.default_rng(0)
    L = env.L
    for t in range(60):
        xy, act = random_actions(rng, 4, 2)
        before = env.get_state()[:, L.cell_m:L.cell_m + 32].reshape(4, 2, 16).sum(-1)
        obs, rew, done = env.step(xy, act)
        after = env.get_state()[:, L.cell_m:L.cell_m + 32].reshape(4, 2, 16).sum(-1)
        np.testing.assert_allclose(rew, after - before, rtol=1e-5, atol=1e-4)
        for n in range(4):
            o1, r1, d1 = single[n].step(xy[n:n + 1], act[n:n + 1])
            assert np.array_equal(o1[0], obs[n]) and np.array_equal(r1[0], rew[n]) and np.array_equal(d1[0], done[n])


def test_env_multithreaded_step_is_identical(oracle):
    cfg = oracle.default_config(num_arenas=24, num_agents=1, num_bots=4, num_viruses=3, arena_size=150.0)
    a = oracle.OracleEnv(cfg, seed=3, nthreads=1); b = oracle.OracleEnv(cfg, seed=3, nthreads=4)
    a.reset(); b.reset()
    rng = np.random.default_rng(1)
    for t in range(20):
        xy, act = random_actions(rng, 24, 1)
        oa = a.step(xy, act); ob = b.step(xy, act)
        assert all(np.array_equal(x, y) for x, y in zip(oa, ob))
    assert np.array_equal(a.get_state().view(np.int32), b.get_state().view(np.int32))


def test_env_mass_decay_rule(oracle):
    """AgarConfig.mass_decay (own-spec rule; 0 = off is the default and leaves every trace unchanged): a lone cell that
    eats nothing keeps (1 - d) of its mass per tick in float32, and stops once the next product would fall below
    start_mass."""
    assert oracle.default_config().mass_decay == 0.0
    f32 = np.float32
    cfg = oracle.default_config(num_arenas=1, num_pellets=0, ticks_per_step=4, mass_decay=0.01, auto_reset=0)
    env = oracle.OracleEnv(cfg, seed=1); env.reset()
    L = env.L
    st = env.get_state(); st[0, L.cell_m] = 50.0; env.set_state(st)
    m, keep = f32(50.0), f32(1.0) - f32(0.01)
    for t in range(60):
        _, rew, _ = env.step(np.zeros((1, 1, 2), f32), np.zeros((1, 1), np.int32))
        m0 = m
        for _ in range(4):
            md = f32(m * keep)
            if md >= f32(10.0):
                m = md
        got = env.get_state()[0, L.cell_m]
        assert got.view(np.int32) == m.view(np.int32), (t, got, m)
        assert rew[0, 0] == f32(m - m0)
    assert f32(10.0) <= m < f32(10.2)
    with pytest.raises(ValueError):
        oracle.OracleEnv(oracle.default_config(mass_decay=0.7))


# ---------------------------------------------------------------- RAM-style / screen extractors
def load_ram_cases():
    z = np.load(os.path.join(GOLD, "ram_golden.npz"))
    keys = ("counts", "K", "px", "py", "vx", "vy", "fx", "fy", "cx", "cy", "cix", "ciy", "cm", "expect")
    return [{k: z[f"{k}_{i}"] for k in keys} for i in range(int(z["n"]))]


def test_ram_extract_matches_reference(oracle):
    """FeatureExtractor.extract (features/extractors.py:99-149) run from /root/reference on 32 entity sets:
    the C restatement reproduces every element bit for bit (float32 values in the reference's float64 vector)."""
    cases = load_ram_cases()
    assert len(cases) == 32
    for c in cases:
        counts = tuple(int(v) for v in c["counts"])
        got, ok = oracle.ram_extract(counts, 0, int(c["K"]), np.stack([c["px"], c["py"]], 1), np.stack([c["vx"], c["vy"]], 1),
                                     np.stack([c["fx"], c["fy"]], 1), c["cx"], c["cy"], c["cix"], c["ciy"], c["cm"])
        assert ok and got.shape == c["expect"].shape
        assert np.array_equal(got.astype(np.float64), c["expect"]), counts


def test_ram_extract_agent_without_cells_and_stable_ties(oracle):
    K = 2
    z = np.zeros(K * 16, np.float32)
    out, ok = oracle.ram_extract((3, 1, 1, 1, 2), 0, K, np.zeros((4, 2), np.float32), np.zeros((0, 2), np.float32),
                                 np.zeros((0, 2), np.float32), z, z, z, z, z)
    assert not ok and not out.any()
    # two pellets at the same distance: the lower index comes first (own spec: numpy leaves it open)
    cm = z.copy(); cm[0] = 10.0
    cx = z.copy(); cy = z.copy(); cx[0] = 50.0; cy[0] = 50.0
    pel = np.array([[53, 54], [47, 46], [50, 51]], np.float32)
    out, ok = oracle.ram_extract((3, 0, 0, 0, 1), 0, K, pel, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32),
                                 cx, cy, z, z, cm)
    assert ok and out[:6].tolist() == [0, 1, 3, 4, -3, -4]


def test_screen_gray_matches_reference(oracle):
    z = np.load(os.path.join(GOLD, "screen_golden.npz"))
    assert np.array_equal(oracle.screen_gray(z["frames"]), z["gray"])
    assert np.array_equal(oracle.screen_gray(z["odd"]), z["odd_gray"])
