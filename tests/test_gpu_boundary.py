"""GPU tests of the reference-facing boundary added in round 2: `agario-full-v0` + `extractor(observation)` (the DQN
trainer's call form, dqn/training.py:96-125,252-256), `make()` with exactly the reference's kwargs, `num_frames`."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32


def _to_action(index, action_shape):
    """dqn/training.py:220-228 restated"""
    indices = np.unravel_index(index, action_shape)
    theta = 2 * np.pi * indices[0] / action_shape[0]
    mag = 1 - indices[1] / action_shape[1]
    return np.cos(theta) * mag, np.sin(theta) * mag, int(indices[2])


def _oracle_lists(oracle, oenv, agent=0):
    """entity lists of arena 0 from the oracle's state, in the FullObservation conventions"""
    L, c = oenv.L, oenv.cfg
    rec = oenv.get_state()[0]
    px, py = rec[L.pellet_x:L.pellet_x + c.num_pellets], rec[L.pellet_y:L.pellet_y + c.num_pellets]
    fx, fy = rec[L.food_x:L.food_x + c.max_foods], rec[L.food_y:L.food_y + c.max_foods]
    vx, vy = rec[L.virus_x:L.virus_x + c.num_viruses], rec[L.virus_y:L.virus_y + c.num_viruses]
    KC = L.cells_total
    cells = {n: rec[getattr(L, "cell_" + n):getattr(L, "cell_" + n) + KC] for n in ("x", "y", "ix", "iy", "m")}
    return (np.stack([px, py], 1)[px >= 0], np.stack([vx, vy], 1), np.stack([fx, fy], 1)[fx >= 0], cells)


def _dqn_loop(oracle, env_kw, okw, extractor_kw, steps, seed, ram_counts=None):
    import agarai_b200 as ab
    # the reference's default DQN setup: dqn/hyperparameters.py:14-49 (action_shape (4, 1, 1)) and :85-107
    action_shape = (4, 1, 1)
    env = ab.make("agario-full-v0", seed=seed, **env_kw)                       # dqn/train.py:55-81
    extractor = ab.GridFeatureExtractor(**extractor_kw)                        # dqn/train.py:99-108
    ram = ab.FeatureExtractor(*ram_counts) if ram_counts else None
    oenv = oracle.OracleEnv(oracle.default_config(**okw), seed=seed)
    oenv.reset(want_obs=False)
    K = okw.get("num_bots", 0) + 1
    rng = np.random.default_rng(seed)
    flags = extractor.flags

    def check(observation):
        pel, vir, foo, cells = _oracle_lists(oracle, oenv)
        assert np.array_equal(observation.pellets, pel) and np.array_equal(observation.foods, foo)
        assert np.array_equal(observation.viruses, vir)
        own = cells["m"][:16] > 0
        assert np.array_equal(observation.agent[:, 4], cells["m"][:16][own]) and np.array_equal(observation.agent[:, 0], cells["x"][:16][own])
        assert len(observation.others) == sum(bool((cells["m"][16 * k:16 * k + 16] > 0).any()) for k in range(1, K))
        fts = extractor(observation)                                           # dqn/training.py:252-256
        want, ok = oracle.grid_extract(extractor.grid_size, extractor.view_size, extractor.arena_size, flags, 0, K, pel, vir, foo,
                                       cells["x"], cells["y"], cells["m"])
        if not ok:
            assert fts is None
            return
        assert fts.dtype == np.float64 and fts.shape == extractor.shape
        assert np.array_equal(fts.astype(f32).view(np.int32), want.view(np.int32)), "grid features differ from the oracle"
        if ram is not None:
            vec = ram(observation)
            wantv, _ = oracle.ram_extract(ram_counts, 0, K, pel, vir, foo, cells["x"], cells["y"], cells["ix"], cells["iy"], cells["m"])
            assert vec.shape == ram.shape and np.array_equal(vec.astype(f32).view(np.int32), wantv.view(np.int32)), "ram features differ"

    first = env.reset()                                                        # dqn/training.py:96
    check(first)
    total_returns = 0.0
    for i in range(steps):                                                     # dqn/training.py:99-123
        index = int(rng.integers(0, int(np.prod(action_shape))))
        action = _to_action(index, action_shape)
        next_state, reward, done, info = env.step(action)
        env.render()
        o_o, o_r, o_d = oenv.step(np.array([[action[:2]]], f32), np.array([[action[2]]], np.int32), want_obs=False)
        assert isinstance(reward, float) and isinstance(done, bool)
        assert f32(reward).view(np.int32) == o_r[0, 0].view(np.int32) and done == bool(o_d[0, 0])
        check(next_state)
        total_returns += reward
        if done:
            break
    env.close(); oenv.close()
    return total_returns


def test_dqn_shaped_loop_on_full_env_with_reference_defaults(oracle):
    """dqn/training.py:96-125 on `agario-full-v0` with the reference's default hyper-parameters
    (dqn/hyperparameters.py:23-28,89-104: arena 45, 30 pellets, frames_per_step 4; grid extractor 128 cells over a
    view of 30, cells / others / viruses channels): every step's FullObservation and `extractor(observation)` against
    the oracle."""
    env_kw = dict(frames_per_step=4, arena_size=45, num_pellets=30, num_viruses=0, num_bots=0, pellet_regen=True)
    okw = dict(ticks_per_step=4, arena_size=45.0, num_pellets=30, num_viruses=0, num_bots=0, pellet_regen=1)
    ext = dict(view_size=30, grid_size=128, arena_size=45, cells=True, others=True, viruses=True, food=False, flat=False)
    assert _dqn_loop(oracle, env_kw, okw, ext, 120, seed=42) > 0.0  # pellets were eaten


def test_dqn_shaped_loop_with_bots_viruses_and_both_extractors(oracle):
    """The same loop in a world with other players, viruses and ejected food, through both of the reference's
    extractors (`GridFeatureExtractor` and the RAM-style `FeatureExtractor`, dqn/train.py:89-108)."""
    env_kw = dict(frames_per_step=4, arena_size=160, num_pellets=300, num_viruses=4, num_bots=5, pellet_regen=True, start_mass=40.0)
    okw = dict(ticks_per_step=4, arena_size=160.0, num_pellets=300, num_viruses=4, num_bots=5, pellet_regen=1, start_mass=40.0)
    ext = dict(view_size=64, grid_size=32, arena_size=160, cells=True, others=True, viruses=True, food=True)
    _dqn_loop(oracle, env_kw, okw, ext, 80, seed=7, ram_counts=(20, 3, 4, 3, 5))


def test_make_with_exactly_the_reference_dqn_kwargs():
    """dqn/train.py:57-81 passes no `multi_agent` key and then steps with (x, y, act) (dqn/training.py:104)."""
    import agarai_b200 as ab
    env_config = {'frames_per_step': 4, 'arena_size': 45, 'num_pellets': 30, 'num_viruses': 0, 'num_bots': 0, 'pellet_regen': True}
    env_config.update({"grid_size": 128, "observe_cells": False, "observe_others": False, "observe_viruses": False,
                       "observe_pellets": True})   # dqn/hyperparameters.py:126-139
    env = ab.make("agario-grid-v0", **env_config)
    first = env.reset()
    assert first.shape == (1, 128, 128)
    x, y, act = _to_action(2, (4, 1, 1))
    state, reward, done, info = env.step((x, y, act))
    assert state.shape == (1, 128, 128) and isinstance(reward, float) and isinstance(done, bool) and isinstance(info, dict)
    env.close()
    # multi-agent call sites pass multi_agent=True (configuration/__init__.py:54) or num_agents > 1
    env = ab.make("agario-grid-v0", num_agents=2, grid_size=16, num_pellets=50, arena_size=100)
    obs, rewards, dones, _ = env.step([(np.array([1.0, 0.0]), 0), (np.array([0.0, 1.0]), 1)])
    assert len(obs) == 2 and rewards.shape == (2,)
    env.close()


def test_num_frames_stacks_the_last_k_grids(oracle):
    """num_frames = k (environment.proto:44, configuration/__init__.py:67): [N, A, k*C, G, G], oldest first, k copies of
    the first grid after a reset; compared with a ring of the oracle's single-frame observations."""
    import agarai_b200 as ab
    k, N, A = 3, 4, 2
    kw = dict(num_agents=A, num_bots=2, num_pellets=200, arena_size=150.0, grid_size=16, view_size=64.0, obs_flags=7)
    env = ab.make("agario-grid-v0", num_envs=N, seed=3, gym_compat=False, num_frames=k, multi_agent=True, **kw)
    oenv = oracle.OracleEnv(oracle.default_config(num_arenas=N, **kw), seed=3)
    C = 3
    assert env.obs_shape == (N, A, k * C, 16, 16) and env.observation_space.shape == (k * C, 16, 16)
    first = oenv.reset()
    ring = [first.copy() for _ in range(k)]
    assert np.array_equal(env.reset().cpu().numpy(), np.concatenate(ring, axis=2))
    rng = np.random.default_rng(0)
    kept = []
    for t in range(12):
        xy = rng.uniform(-1, 1, (N, A, 2)).astype(f32)
        act = rng.integers(0, 3, (N, A)).astype(np.int32)
        o, r, d, _ = env.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        oo, orr, od = oenv.step(xy, act)
        ring = ring[1:] + [oo.copy()]
        assert np.array_equal(o.cpu().numpy().view(np.int32), np.concatenate(ring, axis=2).view(np.int32)), f"step {t}"
        assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(d.cpu().numpy(), od)
        kept.append(o)
        if t == 5:  # masked reset: the reset arenas restart with k copies of their new first grid
            mask = np.array([1, 0, 1, 0], np.uint8)
            go = env.reset(mask=torch.from_numpy(mask)).cpu().numpy()
            fresh = oenv.reset(mask=mask)
            for j in range(k):
                ring[j] = np.where(mask.astype(bool)[:, None, None, None, None], fresh, ring[j])
            assert np.array_equal(go, np.concatenate(ring, axis=2))
    assert kept[0].data_ptr() != kept[1].data_ptr()  # handed-out observations are fresh tensors
    env.close(); oenv.close()
    # the per-arena view and the vector layer take the kwarg as well
    one = ab.make("agario-grid-v0", num_frames=2, grid_size=16, num_pellets=50, arena_size=100, multi_agent=False)
    assert one.reset().shape == (8, 16, 16)
    s, rew, done, _ = one.step((1.0, 0.0, 0))
    assert s.shape == (8, 16, 16)
    one.close()


@pytest.mark.parametrize("world", ["borders_multi_agent", "single_player", "dense_overflow", "cfg2"])
def test_step_host_persistent_buffer_updates_in_place(world):
    """agar_host_obs_persistent: with the caller's promise that h_obs is the same untouched memory from call to call, the
    host side rewrites only what changed since the previous step (bins that were / become occupied, out-of-bounds rows
    and columns that moved).  The buffer must equal the device observation bit for bit after every step: agents
    crossing the arena border (sentinel rows and columns shift), agents dying and arenas resetting, images fuller than
    the list capacity (copied whole, then rebuilt in full), a call with another buffer in between, a masked reset and a
    switch off / on of the mode."""
    import agarai_b200 as ab
    import bench
    kw = {
        "borders_multi_agent": dict(num_arenas=160, num_agents=3, num_bots=4, num_viruses=5, num_pellets=300, arena_size=150.0,
                                    view_size=100.0, grid_size=32, obs_flags=31, start_mass=40.0, virus_pop_mass=60.0, virus_mass=30.0,
                                    auto_reset=1, speed_k=8.0),
        "single_player": dict(num_arenas=300, num_agents=1, num_bots=0, num_pellets=200, arena_size=90.0, view_size=128.0,
                              grid_size=64, obs_flags=7),
        "dense_overflow": dict(num_arenas=1100, num_agents=1, num_bots=1, num_pellets=900, arena_size=200.0, view_size=128.0,
                               grid_size=64, obs_flags=7, speed_k=8.0),
        "cfg2": dict(bench.WORLDS["cfg2"], num_arenas=1024),
    }[world]
    a = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=33)
    b = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=33)
    a.reset(); b.reset()
    b.host_obs_persistent(True)
    N, A = kw["num_arenas"], kw["num_agents"]
    rng = np.random.default_rng(2)
    h_obs = torch.empty(b.obs_shape, dtype=torch.float32).pin_memory()
    other = torch.empty(b.obs_shape, dtype=torch.float32).pin_memory()
    h_obs.fill_(7.0); other.fill_(-3.0)  # stale content before the first delivery into each buffer
    h_r = torch.empty((N, A), dtype=torch.float32).pin_memory(); h_d = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    partial = 0
    for t in range(60):
        xy = rng.uniform(-1, 1, (N, A, 2)).astype(f32)
        xy[: N // 2] = np.sign(xy[: N // 2, :, :1]) * np.array([1.0, 0.3], f32)  # half of the arenas run into the side walls
        act = rng.integers(0, 3, (N, A)).astype(np.int32)
        if t == 20:  # a masked reset on both sides: the device state jumps, the host buffer still holds step 19's grids
            m = torch.from_numpy((rng.random(N) < 0.3).astype(np.uint8)).cuda()
            a.reset(mask=m); b.reset(mask=m)
        if t == 30:
            b.host_obs_persistent(False)
        if t == 33:
            b.host_obs_persistent(True)
        o, r, d, _ = a.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        buf = other if t in (10, 11, 40) else h_obs  # another pointer: full rewrite there, and again on the way back
        b.step_host(xy, act, buf.numpy(), h_r.numpy(), h_d.numpy())
        assert np.array_equal(buf.numpy().view(np.int32), o.cpu().numpy().view(np.int32)), f"{world}: obs differ at step {t}"
        assert np.array_equal(h_r.numpy(), r.cpu().numpy()) and np.array_equal(h_d.numpy(), d.cpu().numpy())
        occ = ((o != 0) & (o != -1)).flatten(2).sum(2)
        partial += int(((occ > 256).any() and (occ <= 256).any()).item())
    if world == "dense_overflow":
        assert partial > 10, "expected steps with both listed and overflowed images"
    a.close(); b.close()


def test_coordinator_with_hundreds_of_workers_updates_its_buffer_in_place():
    """a2c/coordinator.py semantics at a batch size where agar_step_host takes the compact copy: the Coordinator owns its
    pinned observation buffer, hands out copies and declares the buffer persistent; every live env's observation must
    equal the device-resident twin's, step after step."""
    import agarai_b200 as ab
    kw = dict(num_agents=1, multi_agent=False, difficulty="normal", num_pellets=300, num_viruses=3, num_bots=6, arena_size=120.0,
              grid_size=32, observe_viruses=True, start_mass=12.0)
    W = 320
    twin = ab.BatchedAgarioEnv(ab.config_from_gym_kwargs(num_envs=W, auto_reset=0, **kw), seed=4)
    rng = np.random.default_rng(3)
    with ab.Coordinator(lambda: kw, W, seed=4) as coord:
        obs = coord.reset()
        o_dev = twin.reset().cpu().numpy()
        assert all(np.array_equal(obs[i], o_dev[i, 0]) for i in range(W))
        for t in range(40):
            xy = rng.uniform(-1, 1, (W, 2)).astype(f32); act = rng.integers(0, 3, W).astype(np.int32)
            live = [not d for d in coord.dones]
            xy_d = np.where(np.array(live)[:, None], xy, 0).astype(f32); act_d = np.where(live, act, 0).astype(np.int32)
            obs, rs, dones, _ = coord.step([(float(xy[i, 0]), float(xy[i, 1]), int(act[i])) for i in range(W)])
            o, r, d, _ = twin.step(torch.from_numpy(xy_d.reshape(W, 1, 2)).cuda(), torch.from_numpy(act_d.reshape(W, 1)).cuda())
            o = o.cpu().numpy(); r = r.cpu().numpy()
            for i in range(W):
                if live[i]:
                    assert np.array_equal(obs[i].view(np.int32), o[i, 0].view(np.int32)) and rs[i] == r[i, 0], (t, i)
                else:
                    assert obs[i] is None
    twin.close()


@pytest.mark.parametrize("world", ["borders_multi_agent", "single_player", "dense_overflow"])
def test_step_host_compact_copy_rebuilds_identical_grids(world):
    """agar_step_host with batches of >= 256 agents copies per-agent lists of occupied bins + centroids and rebuilds the
    dense grids on host threads: the host buffer must equal, bit for bit, what agar_step wrote on the device (sentinel
    rows / columns at the borders, the K == 1 `others` channel without sentinel, agents without cells, and images
    fuller than the list capacity, which are copied whole)."""
    import agarai_b200 as ab
    kw = {
        "borders_multi_agent": dict(num_arenas=160, num_agents=3, num_bots=4, num_viruses=5, num_pellets=300, arena_size=150.0,
                                    view_size=100.0, grid_size=32, obs_flags=31, start_mass=40.0, virus_pop_mass=60.0, virus_mass=30.0),
        "single_player": dict(num_arenas=300, num_agents=1, num_bots=0, num_pellets=200, arena_size=90.0, view_size=128.0,
                              grid_size=64, obs_flags=7),
        "dense_overflow": dict(num_arenas=1100, num_agents=1, num_bots=1, num_pellets=4000, arena_size=200.0, view_size=128.0,
                               grid_size=64, obs_flags=7),
    }[world]
    a = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=21)
    b = ab.BatchedAgarioEnv(ab.default_config(**kw), seed=21)
    a.reset(); b.reset()
    N, A = kw["num_arenas"], kw["num_agents"]
    rng = np.random.default_rng(1)
    h_obs = torch.empty(b.obs_shape, dtype=torch.float32).pin_memory()
    h_r = torch.empty((N, A), dtype=torch.float32).pin_memory(); h_d = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    for t in range(25):
        xy = rng.uniform(-1, 1, (N, A, 2)).astype(f32)
        act = rng.integers(0, 3, (N, A)).astype(np.int32)
        o, r, d, _ = a.step(torch.from_numpy(xy).cuda(), torch.from_numpy(act).cuda())
        h_obs.fill_(7.0)  # stale content must be overwritten everywhere
        b.step_host(xy, act, h_obs.numpy(), h_r.numpy(), h_d.numpy())
        assert np.array_equal(h_obs.numpy().view(np.int32), o.cpu().numpy().view(np.int32)), f"{world}: obs differ at step {t}"
        assert np.array_equal(h_r.numpy(), r.cpu().numpy()) and np.array_equal(h_d.numpy(), d.cpu().numpy())
    h2d, d2h = b.host_copy_bytes()
    dense = h_obs.numel() * 4
    assert h2d == N * A * 12
    if world == "dense_overflow":
        assert d2h > dense // 2, "expected most images to overflow the list and be copied whole"
    else:
        assert d2h < dense // 4, (d2h, dense)
    a.close(); b.close()
