"""Host-side mirror of the reference interface: action decode, gym kwargs mapping, rollout buffer,
arena sharding (world_size-2 gloo).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

import agarai_b200 as ab
from agarai_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- action decode (ppo/action_conversion.py:9-80; a2c/train.py:67-79; dqn/training.py:220-228) ----
def ref_ppo_convert(index, D, M, split, feed):
    """ppo/action_conversion.py:31-78 restated with numpy float32 (the original is jitted jax)."""
    if index == 0:
        return np.array([0, 0], np.float32), 0
    i = index - 1
    num_moves = D * M
    two_pi = np.float32(2 * np.pi)
    if i < num_moves:
        theta = two_pi * np.float32(i % D) / np.float32(D)
        mag = np.float32(1 - (i // D) / (1 + M))
        return np.array([np.cos(theta) * mag, np.sin(theta) * mag], np.float32), 0
    j = i - num_moves
    theta = two_pi * np.float32(j % D) / np.float32(D)
    act = int(1 + j / D) if (split and feed) else (1 if split else (2 if feed else 0))
    return np.array([np.cos(theta), np.sin(theta)], np.float32), act


@pytest.mark.parametrize("D,M,split,feed", [(8, 1, False, False), (8, 1, True, True), (4, 3, True, False),
                                            (16, 2, False, True)])
def test_ppo_action_table(D, M, split, feed):
    cfg = ab.ActionConfig(D, M, split, feed)
    xy, act = ab.ppo_action_table(cfg)
    assert len(act) == ab.action_size(cfg) == 1 + D * M + D * (int(split) + int(feed))
    conv = ab.ActionConverter.ppo(cfg)
    for index in range(len(act)):
        exy, eact = ref_ppo_convert(index, D, M, split, feed)
        gxy, gact = conv(index)
        assert np.array_equal(gxy, exy) and gact == eact
    assert (np.linalg.norm(xy, axis=1) <= 1 + 1e-6).all()
    # ppo.textproto:19-22: 8 directions, 1 magnitude, no acts -> 9 actions
    assert ab.action_size(ab.ActionConfig(8, 1)) == 9


def test_unravel_action_table_matches_a2c_and_dqn_decode():
    shape = (8, 2, 3)
    xy, act = ab.unravel_action_table(shape)
    for index in range(int(np.prod(shape))):
        i0, i1, i2 = np.unravel_index(index, shape)  # a2c/train.py:74-79
        theta = (2 * np.pi * i0) / shape[0]
        mag = 1 - i1 / shape[1]
        assert np.allclose(xy[index], [np.cos(theta) * mag, np.sin(theta) * mag], atol=1e-7)
        assert act[index] == int(i2)


# ---- gym kwargs (configuration/__init__.py:49-73; a2c/train.py:17-64; dqn/train.py:55-81) ----
def test_gym_kwargs_mapping():
    cfg = ab.config_from_gym_kwargs(
        num_envs=7, num_agents=16, multi_agent=True, difficulty="empty", ticks_per_step=4, arena_size=1000,
        num_pellets=1000, num_viruses=0, num_bots=0, pellet_regen=True, grid_size=64, num_frames=1,
        observe_cells=True, observe_others=True, observe_pellets=True, observe_viruses=False)
    assert (cfg.num_arenas, cfg.num_agents, cfg.num_bots, cfg.num_pellets, cfg.num_viruses) == (7, 16, 0, 1000, 0)
    assert cfg.obs_flags == 7 and cfg.grid_size == 64 and cfg.arena_size == 1000.0 and cfg.pellet_regen == 1
    a2c = ab.config_from_gym_kwargs(num_agents=8, difficulty="normal", ticks_per_step=4, arena_size=500,
                                    num_pellets=1000, num_viruses=25, num_bots=25, pellet_regen=True, grid_size=32,
                                    observe_cells=True, observe_others=True, observe_viruses=True, observe_pellets=True)
    assert (a2c.num_bots, a2c.num_viruses, a2c.grid_size, a2c.obs_flags) == (25, 25, 32, 15)
    dqn = ab.config_from_gym_kwargs(frames_per_step=4, arena_size=45, num_pellets=30, num_viruses=0, num_bots=0,
                                    pellet_regen=True, difficulty="empty")
    assert dqn.ticks_per_step == 4 and dqn.num_agents == 1
    with pytest.raises(TypeError):
        ab.config_from_gym_kwargs(no_such_kwarg=1)
    assert ab.config_from_gym_kwargs(num_frames=4).grid_size == 64  # validated here, applied by the env (frame ring)
    with pytest.raises(ValueError):
        ab.config_from_gym_kwargs(num_frames=0)
    with pytest.raises(NotImplementedError):
        ab.make("agario-screen-v0")
    with pytest.raises(ValueError):
        ab.make("no-such-env-v0")


def test_spaces_are_picklable():
    import pickle
    from agarai_b200.env import Box, Discrete, TupleSpace
    sp = TupleSpace((Box(-1.0, 1.0, (2,)), Discrete(3)))
    assert pickle.loads(pickle.dumps(sp)).spaces[1].n == 3  # a2c/remote_environment.py:78-81 ships them


# ---- rollout buffer (rollout/multi_rollout.py:17-64) ----
def test_device_rollout_matches_multirollout_semantics():
    T, B = 5, 6
    roll = ab.DeviceRollout(capacity=T)
    ref = {"rewards": [], "values": [], "observations": []}
    rng = np.random.default_rng(0)
    for t in range(T):
        step = {"rewards": rng.normal(size=B).astype(np.float32), "values": rng.normal(size=B).astype(np.float32),
                "observations": rng.normal(size=(B, 2, 3, 3)).astype(np.float32)}
        roll.record({k: torch.from_numpy(v) for k, v in step.items()})
        for k, v in step.items():
            ref[k].append(v)
    boot = rng.normal(size=B).astype(np.float32)
    roll.record({"values": torch.from_numpy(boot)})  # extra bootstrap row, ppo/train.py:189-195
    ref["values"].append(boot)
    assert len(roll) == T + 1 and roll["values"].shape == (T + 1, B) and roll["rewards"].shape == (T, B)
    for key in ref:  # transpose_batch: [T][A] -> [A][T, ...]
        expect = [np.array(list(r)) for r in zip(*ref[key])]
        got = roll.batched(key)
        assert got.shape[0] == B
        for a in range(B):
            assert np.array_equal(got[a].numpy(), expect[a])
    with pytest.raises(IndexError):
        roll.record({"rewards": torch.zeros(B)}); roll.record({"rewards": torch.zeros(B)})
    roll.clear()
    assert len(roll) == 0


# ---- sharding over ranks: world_size-2 gloo on CPU ----
def _gloo_worker(rank, world, port, out_q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = sharding.shard_from_env(arenas_per_rank=3)
    kw = dict(num_agents=2, num_bots=2, num_viruses=2, arena_size=120.0, num_pellets=200)
    env = O.OracleEnv(O.default_config(num_arenas=sh.arenas_per_rank, arena_id_base=sh.arena_id_base, **kw), seed=5)
    env.reset()
    rng = np.random.default_rng(0)  # same action stream on every rank; each takes its own rows
    for t in range(10):
        xy = rng.uniform(-1, 1, (sh.total_arenas, 2, 2)).astype(np.float32)
        act = rng.integers(0, 3, (sh.total_arenas, 2)).astype(np.int32)
        lo = sh.arena_id_base
        env.step(xy[lo:lo + sh.arenas_per_rank], act[lo:lo + sh.arenas_per_rank], want_obs=False)
    slow = sharding.max_over_ranks(1.0 + rank)           # rank 1 is "slower"
    thr = sharding.aggregate_throughput(units_per_rank=30, seconds_local=1.0 + rank)
    out_q.put((rank, list(sh.global_ids()), env.get_state(), slow, thr))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_is_invariant_and_reduces_max(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4, 5]          # disjoint cover of the global ids
    assert res[0][3] == res[1][3] == 2.0                             # max over ranks
    assert res[0][4] == res[1][4] == pytest.approx(2 * 30 / 2.0)     # whole-job units / slowest rank
    # one process stepping all 6 arenas gives the same states as the two shards
    kw = dict(num_agents=2, num_bots=2, num_viruses=2, arena_size=120.0, num_pellets=200)
    env = oracle.OracleEnv(oracle.default_config(num_arenas=6, **kw), seed=5)
    env.reset()
    rng = np.random.default_rng(0)
    for t in range(10):
        xy = rng.uniform(-1, 1, (6, 2, 2)).astype(np.float32)
        act = rng.integers(0, 3, (6, 2)).astype(np.int32)
        env.step(xy, act, want_obs=False)
    full = env.get_state().view(np.int32)
    assert np.array_equal(full[:3], res[0][2].view(np.int32)) and np.array_equal(full[3:], res[1][2].view(np.int32))


def test_split_total():
    assert sharding.split_total(262144, 8, 3).arena_id_base == 3 * 32768  # BASELINE configs[3]
    with pytest.raises(ValueError):
        sharding.split_total(10, 4, 0)


# ---- PPO losses restated from rl/__init__.py:12-57 (numpy) vs the torch forms in ppo.py ----
def test_ppo_losses_match_reference_formulas():
    from agarai_b200 import ppo
    rng = np.random.default_rng(1)
    B, NA = 64, 9
    logits, old = rng.normal(size=(B, NA)), rng.normal(size=(B, NA))
    actions = rng.integers(0, NA, B); adv = rng.normal(size=B); eps = 0.2

    def log_softmax(x):
        x = x - x.max(1, keepdims=True)
        return x - np.log(np.exp(x).sum(1, keepdims=True))
    new_lp = np.take_along_axis(log_softmax(logits), actions[:, None], 1)[:, 0]
    old_lp = np.take_along_axis(log_softmax(old), actions[:, None], 1)[:, 0]
    ratio = np.exp(new_lp - old_lp)
    expect = (-np.minimum(adv * ratio, adv * np.clip(ratio, 1 - eps, 1 + eps))).mean()   # rl/__init__.py:31-46
    got = ppo.ppo_actor_loss(torch.tensor(logits), torch.tensor(actions), torch.tensor(adv), torch.tensor(old), eps)
    assert abs(float(got) - expect) < 1e-9
    a = rng.normal(size=B) * 2
    huber = np.where(np.abs(a) < 1, a * a / 2, np.abs(a) - 0.5).mean()                  # rl/__init__.py:49-57
    assert abs(float(ppo.critic_loss(torch.tensor(a), torch.zeros(B, dtype=torch.float64))) - huber) < 1e-12
    p = np.exp(log_softmax(logits))
    ent = (-(p * log_softmax(logits)).sum(1)).mean()                                     # rl.xs(softmax, logits)
    assert abs(float(ppo.policy_entropy(torch.tensor(logits))) - ent) < 1e-9
    net = ppo.ActorCritic((3, 64, 64), 9)
    lo, v = net(torch.zeros(2, 3, 64, 64))
    assert lo.shape == (2, 9) and v.shape == (2,)
    assert net.dense.in_features == 52 * 52 * 5   # 13,520 features before the dense layer


# ---- config front-end (configuration/__init__.py:26-73, environment.proto, config.proto) ----
def test_textproto_config_front_end():
    from agarai_b200 import configuration as cf
    cfg = cf.load("ppo")
    env = cfg.environment
    assert cf.gym_env_name(env) == "agario-grid-v0"
    kw = cf.gym_env_config(env)
    assert list(kw) == ["num_agents", "multi_agent", "difficulty", "ticks_per_step", "arena_size", "num_pellets",
                        "num_viruses", "num_bots", "pellet_regen", "grid_size", "num_frames", "observe_cells",
                        "observe_others", "observe_pellets", "observe_viruses"]  # configuration/__init__.py:53-71
    assert kw == dict(num_agents=16, multi_agent=True, difficulty="empty", ticks_per_step=4, arena_size=1000,
                      num_pellets=1000, num_viruses=0, num_bots=0, pellet_regen=False, grid_size=64, num_frames=1,
                      observe_cells=True, observe_others=True, observe_pellets=True, observe_viruses=False)
    hp = cfg.hyperparameters
    assert (hp.num_episodes, hp.episode_length, hp.num_sgd_steps, hp.batch_size) == (10000, 512, 8, 128)
    assert hp.gamma == 0.75 and hp.gae.lam == 0.1 and hp.loss.ppo.clip_epsilon == 0.2
    assert hp.critic_lr.scale == 10.0 and hp.actor_lr.scale == 1.0 and hp.actor_lr.exponential_decay.decay_steps == 500
    assert hp.actor_lr.HasField("exponential_decay") and not hp.actor_lr.HasField("constant") and hp.seed == 0
    assert cfg.model.feature_extractor.num_features == 1024 and not cfg.HasField("test_environment")
    ac = cf.action_config(env)
    assert ab.action_size(ac) == 9  # ppo.textproto:19-22: 1 + 8*1
    # the kwargs feed straight into the engine's config mapping
    ecfg = ab.config_from_gym_kwargs(num_envs=3, **kw)
    assert (ecfg.num_agents, ecfg.num_pellets, ecfg.obs_flags, ecfg.pellet_regen) == (16, 1000, 7, 0)
    # proto2 defaults and text-format details
    m = cf.parse("environment { agario { difficulty: TRIVIAL } }  # configs/test_config.textproto:6-8")
    assert m.environment.agario.difficulty == "TRIVIAL" and m.environment.num_agents == 1
    assert m.environment.observation.grid_size == 64 and m.environment.observation.ticks_per_step == 4
    assert cf.gym_env_config(m.environment)["difficulty"] == "trivial"
    m = cf.parse('environment: < observation < type: 2 pellets: false > >; hyperparameters { gamma: 5e-1, seed: 0x10 }')
    assert cf.gym_env_name(m.environment) == "agario-ram-v0" and m.hyperparameters.gamma == 0.5 and m.hyperparameters.seed == 16
    with pytest.raises(ValueError):
        cf.parse("hyperparameters { learning_rate: 0.1 }")  # configs/sweeps/gamma_lam.textproto:35 is rejected likewise
    with pytest.raises(ValueError):
        cf.parse("environment { agario { difficulty: EASY } }")  # configuration/test.py:13 names a value that does not exist
    with pytest.raises(ValueError):
        cf.parse("environment { num_agents: 2 ")
    ref = "/root/reference/configuration/configs"
    if os.path.isdir(ref):  # build container only: the reference's own files parse to the same kwargs
        rcfg = cf.load(os.path.join(ref, "ppo.textproto"))
        assert cf.gym_env_config(rcfg.environment) == kw and rcfg.hyperparameters.to_dict() == hp.to_dict()
        assert cf.load(os.path.join(ref, "test_config.textproto")).environment.agario.difficulty == "TRIVIAL"
