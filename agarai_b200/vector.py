"""Vector-env layer with the semantics of the reference's process fan-out (a2c/coordinator.py:17-100,
a2c/remote_environment.py:20-83), on ONE batched engine instead of one OS process per environment.

The reference starts `num_workers` processes, each owning one env, and scatters/gathers over pipes; a CUDA
context cannot be forked and the point of the engine is that all arenas advance in one kernel launch, so
`Coordinator` here owns a `BatchedAgarioEnv` with `num_workers` arenas and keeps the caller-visible contract:

    with Coordinator(get_env, num_workers) as c:
        obs = c.reset()                                  # list over envs                     (:68-74)
        obs, rs, dones, infos = c.step(actions)          # None entries for finished envs,    (:43-66)
                                                         # dones sticky until reset()
        c.observation_space(); c.action_space()          # methods, as in the reference       (:83-89)

`get_env` is the reference's closure argument.  It cannot return a live env here (envs are arenas of one
batch), so it must return the gym.make keyword dict the env would have been built with (or be that dict).
Arena i of the batch is bit-identical to a stand-alone env created with `arena_id_base=i` and the same seed:
the RNG is keyed by the global arena id (tests/test_gpu_parity.py).
"""
from enum import Enum

import numpy as np
import torch

from .env import BatchedAgarioEnv, config_from_gym_kwargs


class RemoteCommand(Enum):  # a2c/remote_environment.py:12-17 (kept for callers that import it)
    step = 1
    reset = 2
    close = 3
    observation_space = 4
    action_space = 5


def _kwargs_of(get_env):
    kw = get_env() if callable(get_env) else get_env
    if not isinstance(kw, dict):
        raise TypeError("get_env must be (or return) the dict of gym.make keyword arguments of one env")
    return dict(kw)


class Coordinator:
    """a2c/coordinator.py:17-100 on the batched engine.  Single-agent envs (the reference's use: `done` is a
    bool) and multi-agent envs (an env is done when all its agents are) are both supported."""

    def __init__(self, get_env, num_workers, device=None, seed=0):
        self.get_env = get_env
        self.num_workers = int(num_workers)
        self.device, self.seed = device, seed
        self.env = None
        self.dones = [False] * self.num_workers

    # ---- lifecycle (open/close, context manager: coordinator.py:27-41, 76-81) ----
    def open(self):
        kw = _kwargs_of(self.get_env)
        self.multi_agent = bool(kw.get("multi_agent", int(kw.get("num_agents", 1)) > 1))
        kw.setdefault("auto_reset", 0)  # a finished env stays finished until reset(), as a worker's env does
        cfg = config_from_gym_kwargs(num_envs=self.num_workers, **kw)
        self.env = BatchedAgarioEnv(cfg, device=self.device, seed=self.seed, num_frames=int(kw.get("num_frames", 1)))
        N, A = self.num_workers, self.env.num_agents
        self._t = torch.zeros((N, A, 2), dtype=torch.float32).pin_memory()
        self._a = torch.zeros((N, A), dtype=torch.int32).pin_memory()
        self._obs = torch.zeros(self.env.obs_shape, dtype=torch.float32).pin_memory()
        self._r = torch.zeros((N, A), dtype=torch.float32).pin_memory()
        self._d = torch.zeros((N, A), dtype=torch.uint8).pin_memory()
        # _obs is private to this object and only step_host writes to it (callers get copies, _env_obs): the library
        # may update it in place from step to step instead of rewriting every grid
        self.env.host_obs_persistent(True)

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def close(self):
        if self.env is not None:
            self.env.close()
            self.env = None

    # ---- stepping ----
    def _env_obs(self, obs_np, i):
        if self.multi_agent:
            return [obs_np[i, a].copy() for a in range(self.env.num_agents)]
        return obs_np[i, 0].copy()

    def step(self, actions):
        """actions[i] is what env i's `step` takes: a list of (xy, act) per agent (multi-agent) or one
        (x, y, act) tuple (single-agent).  Entries of finished envs are ignored (the reference does not even
        send them, coordinator.py:45-48)."""
        actions = list(actions)
        A = self.env.num_agents
        t, a = self._t.numpy(), self._a.numpy()
        for i in range(self.num_workers):
            if self.dones[i]:
                t[i] = 0.0; a[i] = 0
                continue
            if self.multi_agent:
                if len(actions[i]) != A:
                    raise ValueError(f"env {i}: expected {A} actions, got {len(actions[i])}")
                for k, (xy, act) in enumerate(actions[i]):
                    t[i, k, 0], t[i, k, 1], a[i, k] = float(xy[0]), float(xy[1]), int(act)
            else:
                x, y, act = actions[i]
                t[i, 0, 0], t[i, 0, 1], a[i, 0] = float(x), float(y), int(act)
        self.env.step_host(t, a, self._obs.numpy(), self._r.numpy(), self._d.numpy())
        obs_np, r_np, d_np = self._obs.numpy(), self._r.numpy(), self._d.numpy()
        obs, rs, infos = [], [], []
        for i in range(self.num_workers):
            if self.dones[i]:
                obs.append(None); rs.append(None); infos.append(None)
                continue
            obs.append(self._env_obs(obs_np, i))
            if self.multi_agent:
                rs.append(r_np[i].copy())
                self.dones[i] = bool(d_np[i].all())
            else:
                rs.append(float(r_np[i, 0]))
                self.dones[i] = bool(d_np[i, 0])
            infos.append({})
        return obs, rs, self.dones.copy(), infos

    def reset(self):
        self.dones = [False] * self.num_workers
        obs_np = self.env.reset().cpu().numpy()
        return [self._env_obs(obs_np, i) for i in range(self.num_workers)]

    def observation_space(self):
        return self.env.observation_space

    def action_space(self):
        return self.env.action_space


class RemoteEnvironment:
    """a2c/remote_environment.py:20-53: one env behind the same context-manager interface; there is no
    remote process, the env is arena 0 of its own one-arena engine."""

    def __init__(self, get_env, device=None, seed=0):
        self._c = Coordinator(get_env, 1, device=device, seed=seed)

    def __enter__(self):
        self._c.open()
        return self

    def __exit__(self, exc_type, exc_value, tb):
        self._c.close()

    def reset(self):
        return self._c.reset()[0]

    def step(self, actions):
        A = self._c.env.num_agents
        t, a = self._c._t.numpy(), self._c._a.numpy()
        if self._c.multi_agent:
            for k, (xy, act) in enumerate(actions):
                t[0, k, 0], t[0, k, 1], a[0, k] = float(xy[0]), float(xy[1]), int(act)
        else:
            t[0, 0, 0], t[0, 0, 1], a[0, 0] = float(actions[0]), float(actions[1]), int(actions[2])
        c = self._c
        c.env.step_host(t, a, c._obs.numpy(), c._r.numpy(), c._d.numpy())
        obs = c._env_obs(c._obs.numpy(), 0)
        if c.multi_agent:
            return obs, c._r.numpy()[0].copy(), c._d.numpy()[0].astype(bool), {}
        return obs, float(c._r[0, 0]), bool(c._d[0, 0]), {}

    def observation_space(self):
        return self._c.observation_space()

    def action_space(self):
        return self._c.action_space()
