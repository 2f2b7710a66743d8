"""Host-side mirror of the reference's env interface on top of the C-ABI CUDA library.

`BatchedAgarioEnv` is N arenas stepped by one kernel launch, with torch CUDA tensors in and out
(zero-copy: the library writes straight into them).  `GymCompatEnv` is the per-arena view with
the exact calling convention of the reference's call sites:

    env = gym.make(id, **kwargs)          configuration/__init__.py:45-73, ppo/train.py:361-372
    obs = env.reset()                     ppo/train.py:164, a2c/training.py:66, dqn/training.py:96
    obs, rewards, dones, info = env.step(actions)      ppo/train.py:173, a2c/training.py:94
    env.observation_space.shape, env.action_space, env.render(), env.close()

PyTorch is used for device memory and streams only; all compute is in libagar_b200.so and there
is no fallback path: without the library (or without a GPU) construction raises.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .actions import ActionConfig, ppo_action_table

OBS_PELLETS, OBS_CELLS, OBS_OTHERS, OBS_VIRUSES, OBS_FOODS = 1, 2, 4, 8, 16

# gym ids the reference uses (configuration/__init__.py:38-42, dqn/hyperparameters.py:89); "full" hands out the raw
# entity lists (`FullObservation`) that the reference's extractors consume (dqn/training.py:252-256)
GYM_IDS = ("agario-grid-v0", "agario-ram-v0", "agario-full-v0")
# screen frames are drawn by the external package's renderer (only their grayscale reduction is in the
# reference tree: features.ScreenFeatureExtractor)
_UNSUPPORTED_IDS = ("agario-screen-v0",)

# `difficulty` only selects defaults in the original package; explicit kwargs win
_DIFFICULTY_DEFAULTS = {
    "normal": dict(num_pellets=1000, num_viruses=25, num_bots=25),
    "empty": dict(num_pellets=1000, num_viruses=0, num_bots=0),
    "trivial": dict(num_pellets=1000, num_viruses=0, num_bots=0),
}


class Box:
    """Minimal picklable stand-in for gym.spaces.Box (observation_space / action_space are sent
    through pipes by a2c/remote_environment.py:78-81, so they must pickle)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def __repr__(self):
        return f"Discrete({self.n})"


class TupleSpace:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __repr__(self):
        return f"Tuple{self.spaces}"


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def config_from_gym_kwargs(num_envs=1, **kw) -> "_lib.AgarConfig":
    """Maps the reference's gym.make kwargs (configuration/__init__.py:53-71; a2c/train.py:17-64;
    old style dqn/train.py:57-67) onto an AgarConfig.  Unknown keys raise like gym would; any
    AgarConfig field may also be passed directly to override this build's rule constants.
    `num_frames` (environment.proto:44) is not an engine field: it is validated here and applied by
    BatchedAgarioEnv(num_frames=...) / make(), which stack the last k grids of every agent."""
    kw = dict(kw)
    difficulty = str(kw.pop("difficulty", "normal")).lower()
    if difficulty not in _DIFFICULTY_DEFAULTS:
        raise ValueError(f"unknown difficulty {difficulty!r}")
    fields = dict(_DIFFICULTY_DEFAULTS[difficulty])
    kw.pop("multi_agent", None)  # always multi-agent capable; the compat view handles A == 1
    if "frames_per_step" in kw:  # old-style name (dqn/train.py:58)
        kw.setdefault("ticks_per_step", kw.pop("frames_per_step"))
    num_frames = int(kw.pop("num_frames", 1))
    if not 1 <= num_frames <= 16:
        raise ValueError("num_frames must be in [1, 16] (environment.proto:44 default is 1)")
    flags = 0
    for key, bit in (("observe_pellets", OBS_PELLETS), ("observe_cells", OBS_CELLS),
                     ("observe_others", OBS_OTHERS), ("observe_viruses", OBS_VIRUSES),
                     ("observe_foods", OBS_FOODS)):
        default = key != "observe_foods"  # environment.proto:51-54 defaults are all true
        if bool(kw.pop(key, default)):
            flags |= bit
    fields["obs_flags"] = flags
    for key in ("num_agents", "ticks_per_step", "num_pellets", "num_viruses", "num_bots", "grid_size"):
        if key in kw:
            fields[key] = int(kw.pop(key))
    if "arena_size" in kw:
        fields["arena_size"] = float(kw.pop("arena_size"))
    if "pellet_regen" in kw:
        fields["pellet_regen"] = int(bool(kw.pop("pellet_regen")))
    fields["num_arenas"] = int(num_envs)
    fields.update(kw)  # remaining keys must be AgarConfig fields (default_config raises otherwise)
    return _lib.default_config(**fields)


class BatchedAgarioEnv:
    """N independent arenas on one GPU behind the reference's env interface.

    step(target_xy[N,A,2] float32 cuda, act[N,A] int32 cuda) -> (obs[N,A,C,G,G], reward[N,A],
    done[N,A] uint8, info).  Returned tensors are freshly allocated unless `out_obs` is given,
    because callers keep observations for a whole episode (rollout/multi_rollout.py:42-44).
    """

    metadata = {"render.modes": []}

    def __init__(self, cfg: "_lib.AgarConfig", device=None, seed: int = 0, action_config: ActionConfig = None,
                 observation: str = "grid", ram_extractor=None, lib_path: Optional[str] = None, num_frames: int = 1):
        self._h = None
        if observation not in ("grid", "ram"):
            raise ValueError("observation must be 'grid' or 'ram'")
        if not 1 <= int(num_frames) <= 16 or (num_frames != 1 and observation != "grid"):
            raise ValueError("num_frames must be in [1, 16] and applies to the grid observation")
        self.num_frames = int(num_frames)
        self.observation = observation
        self._L = _lib.lib() if lib_path is None else _lib.load(lib_path)  # lib_path: a build.build_specialised() copy
        if not torch.cuda.is_available():
            raise _lib.AgarError("BatchedAgarioEnv needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        _lib.check(self._L.agar_create(C.byref(cfg), self.device.index or 0, C.byref(h)), self._L)
        self._h = h
        self.cfg = _lib.AgarConfig()
        _lib.check(self._L.agar_get_config(self._h, C.byref(self.cfg)), self._L)
        self.layout = _lib.AgarLayout()
        _lib.check(self._L.agar_get_layout(self._h, C.byref(self.layout)), self._L)
        self.num_envs, self.num_agents = self.cfg.num_arenas, self.cfg.num_agents
        self.channels = self._L.agar_obs_channels(C.byref(self.cfg))
        G = self.cfg.grid_size
        # num_frames = k (environment.proto:44, own spec: the external package is absent): an observation stacks the
        # grids of the agent's last k steps along the channel axis, oldest first; after reset() all k are the first
        # grid.  The kernel writes each step's grid into one slot of a device ring [k, N, A, C, G, G]; the stacked
        # tensor handed out is a fresh copy (callers keep observations, rollout/multi_rollout.py:42-44).
        self.frame_shape = (self.num_envs, self.num_agents, self.channels, G, G)
        self.obs_shape = (self.num_envs, self.num_agents, self.num_frames * self.channels, G, G)
        self._ring, self._ring_pos = None, 0
        # per-agent spaces, as the reference's consumers read them (dqn/qn.py:50 takes shape[0] as
        # channels => CHW, which is also GridFeatureExtractor.shape, features/extractors.py:31)
        self.observation_space = Box(-1.0, np.inf, (self.num_frames * self.channels, G, G))
        self.ram_extractor = None
        if observation == "ram":  # agario-ram-v0: FeatureExtractor vectors (features/extractors.py:99-149)
            from .features import FeatureExtractor
            self.ram_extractor = ram_extractor or FeatureExtractor()
            self.obs_shape = (self.num_envs, self.num_agents, self.ram_extractor.size)
            self.observation_space = Box(-np.inf, np.inf, self.ram_extractor.shape)
        self.action_space = TupleSpace((Box(-1.0, 1.0, (2,)), Discrete(3)))
        self.action_config = action_config or ActionConfig()
        self.set_action_table(*ppo_action_table(self.action_config))
        self.seed(seed)

    # ---- lifecycle -----------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._L.agar_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed: int):
        _lib.check(self._L.agar_seed(self._h, C.c_uint64(int(seed) & (2 ** 64 - 1))), self._L)
        self.seed_value = int(seed)
        return [seed]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _new_obs(self):
        return torch.empty(self.obs_shape, dtype=torch.float32, device=self.device)

    # ---- frame ring (num_frames > 1) ------------------------------------------------------
    def _frame_slot(self, advance: bool):
        """the ring slot the kernel writes this call's grid into"""
        if self._ring is None:
            self._ring = torch.zeros((self.num_frames,) + self.frame_shape, dtype=torch.float32, device=self.device)
        if advance:
            self._ring_pos = (self._ring_pos + 1) % self.num_frames
        return self._ring[self._ring_pos]

    def _stacked(self, out_obs):
        """[N, A, k*C, G, G]: the ring's slots oldest first, copied out"""
        k = self.num_frames
        order = [(self._ring_pos + 1 + i) % k for i in range(k)]
        out = self._new_obs() if out_obs is None else out_obs
        N, A, C, G = self.num_envs, self.num_agents, self.channels, self.cfg.grid_size
        torch.stack([self._ring[j] for j in order], dim=2, out=out.view(N, A, k, C, G, G))
        return out

    # ---- reference interface --------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None, out_obs: Optional[torch.Tensor] = None, want_obs: bool = True):
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        ram = self.ram_extractor is not None
        if self.num_frames > 1:
            frame = self._frame_slot(advance=False)
            _lib.check(self._L.agar_reset(self._h, _ptr(mask), _ptr(frame), self._stream()), self._L)
            sel = slice(None) if mask is None else mask.bool()
            for j in range(self.num_frames):  # a reset arena starts with k copies of its first grid
                if j != self._ring_pos:
                    self._ring[j][sel] = frame[sel]
            return self._stacked(out_obs)
        obs = (self._new_obs() if out_obs is None else out_obs) if want_obs else None
        _lib.check(self._L.agar_reset(self._h, _ptr(mask), None if (ram or obs is None) else _ptr(obs), self._stream()), self._L)
        return self.ram_extractor.extract(self, out=obs) if (ram and obs is not None) else obs

    def _outs(self, out_obs, want_obs):
        """(buffer the kernel rasterises into, reward, done); with num_frames > 1 the buffer is the next ring slot"""
        if self.num_frames > 1:
            obs = self._frame_slot(advance=True)
        else:
            obs = (self._new_obs() if out_obs is None else out_obs) if want_obs else None
        reward = torch.empty((self.num_envs, self.num_agents), dtype=torch.float32, device=self.device)
        done = torch.empty((self.num_envs, self.num_agents), dtype=torch.uint8, device=self.device)
        return obs, reward, done

    def _obs_out(self, obs, out_obs, want_obs):
        if self.num_frames > 1:
            return self._stacked(out_obs) if want_obs else None
        if self.ram_extractor is not None and obs is not None:
            self.ram_extractor.extract(self, out=obs)
        return obs

    def step(self, target_xy: torch.Tensor, act: torch.Tensor, out_obs=None, want_obs=True):
        t = target_xy.to(device=self.device, dtype=torch.float32).contiguous()
        a = act.to(device=self.device, dtype=torch.int32).contiguous()
        if t.numel() != self.num_envs * self.num_agents * 2 or a.numel() != self.num_envs * self.num_agents:
            raise ValueError("step: actions must be [num_envs, num_agents, 2] and [num_envs, num_agents]")
        obs, reward, done = self._outs(out_obs, want_obs)
        ram = self.ram_extractor is not None
        _lib.check(self._L.agar_step(self._h, _ptr(t), _ptr(a), None if ram else _ptr(obs), _ptr(reward), _ptr(done),
                                     self._stream()), self._L)
        return self._obs_out(obs, out_obs, want_obs), reward, done, {}

    def set_action_table(self, xy: np.ndarray, act: np.ndarray):
        xy = np.ascontiguousarray(xy, np.float32)
        act = np.ascontiguousarray(act, np.int32)
        _lib.check(self._L.agar_set_action_table(self._h, xy.ctypes.data_as(C.c_void_p),
                                                 act.ctypes.data_as(C.c_void_p), len(act)), self._L)
        self.num_actions = len(act)

    def step_discrete(self, action_index: torch.Tensor, out_obs=None, want_obs=True):
        """env.step(vmap(action_converter)(indices)) of ppo/train.py:172-173 in one launch."""
        idx = action_index.to(device=self.device, dtype=torch.int32).contiguous()
        if idx.numel() != self.num_envs * self.num_agents:
            raise ValueError("step_discrete: indices must be [num_envs, num_agents]")
        obs, reward, done = self._outs(out_obs, want_obs)
        ram = self.ram_extractor is not None
        _lib.check(self._L.agar_step_indexed(self._h, _ptr(idx), None if ram else _ptr(obs), _ptr(reward), _ptr(done),
                                             self._stream()), self._L)
        return self._obs_out(obs, out_obs, want_obs), reward, done, {}

    def observe(self, out_obs=None):
        if self.num_frames > 1:  # the current grid is already in the ring
            return self._stacked(out_obs)
        obs = self._new_obs() if out_obs is None else out_obs
        if self.ram_extractor is not None:
            return self.ram_extractor.extract(self, out=obs)
        _lib.check(self._L.agar_observe(self._h, _ptr(obs), self._stream()), self._L)
        return obs

    def observe_records(self, state: torch.Tensor):
        """Rasterise state snapshots ([n, layout.words] float32) kept instead of observations (one grid per
        snapshot: [n, A, C, G, G] whatever num_frames is)."""
        s = state.to(device=self.device, dtype=torch.float32).contiguous()
        n = s.shape[0]
        G = self.cfg.grid_size
        obs = torch.empty((n, self.num_agents, self.channels, G, G), dtype=torch.float32, device=self.device)
        _lib.check(self._L.agar_observe_records(self._h, _ptr(s), n, _ptr(obs), self._stream()), self._L)
        return obs

    def render(self, mode="human"):
        return None  # dqn/training.py:106 calls it every step; there is nothing to draw headless

    # ---- state / events (parity tests, checkpoint-resume) ----------------------------------
    def get_state(self) -> torch.Tensor:
        s = torch.empty((self.num_envs, self.layout.words), dtype=torch.float32, device=self.device)
        _lib.check(self._L.agar_get_state(self._h, _ptr(s), self._stream()), self._L)
        return s

    def set_state(self, state: torch.Tensor):
        s = state.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(s.shape) != (self.num_envs, self.layout.words):
            raise ValueError("set_state: wrong shape")
        _lib.check(self._L.agar_set_state(self._h, _ptr(s), self._stream()), self._L)
        torch.cuda.current_stream(self.device).synchronize()  # `s` may be a temporary

    def events(self):
        cap = self.cfg.event_capacity
        ev = torch.zeros((self.num_envs, max(cap, 1), 4), dtype=torch.int32, device=self.device)
        cnt = torch.zeros((self.num_envs,), dtype=torch.int32, device=self.device)
        _lib.check(self._L.agar_events(self._h, _ptr(ev), _ptr(cnt), self._stream()), self._L)
        return ev, cnt

    def step_host(self, target_xy: np.ndarray, act: np.ndarray, obs: Optional[np.ndarray],
                  reward: np.ndarray, done: np.ndarray):
        """agar_step_host: numpy (ideally pinned) buffers in and out, copies inside."""
        for arr, dt in ((target_xy, np.float32), (act, np.int32), (reward, np.float32), (done, np.uint8)):
            if arr.dtype != dt or not arr.flags.c_contiguous:
                raise ValueError("step_host: wrong dtype or non-contiguous buffer")
        if obs is not None and (obs.dtype != np.float32 or not obs.flags.c_contiguous or obs.size != int(np.prod(self.obs_shape))):
            raise ValueError("step_host: bad obs buffer")
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        if self.num_frames > 1 and obs is not None:  # stacked grids: the device ring, then one copy
            o, r, d, _ = self.step(torch.from_numpy(target_xy), torch.from_numpy(act))
            torch.from_numpy(obs).copy_(o.reshape(obs.shape))
            torch.from_numpy(reward).copy_(r); torch.from_numpy(done).copy_(d)
            return
        if self.ram_extractor is not None and obs is not None:  # feature vectors: a second launch, then one copy
            _lib.check(self._L.agar_step_host(self._h, vp(target_xy), vp(act), None, vp(reward), vp(done)), self._L)
            torch.from_numpy(obs).copy_(self.ram_extractor.extract(self).reshape(obs.shape))
            return
        _lib.check(self._L.agar_step_host(self._h, vp(target_xy), vp(act), vp(obs), vp(reward), vp(done)), self._L)

    def host_obs_persistent(self, on: bool = True):
        """agar_host_obs_persistent: promise that consecutive step_host calls receive the same, untouched `obs` array;
        the library then rewrites only what changed between two steps (bit-identical grids, several times less host
        memory traffic).  Use it when the env owns its observation buffer, as vector envs do."""
        _lib.check(self._L.agar_host_obs_persistent(self._h, 1 if on else 0), self._L)

    def host_copy_bytes(self):
        """(host->device, device->host) bytes of the last step_host call"""
        h2d, d2h = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.agar_host_copy_bytes(self._h, C.byref(h2d), C.byref(d2h)), self._L)
        return int(h2d.value), int(d2h.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.agar_launch_count(self._h))

    @property
    def kernel_variant(self) -> str:
        """"generic" or the name of the compile-time specialised step kernel this env runs"""
        return self._L.agar_kernel_variant(self._h).decode()


class GymCompatEnv:
    """One arena of a BatchedAgarioEnv with the reference's calling convention.

    Multi-agent (`num_agents` > 1 or multi_agent=True): `step(actions)` takes a list of
    `(np.ndarray[2], int)` (ppo/train.py:172-173, a2c/train.py:79) and returns
    `(obs list, rewards array, dones array, info)`; dones works as a boolean mask
    (ppo/train.py:192) and with all() (:166).  Single-agent (DQN): `step((x, y, act))` returns
    `(obs, float reward, bool done, info)` (dqn/training.py:103-125).  Observations are fresh
    numpy arrays each step (callers keep them, rollout/multi_rollout.py:42-44); a done agent's
    observation is all zeros where the reference extractor would give None
    (features/extractors.py:42)."""

    def __init__(self, env: BatchedAgarioEnv, multi_agent: bool = True, full: bool = False):
        if env.num_envs != 1:
            raise ValueError("GymCompatEnv wraps a single-arena BatchedAgarioEnv (num_envs=1)")
        self.env = env
        self.multi_agent = multi_agent
        # `agario-full-v0` (dqn/hyperparameters.py:89): observations are `FullObservation` entity lists, which the
        # caller feeds to a feature extractor (dqn/training.py:252-256); no grid is rasterised by the step itself
        self.full = full
        self.observation_space = None if full else env.observation_space
        self.action_space = env.action_space
        A = env.num_agents
        # pinned staging so the per-step copies are single async DMA transfers
        self._t = torch.zeros((1, A, 2), dtype=torch.float32).pin_memory()
        self._a = torch.zeros((1, A), dtype=torch.int32).pin_memory()
        self._obs = torch.zeros(env.obs_shape, dtype=torch.float32).pin_memory()
        self._r = torch.zeros((1, A), dtype=torch.float32).pin_memory()
        self._d = torch.zeros((1, A), dtype=torch.uint8).pin_memory()

    def seed(self, seed):
        return self.env.seed(seed)

    def _obs_out(self, obs_np):
        if self.multi_agent:
            return [obs_np[0, i].copy() for i in range(self.env.num_agents)]
        return obs_np[0, 0].copy()

    def _full_out(self):
        from .features import full_observation
        rec = self.env.get_state()[0].cpu().numpy()
        if self.multi_agent:
            return [full_observation(self.env, 0, i, record=rec) for i in range(self.env.num_agents)]
        return full_observation(self.env, 0, 0, record=rec)

    def reset(self):
        if self.full:
            self.env.reset(want_obs=False)
            return self._full_out()
        obs = self.env.reset()
        return self._obs_out(obs.cpu().numpy())

    def step(self, actions):
        A = self.env.num_agents
        if self.multi_agent:
            if len(actions) != A:
                raise ValueError(f"expected {A} actions, got {len(actions)}")
            for i, (xy, act) in enumerate(actions):
                self._t[0, i, 0], self._t[0, i, 1] = float(xy[0]), float(xy[1])
                self._a[0, i] = int(act)
        else:
            x, y, act = actions
            self._t[0, 0, 0], self._t[0, 0, 1] = float(x), float(y)
            self._a[0, 0] = int(act)
        self.env.step_host(self._t.numpy(), self._a.numpy(), None if self.full else self._obs.numpy(), self._r.numpy(),
                           self._d.numpy())
        obs = self._full_out() if self.full else self._obs_out(self._obs.numpy())
        if self.multi_agent:
            return obs, self._r.numpy()[0].copy(), self._d.numpy()[0].astype(bool), {}
        return obs, float(self._r[0, 0]), bool(self._d[0, 0]), {}

    def render(self, mode="human"):
        return None

    def close(self):
        self.env.close()


def make(env_id: str = "agario-grid-v0", num_envs: int = 1, device=None, seed: int = 0,
         action_config: ActionConfig = None, gym_compat: Optional[bool] = None, **kwargs):
    """gym.make(env_id, **kwargs) for the batched engine.  With num_envs == 1 (and gym_compat not
    False) the reference-shaped per-arena view is returned; otherwise the batched tensor env."""
    if env_id in _UNSUPPORTED_IDS:
        raise NotImplementedError(f"{env_id}: frames come from the external package's renderer; the grid, ram and full "
                                  "observation families are served")
    if env_id not in GYM_IDS:
        raise ValueError(f"unknown env id {env_id!r}")
    # the reference's multi-agent call sites pass multi_agent=True explicitly (configuration/__init__.py:54); the
    # old-style single-agent ones (dqn/train.py:57-81) pass neither key and then call env.step((x, y, act))
    multi_agent = bool(kwargs.get("multi_agent", int(kwargs.get("num_agents", 1)) > 1))
    ram_extractor = kwargs.pop("ram_extractor", None)
    specialise = bool(kwargs.pop("specialise", False))
    num_frames = int(kwargs.get("num_frames", 1))
    if env_id == "agario-full-v0":
        for key in ("grid_size", "num_frames", "observe_cells", "observe_others", "observe_viruses", "observe_pellets"):
            kwargs.pop(key, None)  # observation kwargs of the other families have no meaning here
        num_frames = 1
    cfg = config_from_gym_kwargs(num_envs=num_envs, **kwargs)
    lib_path = None
    if specialise and cfg.event_capacity == 0:  # compile (once) a step kernel with this world's sizes as constants
        from . import build as _build
        lib_path = _build.build_specialised(cfg)
    env = BatchedAgarioEnv(cfg, device=device, seed=seed, action_config=action_config,
                           observation="ram" if env_id == "agario-ram-v0" else "grid", ram_extractor=ram_extractor,
                           lib_path=lib_path, num_frames=num_frames)
    if gym_compat is None:
        gym_compat = num_envs == 1
    if env_id == "agario-full-v0" and not gym_compat:
        raise ValueError("agario-full-v0 hands out host-side entity lists: use it through the per-arena view (num_envs=1)")
    return GymCompatEnv(env, multi_agent=multi_agent, full=env_id == "agario-full-v0") if gym_compat else env
