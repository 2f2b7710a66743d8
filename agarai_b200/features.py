"""Device-side mirrors of the reference's feature extractors (features/extractors.py), same class names,
constructor arguments and `shape` / `size` attributes.  The reference's extractors take a `FullObservation`
of ONE agent on the host.  These take either the batched engine (the state already lives in HBM; the whole
batch comes from one kernel launch) or, with the reference's own call form `extractor(observation)`
(dqn/training.py:252-256), a `FullObservation`: its entity lists are packed into one arena record, uploaded
and run through the same kernels (agar_observe_records / agar_observe_ram).  No fallback: without the CUDA
library they raise.

    GridFeatureExtractor    features/extractors.py:11-96    -> agar_observe          (kernel raster_arena)
    FeatureExtractor        features/extractors.py:99-149   -> agar_observe_ram      (ram_features_kernel)
    ScreenFeatureExtractor  features/extractors.py:152-164  -> agar_screen_gray      (screen_gray_kernel)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


import collections

FullObservation = collections.namedtuple("FullObservation", ["pellets", "viruses", "foods", "agent", "others"])


def _is_full_observation(x):
    return all(hasattr(x, f) for f in ("pellets", "viruses", "foods", "agent", "others"))


def _bucket(n, step, lo):
    return max(lo, -(-int(n) // step) * step)


class _RecordUploader:
    """Packs the entity lists of a `FullObservation` into one arena record and keeps the private one-arena engine
    handles that rasterise such records (one per capacity bucket).  Entities keep their order: the agent's cells take
    the slots of player 0, other player k the slots of player 1+k, so mass sums are taken in the reference's order
    (features/extractors.py:46-56)."""

    def __init__(self, **world):
        self.world = world          # view_size, grid_size, arena_size, obs_flags
        self._envs = {}

    def env_for(self, obs, device=None):
        from .env import BatchedAgarioEnv
        n_p, n_v, n_f, n_o = len(obs.pellets), len(obs.viruses), len(obs.foods), len(obs.others)
        if n_o > 63:
            raise ValueError("FullObservation with more than 63 other players")
        key = (_bucket(n_p, 256, 4), n_v, _bucket(n_f, 64, 4), 1 << max(0, int(n_o)).bit_length(), str(device))
        env = self._envs.get(key)
        if env is None:
            cfg = _lib.default_config(num_arenas=1, num_agents=1, num_bots=key[3] - 1, num_pellets=key[0], num_viruses=key[1],
                                      max_foods=key[2], **self.world)
            env = self._envs[key] = BatchedAgarioEnv(cfg, device=device)
        return env

    @staticmethod
    def pack(env, obs):
        """float32 [words] host record of `env`'s layout holding the observation's entities"""
        L, c = env.layout, env.cfg
        rec = np.zeros(L.words, np.float32)
        rec[L.pellet_x:L.pellet_x + L.P_pad] = -1.0          # x < 0: absent pellet
        rec[L.pellet_y:L.pellet_y + L.P_pad] = -1.0
        rec[L.food_x:L.food_x + L.F_pad] = -1.0              # x < 0: free food slot

        def xy(dst_x, dst_y, ents):
            e = np.asarray(ents, np.float32).reshape(-1, 2) if len(ents) else np.zeros((0, 2), np.float32)
            rec[dst_x:dst_x + len(e)] = e[:, 0]
            rec[dst_y:dst_y + len(e)] = e[:, 1]

        xy(L.pellet_x, L.pellet_y, obs.pellets)
        xy(L.virus_x, L.virus_y, obs.viruses)
        xy(L.food_x, L.food_y, obs.foods)
        for k, cells in enumerate([obs.agent] + list(obs.others)):
            cells = np.asarray(cells, np.float32).reshape(-1, 5) if len(cells) else np.zeros((0, 5), np.float32)
            if len(cells) > 16:
                raise ValueError("a player with more than 16 cells")
            o = k * 16
            for col, seg in enumerate((L.cell_x, L.cell_y, L.cell_ix, L.cell_iy, L.cell_m)):  # 5-wide rows, mass last
                rec[seg + o:seg + o + len(cells)] = cells[:, col]
        return rec

    def upload(self, obs, device=None):
        env = self.env_for(obs, device)
        rec = torch.from_numpy(self.pack(env, obs)).to(env.device)
        return env, rec


class GridFeatureExtractor:
    """features/extractors.py:11-34: the C x G x G grid.  View size, grid size and the channel toggles are
    part of the engine's config (they fix the kernel's shared-memory plan), so this object only checks that
    the env was built with the same ones and exposes `shape`."""

    def __init__(self, view_size, grid_size, arena_size, cells=True, others=True, viruses=True, food=False, flat=False,
                 pellets=True):
        if flat:  # features/extractors.py:33-34 computes tuple(np.prod(shape)), which raises for every input
            raise TypeError("GridFeatureExtractor(flat=True): 'numpy.int64' object is not iterable (as in the reference)")
        self.view_size, self.grid_size, self.arena_size = float(view_size), int(grid_size), float(arena_size)
        self.cells, self.others, self.viruses, self.foods, self.flat = bool(cells), bool(others), bool(viruses), bool(food), False
        self.box_size = view_size / grid_size
        self.flags = (1 if pellets else 0) | (2 if cells else 0) | (4 if others else 0) | (8 if viruses else 0) | (16 if food else 0)
        self.depth = bin(self.flags).count("1")
        self.shape = (self.depth, self.grid_size, self.grid_size)  # features/extractors.py:31
        self._up = _RecordUploader(view_size=self.view_size, grid_size=self.grid_size, arena_size=self.arena_size, obs_flags=self.flags)

    def __call__(self, observation):
        return self.extract(observation)

    def extract(self, observation):
        """`observation` is a `FullObservation` (the reference's call form: returns the agent's float64 [C, G, G] grid
        as a numpy array, or None for an agent without cells, features/extractors.py:42) or a batched engine (returns
        the float32 CUDA tensor [num_envs, num_agents, C, G, G] of every agent)."""
        if _is_full_observation(observation):
            if len(observation.agent) == 0:
                return None
            env, rec = self._up.upload(observation)
            return env.observe_records(rec[None])[0, 0].cpu().numpy().astype(np.float64)
        env = observation
        c = env.cfg
        if (c.grid_size, c.obs_flags) != (self.grid_size, self.flags) or float(c.view_size) != self.view_size \
                or float(c.arena_size) != self.arena_size:
            raise ValueError("GridFeatureExtractor: the env was configured with another grid (grid_size, view_size, "
                             "arena_size and the channel toggles are AgarConfig fields)")
        return env.observe()


class FeatureExtractor:
    """features/extractors.py:99-149: fixed-length vector of the nearest entities (580 floats by default)."""

    def __init__(self, num_pellet=50, num_virus=5, num_food=10, num_other=5, num_cell=15):
        self.num_pellet, self.num_virus, self.num_food = num_pellet, num_virus, num_food
        self.num_other, self.num_cell = num_other, num_cell
        self.size = 2 * num_pellet + 2 * num_virus + 2 * num_food + 5 * (1 + num_other) * num_cell
        self.shape = (self.size,)
        self.filler_value = -1000
        self._up = _RecordUploader(view_size=128.0, grid_size=4, arena_size=1000.0, obs_flags=1)  # the grid fields are unused here

    def __call__(self, observation):
        return self.extract(observation)

    def extract(self, env, out=None):
        """For a batched engine: float32 [num_envs, num_agents, size] for the env's current state (zeros for an agent
        without cells, where the reference returns None).  For a `FullObservation` (the reference's call form): the
        agent's float64 [size] vector as a numpy array, or None for an agent without cells."""
        if _is_full_observation(env):
            if len(env.agent) == 0:
                return None
            priv, rec = self._up.upload(env)
            priv.set_state(rec[None])
            return self.extract(priv)[0, 0].cpu().numpy().astype(np.float64)
        L = env._L  # the library that owns the handle (a specialised copy keeps its own error slot)
        if out is None:
            out = torch.empty((env.num_envs, env.num_agents, self.size), dtype=torch.float32, device=env.device)
        _lib.check(L.agar_observe_ram(env._h, self.num_pellet, self.num_virus, self.num_food, self.num_other,
                                      self.num_cell, _ptr(out), _stream(env.device)), L)
        return out


class ScreenFeatureExtractor:
    """features/extractors.py:152-164: colour frames -> grayscale by the mean over the trailing axis."""

    def __init__(self, num_frames, screen_len):
        self.screen_len = screen_len
        self.shape = (num_frames, screen_len, screen_len)

    def __call__(self, observation):
        return self.extract(observation)

    def extract(self, observation: torch.Tensor, dtype=torch.float64):
        """observation: uint8 CUDA tensor [..., 3]; returns [...] in float64 (the reference's dtype, bit-exact)
        or float32."""
        if not isinstance(observation, torch.Tensor) or not observation.is_cuda:
            raise _lib.AgarError("ScreenFeatureExtractor needs a CUDA uint8 tensor (there is no CPU fallback)")
        if observation.dtype != torch.uint8 or observation.shape[-1] != 3:
            raise ValueError("expected uint8 frames with a trailing colour axis of 3")
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("dtype must be float64 or float32")
        x = observation.contiguous()
        out = torch.empty(x.shape[:-1], dtype=dtype, device=x.device)
        _lib.check(_lib.lib().agar_screen_gray(_ptr(x), int(np.prod(x.shape[:-1])), _ptr(out),
                                               1 if dtype == torch.float64 else 0, _stream(x.device)))
        return out


def full_observation(env, arena: int = 0, agent: int = 0, record=None) -> "FullObservation":
    """The raw entity lists of one agent's view, in the shape of the external package's `FullObservation`
    that the reference's extractors read (features/extractors.py:40-64,108,140): present pellets / foods
    [n,2], viruses [V,2], the agent's cells and every other player that has cells as [c,5] rows of
    (x, y, impulse_x, impulse_y, mass).  Host-side view of the device state: what `agario-full-v0` hands out
    (env.GymCompatEnv); `record` is an already downloaded host record of the arena."""
    L, c = env.layout, env.cfg
    rec = env.get_state()[arena].cpu().numpy() if record is None else record
    seg = lambda off, n: rec[off:off + n]
    px, py = seg(L.pellet_x, c.num_pellets), seg(L.pellet_y, c.num_pellets)
    fx, fy = seg(L.food_x, c.max_foods), seg(L.food_y, c.max_foods)
    pellets = np.stack([px, py], 1)[px >= 0]
    foods = np.stack([fx, fy], 1)[fx >= 0]
    viruses = np.stack([seg(L.virus_x, c.num_viruses), seg(L.virus_y, c.num_viruses)], 1)

    def player(k):
        s = slice(k * 16, k * 16 + 16)
        rows = np.stack([seg(L.cell_x, L.cells_total)[s], seg(L.cell_y, L.cells_total)[s], seg(L.cell_ix, L.cells_total)[s],
                         seg(L.cell_iy, L.cells_total)[s], seg(L.cell_m, L.cells_total)[s]], 1)
        return rows[rows[:, 4] > 0]

    others = [p for p in (player(k) for k in range(L.num_players) if k != agent) if p.shape[0] > 0]
    return FullObservation(pellets, viruses, foods, player(agent), others)
