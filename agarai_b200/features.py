"""Device-side mirrors of the reference's feature extractors (features/extractors.py), same class names,
constructor arguments and `shape` / `size` attributes.  The reference's extractors take a `FullObservation`
of ONE agent on the host; these take the batched engine (the state already lives in HBM) or a device tensor
and return the whole batch from one kernel launch.  No fallback: without the CUDA library they raise.

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


class GridFeatureExtractor:
    """features/extractors.py:11-34: the C x G x G grid.  View size, grid size and the channel toggles are
    part of the engine's config (they fix the kernel's shared-memory plan), so this object only checks that
    the env was built with the same ones and exposes `shape`."""

    def __init__(self, view_size, grid_size, arena_size, cells=True, others=True, viruses=True, food=False, pellets=True):
        self.view_size, self.grid_size, self.arena_size = float(view_size), int(grid_size), float(arena_size)
        self.flags = (1 if pellets else 0) | (2 if cells else 0) | (4 if others else 0) | (8 if viruses else 0) | (16 if food else 0)
        self.depth = bin(self.flags).count("1")
        self.shape = (self.depth, self.grid_size, self.grid_size)  # features/extractors.py:31

    def __call__(self, env):
        return self.extract(env)

    def extract(self, env):
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

    def __call__(self, env):
        return self.extract(env)

    def extract(self, env, out=None):
        """float32 [num_envs, num_agents, size] for the env's current state (zeros for an agent without
        cells, where the reference returns None)."""
        L = _lib.lib()
        if out is None:
            out = torch.empty((env.num_envs, env.num_agents, self.size), dtype=torch.float32, device=env.device)
        _lib.check(L.agar_observe_ram(env._h, self.num_pellet, self.num_virus, self.num_food, self.num_other,
                                      self.num_cell, _ptr(out), _stream(env.device)))
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


import collections

FullObservation = collections.namedtuple("FullObservation", ["pellets", "viruses", "foods", "agent", "others"])


def full_observation(env, arena: int = 0, agent: int = 0) -> "FullObservation":
    """The raw entity lists of one agent's view, in the shape of the external package's `FullObservation`
    that the reference's extractors read (features/extractors.py:40-64,108,140): present pellets / foods
    [n,2], viruses [V,2], the agent's cells and every other player that has cells as [c,5] rows of
    (x, y, impulse_x, impulse_y, mass).  Host-side view of the device state (a debugging and interop aid for
    `agario-full-v0` style consumers; the accelerated extractors above never go through it)."""
    L, c = env.layout, env.cfg
    rec = env.get_state()[arena].cpu().numpy()
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
