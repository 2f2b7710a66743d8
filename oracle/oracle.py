"""ctypes binding of the CPU oracle (oracle/agar_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; nothing under agarai_b200/ does.  See the header of agar_oracle.c for what
is pinned against the reference and what is not (the env step is "parity unpinned").
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


class AgarConfig(C.Structure):
    """Mirror of `AgarConfig` in include/agar_b200.h (kept field-for-field identical)."""
    _fields_ = [
        ("num_arenas", C.c_int32), ("num_agents", C.c_int32), ("num_bots", C.c_int32),
        ("num_pellets", C.c_int32), ("num_viruses", C.c_int32), ("max_foods", C.c_int32),
        ("ticks_per_step", C.c_int32), ("grid_size", C.c_int32), ("obs_flags", C.c_int32),
        ("pellet_regen", C.c_int32), ("auto_reset", C.c_int32), ("event_capacity", C.c_int32),
        ("bots_split", C.c_int32), ("reserved0", C.c_int32), ("arena_id_base", C.c_int64),
        ("arena_size", C.c_float), ("view_size", C.c_float), ("start_mass", C.c_float),
        ("pellet_mass", C.c_float), ("food_mass", C.c_float), ("virus_mass", C.c_float),
        ("split_min_mass", C.c_float), ("feed_min_mass", C.c_float), ("virus_pop_mass", C.c_float),
        ("eat_ratio", C.c_float), ("r2_per_mass", C.c_float), ("speed_k", C.c_float),
        ("target_reach", C.c_float), ("split_speed", C.c_float), ("impulse_decay", C.c_float),
        ("feed_speed", C.c_float), ("food_decay", C.c_float), ("food_spawn_gap", C.c_float),
        ("pop_speed", C.c_float), ("merge_delay", C.c_float), ("pop_pieces", C.c_int32),
        ("bot_roam_period", C.c_int32), ("bot_sense_radius", C.c_float), ("mass_decay", C.c_float),
    ]


class AgarLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "words", "pellet_x", "pellet_y", "virus_x", "virus_y", "dyn_begin", "cell_x", "cell_y",
        "cell_ix", "cell_iy", "cell_m", "cell_t", "food_x", "food_y", "food_vx", "food_vy",
        "agent_done", "tick", "episode_steps", "dyn_end", "num_players", "cells_total",
        "P_pad", "V_pad", "F_pad", "A_pad")]


def build(force=False):
    """Compile oracle/agar_oracle.c with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "agar_oracle.c")
    deps = [src, os.path.join(_HERE, "..", "include", "agar_b200.h"),
            os.path.join(_HERE, "..", "include", "agar_spec.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        # never rebuild on the GPU box if a prebuilt .so travelled with the snapshot and gcc is missing
        try:
            build()
        except Exception:
            if not os.path.exists(_SO):
                raise
        L = C.CDLL(_SO)
        f32p, i32p, u8p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(AgarConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_words.argtypes = [C.c_void_p]
        L.oracle_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_reset.argtypes = [C.c_void_p, u8p, f32p, C.c_int]
        L.oracle_step.argtypes = [C.c_void_p, f32p, i32p, f32p, f32p, u8p, C.c_int]
        L.oracle_observe.argtypes = [C.c_void_p, f32p, C.c_int]
        L.oracle_observe_records.argtypes = [C.c_void_p, f32p, C.c_int, f32p]
        L.oracle_get_state.argtypes = [C.c_void_p, f32p]
        L.oracle_set_state.argtypes = [C.c_void_p, f32p]
        L.oracle_events.argtypes = [C.c_void_p, i32p, i32p]
        L.oracle_config_default.argtypes = [C.POINTER(AgarConfig)]
        L.oracle_layout.argtypes = [C.POINTER(AgarConfig), C.POINTER(AgarLayout)]
        L.oracle_gae.argtypes = [f32p, f32p, u8p, f32p, C.c_double, C.c_double, f32p, f32p, C.c_int, C.c_int]
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_centroid.argtypes = [f32p, f32p, f32p, f32p]
        L.oracle_centroid.restype = C.c_int
        L.oracle_grid_extract.restype = C.c_int
        L.oracle_grid_extract.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                          f32p, f32p, C.c_int, f32p, f32p, C.c_int, f32p, f32p, C.c_int,
                                          f32p, f32p, f32p, f32p]
        L.oracle_ram_feature_size.restype = C.c_int
        L.oracle_ram_feature_size.argtypes = [C.c_int] * 5
        L.oracle_ram_extract.restype = C.c_int
        L.oracle_ram_extract.argtypes = [C.c_int] * 7 + [f32p, f32p, C.c_int, f32p, f32p, C.c_int, f32p, f32p, C.c_int,
                                                       f32p, f32p, f32p, f32p, f32p, f32p]
        L.oracle_observe_ram.argtypes = [C.c_void_p] + [C.c_int] * 5 + [f32p]
        L.oracle_screen_gray.argtypes = [u8p, C.c_longlong, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def default_config(**kw):
    cfg = AgarConfig()
    lib().oracle_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown AgarConfig field {k!r}")
        setattr(cfg, k, v)
    return cfg


def layout(cfg):
    L = AgarLayout()
    lib().oracle_layout(C.byref(cfg), C.byref(L))
    return L


def obs_channels(cfg):
    return bin(cfg.obs_flags & 31).count("1")


class OracleEnv:
    """Sequential CPU definition of the engine's rules; numpy in, numpy out."""

    def __init__(self, cfg, seed=0, nthreads=1):
        self.cfg = cfg
        self._h = lib().oracle_create(C.byref(cfg))
        if not self._h:
            raise ValueError("oracle_create rejected the config")
        self.L = layout(cfg)
        self.N, self.A = cfg.num_arenas, cfg.num_agents
        self.C, self.G = obs_channels(cfg), cfg.grid_size
        self.nthreads = nthreads
        lib().oracle_seed(self._h, seed)

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    __del__ = close

    def seed(self, seed):
        lib().oracle_seed(self._h, seed)

    def _obs_buf(self):
        return np.empty((self.N, self.A, self.C, self.G, self.G), np.float32)

    def reset(self, mask=None, want_obs=True):
        obs = self._obs_buf() if want_obs else None
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().oracle_reset(self._h, _p(m, C.c_uint8), _p(obs, C.c_float), self.nthreads)
        return obs

    def step(self, target_xy, act, want_obs=True, out=None):
        t = np.ascontiguousarray(target_xy, np.float32).reshape(self.N, self.A, 2)
        a = np.ascontiguousarray(act, np.int32).reshape(self.N, self.A)
        obs = (out if out is not None else self._obs_buf()) if want_obs else None
        rew = np.empty((self.N, self.A), np.float32)
        done = np.empty((self.N, self.A), np.uint8)
        lib().oracle_step(self._h, _p(t, C.c_float), _p(a, C.c_int32), _p(obs, C.c_float),
                          _p(rew, C.c_float), _p(done, C.c_uint8), self.nthreads)
        return obs, rew, done

    def observe(self):
        obs = self._obs_buf()
        lib().oracle_observe(self._h, _p(obs, C.c_float), self.nthreads)
        return obs

    def observe_records(self, state):
        s = np.ascontiguousarray(state, np.float32)
        obs = np.empty((s.shape[0], self.A, self.C, self.G, self.G), np.float32)
        lib().oracle_observe_records(self._h, _p(s, C.c_float), s.shape[0], _p(obs, C.c_float))
        return obs

    def get_state(self):
        s = np.empty((self.N, self.L.words), np.float32)
        lib().oracle_get_state(self._h, _p(s, C.c_float))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32)
        assert s.shape == (self.N, self.L.words)
        lib().oracle_set_state(self._h, _p(s, C.c_float))

    def observe_ram(self, n_pellet=50, n_virus=5, n_food=10, n_other=5, n_cell=15):
        size = lib().oracle_ram_feature_size(n_pellet, n_virus, n_food, n_other, n_cell)
        out = np.empty((self.N, self.A, size), np.float32)
        lib().oracle_observe_ram(self._h, n_pellet, n_virus, n_food, n_other, n_cell, _p(out, C.c_float))
        return out

    def events(self):
        cap = self.cfg.event_capacity
        ev = np.zeros((self.N, max(cap, 1), 4), np.int32)
        cnt = np.zeros((self.N,), np.int32)
        if cap > 0:
            lib().oracle_events(self._h, _p(ev, C.c_int32), _p(cnt, C.c_int32))
        return ev, cnt


def gae(reward, value, gamma, lam, done=None, end_value=None, want_adv=True, want_ret=True):
    """oracle_gae over time-major [T,N] float32 buffers (rl/test.py:12-30 restated in C)."""
    r = np.ascontiguousarray(reward, np.float32)
    v = np.ascontiguousarray(value, np.float32)
    T, N = r.shape
    assert v.shape == (T + 1, N)
    d = None if done is None else np.ascontiguousarray(done, np.uint8)
    ev = None if end_value is None else np.ascontiguousarray(end_value, np.float32)
    adv = np.empty_like(r) if want_adv else None
    ret = np.empty_like(r) if want_ret else None
    lib().oracle_gae(_p(r, C.c_float), _p(v, C.c_float), _p(d, C.c_uint8), _p(ev, C.c_float),
                     float(gamma), float(lam), _p(adv, C.c_float), _p(ret, C.c_float), T, N)
    return adv, ret


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def centroid(x, y, m):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    m = np.ascontiguousarray(m, np.float32)
    assert x.shape == (16,)
    out = np.zeros(2, np.float32)
    ok = lib().oracle_centroid(_p(x, C.c_float), _p(y, C.c_float), _p(m, C.c_float), _p(out, C.c_float))
    return out if ok else None


def grid_extract(G, view, arena, flags, agent, K, pellets, viruses, foods, cell_x, cell_y, cell_m):
    """oracle_grid_extract on raw entity arrays; pellets/viruses/foods are [n,2] float32."""
    def split(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 2)
        return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), a.shape[0]
    px, py, P = split(pellets); vx, vy, V = split(viruses); fx, fy, F = split(foods)
    cx = np.ascontiguousarray(cell_x, np.float32); cy = np.ascontiguousarray(cell_y, np.float32)
    cm = np.ascontiguousarray(cell_m, np.float32)
    assert cx.shape == (K * 16,)
    Cn = bin(flags & 31).count("1")
    out = np.empty((Cn, G, G), np.float32)
    ok = lib().oracle_grid_extract(G, view, arena, flags, agent, K, _p(px, C.c_float), _p(py, C.c_float), P,
                                   _p(vx, C.c_float), _p(vy, C.c_float), V, _p(fx, C.c_float), _p(fy, C.c_float), F,
                                   _p(cx, C.c_float), _p(cy, C.c_float), _p(cm, C.c_float), _p(out, C.c_float))
    return out, bool(ok)


def ram_extract(counts, agent, K, pellets, viruses, foods, cell_x, cell_y, cell_ix, cell_iy, cell_m):
    """oracle_ram_extract on raw entity arrays (FeatureExtractor.extract, features/extractors.py:99-149);
    counts = (n_pellet, n_virus, n_food, n_other, n_cell); pellets/viruses/foods are [n,2] float32."""
    def split(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 2)
        return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), a.shape[0]
    px, py, P = split(pellets); vx, vy, V = split(viruses); fx, fy, F = split(foods)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (cell_x, cell_y, cell_ix, cell_iy, cell_m)]
    assert all(a.shape == (K * 16,) for a in arrs)
    out = np.empty(lib().oracle_ram_feature_size(*counts), np.float32)
    ok = lib().oracle_ram_extract(*counts, agent, K, _p(px, C.c_float), _p(py, C.c_float), P, _p(vx, C.c_float),
                                  _p(vy, C.c_float), V, _p(fx, C.c_float), _p(fy, C.c_float), F,
                                  *[_p(a, C.c_float) for a in arrs], _p(out, C.c_float))
    return out, bool(ok)


def screen_gray(rgb):
    """oracle_screen_gray: uint8[..., 3] -> float64[...] (ScreenFeatureExtractor.extract, :152-164)."""
    a = np.ascontiguousarray(rgb, np.uint8)
    out = np.empty(a.shape[:-1], np.float64)
    lib().oracle_screen_gray(_p(a, C.c_uint8), out.size, _p(out, C.c_double))
    return out
