"""CPU coverage of the host side of agar_step_host's compact copy (agarai_b200/csrc/agar_host_expand.h: plain C++, the
functions the library's worker threads run): the grid rebuilt from an agent's occupied-bin list + centroid, and the
in-place update of the persistent host buffer from the previous step's list to this step's, must both equal the oracle's
dense grid bit for bit.  The lists are taken from the oracle's own images the way obs_compact_kernel takes them from the
device image (every entry that is neither +0.0 nor -1.0); the harness (tests/host_expand_harness.cpp) is built with g++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32
ENTRY = np.dtype([("idx", "<u4"), ("val", "<f4")])


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("he") / "host_expand.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", out,
                           os.path.join(ROOT, "tests", "host_expand_harness.cpp")])
    L = C.CDLL(out)
    fp, vp = C.POINTER(C.c_float), C.c_void_p
    L.he_expand.argtypes = [C.c_int] * 4 + [C.c_float, fp, fp, vp, C.c_int, fp]
    L.he_delta.argtypes = [C.c_int] * 4 + [C.c_float, fp, fp, vp, C.c_int, fp, vp, C.c_int, fp]
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _lists_and_meta(oracle, env, obs):
    """Per agent: (entries, meta) as the device side produces them."""
    N, A = env.N, env.A
    st = env.get_state()
    Lc = env.L
    out = []
    for n in range(N):
        for a in range(A):
            img = obs[n, a].reshape(-1)
            bits = img.view(np.uint32)
            idx = np.flatnonzero((bits != 0) & (bits != 0xBF800000)).astype(np.uint32)
            ent = np.zeros(len(idx), ENTRY)
            ent["idx"], ent["val"] = idx, img[idx]
            sl = slice(a * 16, a * 16 + 16)
            x, y, m = (st[n, o:o + Lc.cells_total][sl].copy() for o in (Lc.cell_x, Lc.cell_y, Lc.cell_m))
            c = oracle.centroid(x, y, m)
            meta = np.array([0, 0, 0, 0] if c is None else [c[0], c[1], 1, 0], f32)
            out.append((ent, meta))
    return out


WORLDS = {
    "walls_five_channels": dict(num_arenas=6, num_agents=3, num_bots=4, num_viruses=5, num_pellets=300, arena_size=150.0, view_size=100.0,
                                grid_size=32, obs_flags=31, start_mass=40.0, virus_pop_mass=60.0, virus_mass=30.0, auto_reset=1,
                                speed_k=8.0, max_foods=16),
    "single_player_no_sentinel_in_others": dict(num_arenas=5, num_agents=1, num_bots=0, num_pellets=200, arena_size=90.0,
                                                view_size=128.0, grid_size=64, obs_flags=7, speed_k=8.0),
    # grid sides that are not powers of two (the index -> (channel, row, column) split uses multiply-shift) and the largest one
    "grid_20": dict(num_arenas=4, num_agents=2, num_bots=3, num_viruses=3, num_pellets=250, arena_size=110.0, view_size=90.0,
                    grid_size=20, obs_flags=15, start_mass=30.0, speed_k=8.0, auto_reset=1),
    "grid_128_dqn_like": dict(num_arenas=3, num_agents=1, num_bots=0, num_pellets=30, arena_size=45.0, view_size=30.0, grid_size=128,
                              obs_flags=15, speed_k=6.0),
    "cfg2": dict(num_arenas=4, num_agents=1, num_bots=25, num_viruses=25, num_pellets=1000, arena_size=500.0, grid_size=64,
                 obs_flags=7, auto_reset=1, max_foods=64),
}


@pytest.mark.parametrize("world", sorted(WORLDS))
def test_rebuild_and_in_place_update_equal_the_dense_grid(harness, oracle, world):
    kw = WORLDS[world]
    cfg = oracle.default_config(**kw)
    env = oracle.OracleEnv(cfg, seed=5, nthreads=4)
    N, A, G = env.N, env.A, cfg.grid_size
    Cn, K = env.C, cfg.num_agents + cfg.num_bots
    oob = np.array([f32(float(i - G // 2) * (float(cfg.view_size) / float(G))) for i in range(G)], f32)
    geo = (G, Cn, K, cfg.obs_flags, C.c_float(cfg.arena_size), _fp(oob))
    rng = np.random.default_rng(9)
    obs = env.reset()
    prev = _lists_and_meta(oracle, env, obs)
    buf = obs.copy()                                  # the persistent buffer: holds the previous step's grids
    flips = 0
    for t in range(80):
        xy = rng.uniform(-1, 1, (N, A, 2)).astype(f32)
        xy[: N // 2] = np.sign(xy[: N // 2, :, :1]) * np.array([1.0, 0.4], f32)   # half of the arenas run into the side walls
        act = rng.integers(0, 3, (N, A)).astype(np.int32)
        if t == 30:                                   # a masked reset: agents jump, the buffer still holds step 29's grids
            obs = env.reset(mask=(rng.random(N) < 0.5).astype(np.uint8))
        else:
            obs, _, _ = env.step(xy, act)
        cur = _lists_and_meta(oracle, env, obs)
        for g, ((pe, pm), (ne, nm)) in enumerate(zip(prev, cur)):
            n, a = divmod(g, A)
            want = obs[n, a].view(np.int32)
            full = np.full((Cn, G, G), 7.0, f32)      # stale content must be overwritten everywhere
            harness.he_expand(*geo, _fp(full), ne.ctypes.data, len(ne), _fp(nm))
            assert np.array_equal(full.view(np.int32), want), f"{world}: rebuild differs, step {t} agent {g}"
            harness.he_delta(*geo, _fp(buf[n, a]), pe.ctypes.data, len(pe), _fp(pm), ne.ctypes.data, len(ne), _fp(nm))
            assert np.array_equal(buf[n, a].view(np.int32), want), f"{world}: in-place update differs, step {t} agent {g}"
            flips += int(pm[2] != nm[2])
        prev = cur
    if world == "walls_five_channels":
        assert flips > 0, "no agent died or respawned: the full-rewrite branch of the update was not exercised"
    env.close()
