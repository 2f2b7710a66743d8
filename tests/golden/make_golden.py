"""Generates the golden fixtures of tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

* grid_*.npz — inputs and outputs of `GridFeatureExtractor.extract` and `position`
  (features/extractors.py:39-96, :207-212), imported from /root/reference with the one missing
  type import (`gym_agario.envs.FullEnv.FullObservation`, an external package) stubbed by a
  namedtuple of the fields the extractor reads.  Inputs are float32 arrays, so under numpy >= 2
  (NEP 50 weak scalars) the reference evaluates in float32 and accumulates into float64.
* gae_golden.npz — the numpy statements the reference's own tests hold rl.gae / rl.make_returns
  to (`normal_gae`, `make_returns`, rl/test.py:12-30) evaluated on seeded float32 inputs, plus
  a2c.losses.make_returns (a2c/losses.py:66-78).  rl/test.py itself cannot be imported (it
  imports jax and parameterized at module top; neither is installed), so those two 10-line
  functions are extracted from its source text with `ast` and executed unchanged.
"""
import ast
import collections
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def import_extractors():
    m = types.ModuleType("gym_agario"); e = types.ModuleType("gym_agario.envs")
    f = types.ModuleType("gym_agario.envs.FullEnv")
    f.FullObservation = collections.namedtuple("FullObservation", ["pellets", "viruses", "foods", "agent", "others"])
    sys.modules.update({"gym_agario": m, "gym_agario.envs": e, "gym_agario.envs.FullEnv": f})
    sys.path.insert(0, REF)
    from features import extractors
    return extractors, f.FullObservation


def functions_from_source(path, names, extra_globals):
    """exec the named top-level functions of a reference file without importing the module."""
    tree = ast.parse(open(path).read())
    ns = dict(extra_globals)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def cells5(rng, c, arena, lo, hi, centre=None, spread=None):
    p = np.zeros((c, 5), f32)
    if centre is None:
        p[:, 0:2] = rng.uniform(0, arena, (c, 2))
    else:
        p[:, 0:2] = np.clip(centre + rng.uniform(-spread, spread, (c, 2)), 0, arena)
    p[:, 2:4] = rng.normal(0, 1, (c, 2))
    p[:, 4] = rng.uniform(lo, hi, c)
    return p


def grid_case(ex_mod, Obs, rng, G, view, arena, flags, P, V, F, n_others, agent_cells, where):
    """One extract() call; returns dict of arrays (engine slot layout + reference output)."""
    cells, others_on, viruses_on, foods_on = bool(flags & 2), bool(flags & 4), bool(flags & 8), bool(flags & 16)
    assert flags & 1
    if where == "corner":
        centre = np.array([rng.uniform(0, view / 3), rng.uniform(0, view / 3)])
    elif where == "edge":
        centre = np.array([arena - rng.uniform(0, view / 3), rng.uniform(0, arena)])
    else:
        centre = rng.uniform(view, arena - view, 2) if arena > 2 * view else rng.uniform(0, arena, 2)
    agent = cells5(rng, agent_cells, arena, 10, 400, centre, view / 6)
    pellets = rng.uniform(0, arena, (P, 2)).astype(f32)
    if P:  # a dense clump inside the view so that bins hold several pellets
        k = P // 4
        pellets[:k] = np.clip(centre + rng.uniform(-view / 2, view / 2, (k, 2)), 0, np.nextafter(f32(arena), f32(0))).astype(f32)
    viruses = rng.uniform(0, arena, (V, 2)).astype(f32)
    if V:
        viruses[: V // 2] = np.clip(centre + rng.uniform(-view / 2, view / 2, (V // 2, 2)), 0, arena).astype(f32)
    foods = np.clip(centre + rng.uniform(-view, view, (F, 2)), 0, arena).astype(f32)
    others = []
    for k in range(n_others):
        c = int(rng.integers(0, 17))  # 0 cells = a dead player, still listed
        near = rng.random() < 0.7
        others.append(cells5(rng, c, arena, 10, 2000, centre if near else None, view / 2))
    ex = ex_mod.GridFeatureExtractor(float(f32(view)), G, float(f32(arena)) if arena != int(arena) else int(arena),
                                     cells=cells, others=others_on, viruses=viruses_on, food=foods_on)
    out = ex.extract(Obs(pellets, viruses, foods, agent, others))
    loc = ex_mod.position(agent)
    assert out.shape == (bin(flags).count("1"), G, G) and loc.dtype == f32
    # engine layout: K = 1 + n_others players x 16 slots, agent = player 0, alive cells first
    K = 1 + n_others
    cx = np.zeros(K * 16, f32); cy = np.zeros(K * 16, f32); cm = np.zeros(K * 16, f32)
    for k, pl in enumerate([agent] + others):
        n = pl.shape[0]
        # scatter the cells over the 16 slots in order, leaving random holes (empty slots)
        slots = np.sort(rng.choice(16, n, replace=False))
        cx[k * 16 + slots] = pl[:, 0]; cy[k * 16 + slots] = pl[:, 1]; cm[k * 16 + slots] = pl[:, 4]
    return dict(G=G, view=f32(view), arena=f32(arena), flags=flags, K=K, pellets=pellets, viruses=viruses,
                foods=foods, cell_x=cx, cell_y=cy, cell_m=cm, expect=out, loc=loc)


def main():
    ex_mod, Obs = import_extractors()
    rng = np.random.default_rng(20260119)
    cases = []
    # (G, view, arena, flags, P, V, F, n_others, agent_cells, where)
    specs = [
        (64, 128, 1000, 7, 1000, 0, 0, 0, 1, "mid"),      # BASELINE cfg1: 3x64x64, no other player
        (64, 128, 1000, 7, 1000, 0, 0, 0, 3, "corner"),
        (64, 128, 1000, 7, 1000, 0, 0, 15, 5, "mid"),     # cfg4: 16 agents
        (64, 128, 1000, 7, 1000, 0, 0, 15, 16, "edge"),
        (64, 128, 500, 15, 1000, 25, 0, 25, 2, "mid"),    # cfg2: bots + viruses
        (64, 128, 500, 15, 1000, 25, 0, 25, 8, "corner"),
        (64, 128, 500, 31, 1000, 25, 40, 25, 9, "edge"),  # all five channels
        (32, 100, 500, 15, 1000, 25, 0, 7, 4, "mid"),     # a2c hyperparameters: G=32, non-dyadic box
        (32, 100, 500, 15, 300, 25, 0, 7, 12, "corner"),
        (128, 30, 45, 7, 30, 0, 0, 2, 2, "mid"),          # dqn/hyperparameters.py:23-28,97-98
        (128, 30, 45, 31, 30, 5, 10, 2, 1, "corner"),
        (64, 90, 777.5, 31, 500, 10, 20, 3, 7, "edge"),   # awkward sizes
        (8, 64, 200, 3, 50, 0, 0, 0, 1, "mid"),           # tiny grid, pellets+cells only
        (64, 128, 1000, 1, 1000, 0, 0, 0, 2, "corner"),   # pellets only
    ]
    for rep in range(3):
        for sp in specs:
            cases.append(grid_case(ex_mod, Obs, rng, *sp))
    keys = cases[0].keys()
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), n=len(cases), numpy_version=np.__version__,
                        **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in keys})
    print("grid cases:", len(cases))

    # position() alone, every cell count 1..16 many times (pins the pairwise summation order)
    xs, ys, ms, locs = [], [], [], []
    for trial in range(4000):
        c = 1 + trial % 16
        x = np.zeros(16, f32); y = np.zeros(16, f32); m = np.zeros(16, f32)
        slots = np.sort(rng.choice(16, c, replace=False))
        x[slots] = rng.uniform(0, 1000, c); y[slots] = rng.uniform(0, 1000, c); m[slots] = rng.uniform(1, 5000, c)
        pl = np.zeros((c, 5), f32); pl[:, 0] = x[slots]; pl[:, 1] = y[slots]; pl[:, 4] = m[slots]
        xs.append(x); ys.append(y); ms.append(m); locs.append(ex_mod.position(pl))
    np.savez_compressed(os.path.join(HERE, "position_golden.npz"), x=np.array(xs), y=np.array(ys), m=np.array(ms),
                        loc=np.array(locs), numpy_version=np.__version__)

    # GAE / returns
    normal_gae, make_returns = functions_from_source(os.path.join(REF, "rl/test.py"), ["normal_gae", "make_returns"],
                                                     {"onp": np})
    (a2c_make_returns,) = functions_from_source(os.path.join(REF, "a2c/losses.py"), ["make_returns"], {"np": np})
    T, N = 100, 64
    r = rng.normal(0, 1, (T, N)).astype(f32)
    v = rng.normal(0, 1, (T + 1, N)).astype(f32)
    gammas = rng.uniform(0, 1, N); lams = rng.uniform(0, 1, N)
    gammas[:4] = [0.75, 1.0, 0.0, 0.5]; lams[:4] = [0.1, 1.0, 0.0, 0.0]
    adv = np.stack([normal_gae(r[:, n].copy(), v[:, n].copy(), float(gammas[n]), float(lams[n])) for n in range(N)], 1)
    ret0 = np.stack([make_returns(r[:, n].copy(), float(gammas[n]), end_value=0.0) for n in range(N)], 1)
    ret2 = np.stack([make_returns(r[:, n].copy(), float(gammas[n]), end_value=2.0) for n in range(N)], 1)
    reta2c = np.stack([a2c_make_returns(r[:, n].copy(), float(gammas[n])) for n in range(N)], 1)
    assert adv.dtype == f32 and ret0.dtype == f32
    # test/returns_test.py:35-46 vector (float64 rewards, gamma .75), via the a2c implementation
    np.random.seed(10)
    rt = np.random.randn(30)
    rt_ret = a2c_make_returns(rt, 0.75)
    np.savez_compressed(os.path.join(HERE, "gae_golden.npz"), r=r, v=v, gamma=gammas, lam=lams, adv=adv, ret0=ret0,
                        ret2=ret2, reta2c=reta2c, rt=rt, rt_ret=rt_ret, numpy_version=np.__version__)
    print("gae golden written")


if __name__ == "__main__":
    main()
