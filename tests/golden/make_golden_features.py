"""Golden fixtures for the RAM-style and screen feature extractors, made by running the REFERENCE's own
code (features/extractors.py:99-164, helpers :171-212) imported from /root/reference.

    python tests/golden/make_golden_features.py        (build container only: needs /root/reference)

Writes tests/golden/ram_golden.npz and screen_golden.npz.  Inputs are float32 entity arrays in the engine's
slot layout; the reference is fed what a FullObservation of that state holds: the present pellets / foods in
index order, the players that have cells in player order, cell rows (x, y, impulse_x, impulse_y, mass).
numpy's default argsort leaves the order of equal keys unspecified, so a case in which two candidates of one
sort have exactly equal float32 keys and different rows is re-drawn (the spec here is the stable order).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_extractors  # noqa: E402

f32 = np.float32


def has_ambiguous_tie(keys, rows):
    order = np.argsort(keys, kind="stable")
    k = np.asarray(keys)[order]
    for i in np.nonzero(k[1:] == k[:-1])[0]:
        if not np.array_equal(rows[order[i]], rows[order[i + 1]]):
            return True
    return False


def ram_case(ex_mod, Obs, rng, counts, arena, P, V, F, K, max_cells, absent_frac):
    n_pellet, n_virus, n_food, n_other, n_cell = counts
    while True:
        px = rng.uniform(0, arena, P).astype(f32); py = rng.uniform(0, arena, P).astype(f32)
        if P:
            gone = rng.random(P) < absent_frac
            px[gone] = -1.0; py[gone] = -1.0
        vx = rng.uniform(0, arena, V).astype(f32); vy = rng.uniform(0, arena, V).astype(f32)
        fx = rng.uniform(0, arena, F).astype(f32); fy = rng.uniform(0, arena, F).astype(f32)
        if F:
            free = rng.random(F) < 0.5
            fx[free] = -1.0; fy[free] = 0.0
        cx = np.zeros(K * 16, f32); cy = np.zeros(K * 16, f32); cix = np.zeros(K * 16, f32)
        ciy = np.zeros(K * 16, f32); cm = np.zeros(K * 16, f32)
        players = []
        for k in range(K):
            c = int(rng.integers(1 if k == 0 else 0, max_cells + 1))
            slots = np.sort(rng.choice(16, c, replace=False))
            centre = rng.uniform(0, arena, 2)
            g = k * 16 + slots
            cx[g] = np.clip(centre[0] + rng.uniform(-20, 20, c), 0, arena); cy[g] = np.clip(centre[1] + rng.uniform(-20, 20, c), 0, arena)
            cix[g] = rng.normal(0, 2, c); ciy[g] = rng.normal(0, 2, c)
            cm[g] = rng.uniform(10, 500, c)
            if c >= 2 and rng.random() < 0.5:  # twins: equal sort keys everywhere, identical rows (any tie order agrees)
                for arr in (cx, cy, cix, ciy, cm):
                    arr[g[1]] = arr[g[0]]
            players.append(np.stack([cx[g], cy[g], cix[g], ciy[g], cm[g]], 1).astype(f32))
        agent = players[0]
        others = [pl for pl in players[1:] if pl.shape[0] > 0]
        pellets = np.stack([px, py], 1)[px >= 0]
        viruses = np.stack([vx, vy], 1)
        foods = np.stack([fx, fy], 1)[fx >= 0]
        loc = ex_mod.position(agent)
        # ambiguity checks on every sort the reference performs
        amb = False
        for ent in (pellets, viruses, foods):
            if ent.shape[0]:
                amb |= has_ambiguous_tie(np.linalg.norm(ent - loc, axis=1), ent)
        for pl in [agent] + others:
            amb |= has_ambiguous_tie(pl[:, -1], pl)
            amb |= has_ambiguous_tie(np.linalg.norm(pl[:, :2] - loc, axis=1), pl)
        if others:
            d = [np.linalg.norm(loc - ex_mod.position(pl)) for pl in others]
            amb |= len(set(map(float, d))) != len(d)
        if amb:
            continue
        ex = ex_mod.FeatureExtractor(num_pellet=n_pellet, num_virus=n_virus, num_food=n_food, num_other=n_other, num_cell=n_cell)
        out = ex.extract(Obs(pellets, viruses, foods, agent, others))
        assert out.shape == (ex.size,)
        return dict(counts=np.array(counts), K=K, px=px, py=py, vx=vx, vy=vy, fx=fx, fy=fy, cx=cx, cy=cy, cix=cix, ciy=ciy,
                    cm=cm, expect=out)


def main():
    ex_mod, Obs = import_extractors()
    rng = np.random.default_rng(20261020)
    cases = []
    specs = [  # (counts, arena, P, V, F, K, max_cells, absent_frac)
        ((50, 5, 10, 5, 15), 1000, 1000, 25, 64, 26, 4, 0.0),   # reference defaults on the cfg2 world
        ((50, 5, 10, 5, 15), 500, 1000, 25, 64, 26, 16, 0.1),
        ((50, 5, 10, 5, 15), 1000, 1000, 0, 64, 1, 3, 0.0),     # single agent, no viruses: filler players
        ((50, 5, 10, 5, 15), 1000, 1000, 0, 64, 16, 16, 0.0),   # cfg4: 16 agents
        ((50, 5, 10, 5, 15), 45, 30, 5, 8, 3, 2, 0.3),          # dqn world: fewer pellets than slots
        ((8, 2, 4, 3, 16), 200, 100, 3, 16, 6, 16, 0.5),
        ((1, 1, 1, 1, 1), 100, 10, 1, 4, 2, 5, 0.0),
        ((64, 8, 8, 10, 4), 300, 500, 12, 32, 8, 9, 0.2),       # more player slots than players
    ]
    for rep in range(4):
        for sp in specs:
            cases.append(ram_case(ex_mod, Obs, rng, *sp))
    keys = cases[0].keys()
    np.savez_compressed(os.path.join(HERE, "ram_golden.npz"), n=len(cases), numpy_version=np.__version__,
                        **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cases) for k in keys})
    print("ram cases:", len(cases))

    sx = ex_mod.ScreenFeatureExtractor(2, 40)
    frames = rng.integers(0, 256, (2, 2, 40, 40, 3), dtype=np.uint8)
    frames[0, 0] = 255; frames[0, 1] = 0
    odd = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)   # size not a multiple of 16 pixels
    np.savez_compressed(os.path.join(HERE, "screen_golden.npz"), frames=frames, gray=sx.extract(frames), odd=odd,
                        odd_gray=sx.extract(odd), numpy_version=np.__version__)
    print("screen golden written", sx.extract(frames).dtype)


if __name__ == "__main__":
    main()
