"""Generates tests/golden/astar_*.npz from the LIVE reference route planner (build container only).

    python tests/golden/make_golden_astar.py

Maps = the reference's own `is_road_map / road_type_map / allowed_dirs_map` of a committed layout fixture, plus seeded
sparse occupancy / stop maps and a density map; queries = seeded (start, goal, flags) tuples covering every flag
combination `vehicle_base.py` uses (:231-398).  Each query's path comes from the unmodified
`Simulation.utilities.pathfinding.astar_numba.astar_numba` (Numba) and is stored as first cell + 2-bit steps.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {"default12345": ("layout_default12345.npz", 12345, 360), "s14_150x110_carve": ("layout_s14_150x110_carve.npz", 14, 240)}
FLAG_SETS = [   # (respect_awareness, soft_obstacles, ignore_flow, maximum_steps): the call sites of vehicle_base.py
    (False, False, False, 0x7FFFFFFF),   # :280 plain route
    (False, True, False, 0x7FFFFFFF),    # :295 route through soft obstacles
    (False, False, True, 6),             # :231 contraflow overtake
    (False, True, True, 20),             # :388 stuck detour
    (True, False, False, 0x7FFFFFFF),    # VEHICLE_RESPECT_AWARENESS on
    (True, True, True, 40),
]


def make_inputs(fixture, seed, n_queries):
    d = np.load(os.path.join(HERE, fixture), allow_pickle=True)
    road, rtype, adirs = d["is_road_map"], d["road_type_map"], d["allowed_dirs_map"]
    H, W = road.shape
    rng = np.random.default_rng(seed)
    cells = np.flatnonzero(road.reshape(-1) == 1)
    occ = np.zeros(H * W, np.uint8)
    occ[rng.choice(cells, size=len(cells) // 12, replace=False)] = 1
    stop = np.zeros(H * W, np.uint8)
    stop[rng.choice(cells, size=len(cells) // 40, replace=False)] = 1
    dens = np.round(rng.random((H, W)) * 0.6, 3)
    q = np.zeros((n_queries, 8), np.int32)   # sx, sy, gx, gy, respect, soft, ignore, max_steps
    for i in range(n_queries):
        a = rng.choice(cells)
        fl = FLAG_SETS[i % len(FLAG_SETS)]
        if fl[3] < 100 or i % 5 == 0:      # bounded searches and some short ones: a goal nearby
            near = cells[(np.abs(cells % W - a % W) + np.abs(cells // W - a // W)) <= (6 if fl[3] < 100 else 15)]
            b = rng.choice(near)
        else:
            b = rng.choice(cells)
        q[i] = (a % W, a // W, b % W, b // W, fl[0], fl[1], fl[2], fl[3])
    q[0, 2:4] = q[0, 0:2]                  # start == goal
    return dict(W=W, H=H, is_road_map=road.astype(np.uint8), road_type_map=rtype.astype(np.uint8), allowed_dirs_map=adirs.astype(np.uint8),
                occupancy=occ.reshape(H, W), stop_map=stop.reshape(H, W), density=dens, queries=q)


def encode(paths, W):
    first = np.full(len(paths), -1, np.int32)
    off = np.zeros(len(paths) + 1, np.int64)
    steps = []
    for i, p in enumerate(paths):
        c = np.asarray([y * W + x for x, y in p], np.int64)
        off[i + 1] = off[i] + len(c)
        if len(c):
            first[i] = c[0]
            dd = np.diff(c)
            code = np.select([dd == W, dd == 1, dd == -W, dd == -1], [0, 1, 2, 3], default=255).astype(np.uint8)
            assert not (code == 255).any()
            steps.append(np.append(code, 0))   # one code per cell (the last is padding)
    return first, off, (np.concatenate(steps) if steps else np.zeros(0, np.uint8))


def main():
    from oracle.refharness import stubs
    if hasattr(stubs, "install"):
        stubs.install()
    import Simulation.utilities.pathfinding  # noqa: F401
    ref = sys.modules["Simulation.utilities.pathfinding.astar_numba"].astar_numba
    from oracle import oracle as O
    for name, (fixture, seed, nq) in CASES.items():
        inp = make_inputs(fixture, seed, nq)
        W, H = inp["W"], inp["H"]
        i8 = lambda a: np.ascontiguousarray(a.astype(np.int8))     # the dtypes CityModel holds (city_model.py:109-115)
        maps = (i8(inp["occupancy"]), i8(inp["stop_map"]), i8(inp["is_road_map"]), i8(inp["road_type_map"]), i8(inp["allowed_dirs_map"]))
        ora = O.OracleAstar(inp["occupancy"], inp["stop_map"], inp["is_road_map"], inp["road_type_map"], inp["allowed_dirs_map"], inp["density"])
        paths, found = [], 0
        for sx, sy, gx, gy, ra, so, ig, ms in inp["queries"]:
            p = [(int(x), int(y)) for x, y in ref(W, H, int(sx), int(sy), int(gx), int(gy), *maps, bool(ra), 10, inp["density"], bool(so), bool(ig), int(ms))]
            mine = ora.query(sx, sy, gx, gy, bool(ra), 10, bool(so), bool(ig), int(ms))
            assert mine == p, (name, (sx, sy, gx, gy, ra, so, ig, ms), len(p), len(mine))
            paths.append(p)
            found += bool(p)
        first, off, steps = encode(paths, W)
        out = os.path.join(HERE, f"astar_{name}.npz")
        np.savez_compressed(out, fixture=fixture, occupancy=inp["occupancy"], stop_map=inp["stop_map"], density=inp["density"],
                            queries=inp["queries"], path_first=first, path_off=off, path_steps=steps)
        print(name, "queries", len(paths), "with a path", found, "cells", int(off[-1]), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
