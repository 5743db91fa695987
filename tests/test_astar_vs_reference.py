"""Pins the route-planner oracle (oracle/astar_oracle.c) against the LIVE reference planner (Numba) on fresh queries:
other seeds and densities than the committed vectors, blocked maps included."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.reference

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("fixture,seed,occ_div", [("layout_s7_carve.npz", 101, 6), ("layout_s15_96x128.npz", 202, 20),
                                                  ("layout_s26_hw4.npz", 303, 3)])
def test_astar_oracle_matches_reference(fixture, seed, occ_div):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_astar as M
    from oracle.refharness import stubs
    if hasattr(stubs, "install"):
        stubs.install()
    import Simulation.utilities.pathfinding  # noqa: F401
    ref = sys.modules["Simulation.utilities.pathfinding.astar_numba"].astar_numba
    inp = M.make_inputs(fixture, seed, 150)
    rng = np.random.default_rng(seed)
    road = np.flatnonzero(inp["is_road_map"].reshape(-1) == 1)
    occ = np.zeros(inp["W"] * inp["H"], np.uint8)
    occ[rng.choice(road, size=len(road) // occ_div, replace=False)] = 1       # much denser traffic than the vectors
    inp["occupancy"] = occ.reshape(inp["H"], inp["W"])
    W, H = inp["W"], inp["H"]
    i8 = lambda a: np.ascontiguousarray(a.astype(np.int8))
    maps = (i8(inp["occupancy"]), i8(inp["stop_map"]), i8(inp["is_road_map"]), i8(inp["road_type_map"]), i8(inp["allowed_dirs_map"]))
    ora = O.OracleAstar(inp["occupancy"], inp["stop_map"], inp["is_road_map"], inp["road_type_map"], inp["allowed_dirs_map"], inp["density"])
    found = 0
    for sx, sy, gx, gy, ra, so, ig, ms in inp["queries"]:
        want = [(int(x), int(y)) for x, y in ref(W, H, int(sx), int(sy), int(gx), int(gy), *maps, bool(ra), 10, inp["density"], bool(so), bool(ig), int(ms))]
        got = ora.query(sx, sy, gx, gy, bool(ra), 10, bool(so), bool(ig), int(ms))
        assert got == want, (sx, sy, gx, gy, ra, so, ig, ms)
        found += bool(want)
    assert found > 20
