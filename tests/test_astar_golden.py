"""The route-planner oracle (oracle/astar_oracle.c) against the committed vectors of the live reference
(tests/golden/astar_*.npz, made by tests/golden/make_golden_astar.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
from golden_util import load_astar

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "astar_*.npz")))


def test_astar_fixtures_exist():
    assert len(FIXTURES) >= 2


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[6:-4])
def test_astar_oracle_reproduces_golden(path):
    r = load_astar(path)
    ora = O.OracleAstar(r["occupancy"], r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
    found = 0
    for q, want in zip(r["queries"], r["paths"]):
        sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in q)
        got = ora.query(sx, sy, gx, gy, bool(ra), 10, bool(so), bool(ig), ms)
        assert [y * r["W"] + x for x, y in got] == list(want), tuple(q)
        found += len(want) > 0
    assert found > len(r["queries"]) // 2
    kinds = {tuple(q[4:8]) for q in r["queries"].tolist()}
    assert len(kinds) == 6          # every flag combination vehicle_base.py uses
