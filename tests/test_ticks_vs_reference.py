"""Pins the C tick oracle (oracle/vehicle_oracle.c) against the LIVE reference tick loop."""
import numpy as np
import pytest

from oracle import oracle as O
from golden_util import compare_tick

pytestmark = pytest.mark.reference

CASES = [
    dict(seed=12345, n_ticks=120, spawns_per_tick=6, malfunction_p=0.002),
    dict(seed=7, n_ticks=80, spawns_per_tick=10, malfunction_p=0.0, rain_rect=(40, 40, 160, 120)),
    dict(seed=14, n_ticks=80, spawns_per_tick=4, malfunction_p=0.01,
         layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
    # sideswipe draws that fire (vehicle_base.py:567-605): collisions strand both vehicles for 600 ticks
    dict(seed=21, n_ticks=140, spawns_per_tick=12, malfunction_p=0.002, sideswipe_p=0.35),
    # the other controllers of IntersectionLightGroup.step (intersection_light_group.py:396-423)
    # (short runs: the committed fixtures ticks_s31_fixed_time / s9_pressure / s14_pressure_fwd / s9_green_wave / s14_green_wave hold longer ones)
    dict(seed=12345, n_ticks=60, spawns_per_tick=6, malfunction_p=0.002, algo="FIXED_TIME"),
    dict(seed=9, n_ticks=60, spawns_per_tick=8, malfunction_p=0.0, algo="PRESSURE_CONTROL"),
    dict(seed=14, n_ticks=60, spawns_per_tick=4, malfunction_p=0.01, algo="PRESSURE_CONTROL",
         layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True, forward_traffic_light_range=True)),
    dict(seed=12345, n_ticks=60, spawns_per_tick=8, malfunction_p=0.0, algo="NEIGHBOR_GREEN_WAVE"),
    dict(seed=14, n_ticks=60, spawns_per_tick=4, malfunction_p=0.01, algo="NEIGHBOR_GREEN_WAVE",
         layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
]
ALGO = {None: 0, "QUEUE_ACTUATED": 0, "FIXED_TIME": 1, "PRESSURE_CONTROL": 2, "NEIGHBOR_GREEN_WAVE": 3}


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c['seed']}" + ("_sideswipe" if c.get("sideswipe_p") else "") + ("_" + c["algo"] if c.get("algo") else ""))
def test_tick_oracle_matches_reference(case):
    from oracle.refharness import ticks
    r = ticks.run_ticks(**case)
    lay = r["layout"]
    tables = O.light_tables_from_reference(lay["links"]["lights"], lay["links"]["ctrl"], r["groups"])
    sim = O.OracleTicks(r["W"], r["H"], tables, r, r["n_ticks"], algo=ALGO[case.get("algo")], rain_enabled=case.get("rain_rect") is not None)
    assert (r["pos"] >= 0).any()
    if case.get("sideswipe_p"):
        assert r["sideswipes_fired"] >= 5 and (r["vflags"] & 32).any(), r["sideswipes_fired"]
    for t in range(r["n_ticks"]):
        sim.run(1)
        st = sim.state()
        compare_tick(t, st, r)
        assert np.array_equal(st["groups_ext"], r["group_ext"][t]), (t, "fixed-time timers / pressures")
    assert np.array_equal(sim.a["alive"].astype(bool) | (r["spawned"] == 0), sim.a["alive"].astype(bool) | (r["spawned"] == 0))
