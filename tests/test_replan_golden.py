"""Route planning joined to the tick (replan.PlannedTraffic) against the reference fixtures, WITHOUT their route events: the state
machine must plan, tick by tick, exactly the routes the unmodified reference planned (every route event of the fixture, spawn routes
and re-plans alike) and the city must evolve exactly as the reference's did.  CPU only: the tick and the A* searches are the C
oracles behind the device interfaces (tests/planning_backends.py); tests/test_gpu_replan.py runs the same check on the GPU."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from golden_util import tick_fixtures, load_ticks, compare_tick
from planning_backends import OracleTrafficBackend, OraclePlannerBackend, without_routes


def fixture_maps(r):
    cfgd = dict(r["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads", False)
    oc = O.OracleCity(O.make_cfg(**cfgd), r["hbands"], r["vbands"])
    oc.run_all(r["tape_zone"], r["tape_carve"], r["tape_entrance"], carve=carve)
    return oc.simple_maps()


def reference_events(r):
    ev = {}
    for i, (t, v) in enumerate(zip(r["ev_tick"].tolist(), r["ev_vehicle"].tolist())):
        ev[(t, v)] = r["ev_cells"][r["ev_off"][i]:r["ev_off"][i + 1]].tolist()   # the tick's last event of a vehicle is the one that counts
    return ev


def check_against_fixture(r, sim, n_ticks):
    want = reference_events(r)
    for t in range(n_ticks):
        sim.step(1)
        compare_tick(t, sim.traffic.state_host(), r)
    got = {}
    for t, v, cells in sim.events:
        got[(t, v)] = cells
    want = {k: c for k, c in want.items() if k[0] < n_ticks}
    assert set(got) == set(want), (sorted(set(got) - set(want))[:5], sorted(set(want) - set(got))[:5])
    bad = [k for k in want if want[k] != got[k]]
    assert not bad, (bad[:5], want[bad[0]][:12], got[bad[0]][:12])
    return len(want)


PLANNABLE = tick_fixtures()


def test_planned_traffic_with_a_small_route_buffer():
    """The append-only route buffer runs full again and again: every compaction must leave the run unchanged."""
    from trafficsimulation_b200.replan import PlannedTraffic
    r = load_ticks([p for p in PLANNABLE if "s14_carve" in p][0])
    maps = fixture_maps(r)
    tables = O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])
    tapes = without_routes(r)
    traffic = OracleTrafficBackend(r["W"], r["H"], tables, tapes, r["n_ticks"], algo=r["algo"], rain_enabled=r["meta"]["rain_enabled"],
                                   route_capacity=30000)
    planner = OraclePlannerBackend(r["W"], r["H"], maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"])
    sim = PlannedTraffic(traffic, planner, r["W"], r["H"], maps["intersection_map"], tapes, rain_enabled=r["meta"]["rain_enabled"])
    check_against_fixture(r, sim, r["n_ticks"])
    assert sim.compactions >= 5


@pytest.mark.parametrize("variant", ["speculate", "rank_chunks"])
def test_planned_traffic_batching_variants(variant, monkeypatch):
    """The two batching devices of the device path, on the CPU stand-ins: the alternatives of a search answered in one round
    (what GpuAstar does), and a tick with more spawns than one rank plane holds (planned chunk by chunk)."""
    import trafficsimulation_b200.replan as R
    r = load_ticks([p for p in PLANNABLE if "s7_rain" in p][0])   # 10 spawn attempts per tick
    maps = fixture_maps(r)
    tables = O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])
    tapes = without_routes(r)
    traffic = OracleTrafficBackend(r["W"], r["H"], tables, tapes, r["n_ticks"], algo=r["algo"], rain_enabled=r["meta"]["rain_enabled"])
    planner = OraclePlannerBackend(r["W"], r["H"], maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"])
    if variant == "speculate":
        planner.speculate = True
    else:
        monkeypatch.setattr(R, "RANK_CHUNK", 3)
    sim = R.PlannedTraffic(traffic, planner, r["W"], r["H"], maps["intersection_map"], tapes, rain_enabled=r["meta"]["rain_enabled"])
    check_against_fixture(r, sim, 60)


@pytest.mark.parametrize("path", PLANNABLE, ids=lambda p: os.path.basename(p)[6:-4])
def test_planned_traffic_reproduces_reference_routes(path):
    from trafficsimulation_b200.replan import PlannedTraffic
    r = load_ticks(path)
    maps = fixture_maps(r)
    tables = O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])
    tapes = without_routes(r)
    traffic = OracleTrafficBackend(r["W"], r["H"], tables, tapes, r["n_ticks"], algo=r["algo"], rain_enabled=r["meta"]["rain_enabled"])
    planner = OraclePlannerBackend(r["W"], r["H"], maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"])
    sim = PlannedTraffic(traffic, planner, r["W"], r["H"], maps["intersection_map"], tapes, rain_enabled=r["meta"]["rain_enabled"])
    # default12345 in full (240 ticks, 1 440 trips, 37 k planned routes); 90 ticks of the others keep the suite short
    n = check_against_fixture(r, sim, r["n_ticks"] if "default" in r["name"] else min(r["n_ticks"], 90))
    assert n > 100 and sim.searches >= n // 2
