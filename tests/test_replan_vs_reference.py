"""Pins the route-planning loop (trafficsimulation_b200/replan.py) against the LIVE reference: fresh seeds and parameter corners the
committed fixtures do not hold (many malfunctions -> contraflow overtakes of stranded vehicles; dense spawning -> stuck detours).
The unmodified reference runs under the harness, which records every route it plans; the loop gets the same tapes minus the routes.
Build container only (needs /root/reference); the tick and the searches are the C oracles behind the device interfaces."""
import numpy as np
import pytest

from oracle import oracle as O
from planning_backends import OracleTrafficBackend, OraclePlannerBackend, without_routes
from test_replan_golden import check_against_fixture

pytestmark = pytest.mark.reference

CASES = [
    dict(seed=5, n_ticks=90, spawns_per_tick=8, malfunction_p=0.03),
    dict(seed=33, n_ticks=70, spawns_per_tick=14, malfunction_p=0.004),
    dict(seed=27, n_ticks=90, spawns_per_tick=12, malfunction_p=0.004, sideswipe_p=0.4),   # sideswipe collisions strand both vehicles mid phase A
    dict(seed=18, n_ticks=80, spawns_per_tick=5, malfunction_p=0.02, rain_rect=(30, 30, 120, 90),
         layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c['seed']}")
def test_planned_traffic_matches_live_reference(case):
    from oracle.refharness import ticks
    from trafficsimulation_b200.replan import PlannedTraffic
    r = ticks.run_ticks(**case)
    lay = r["layout"]
    r["ev_off"], r["ev_cells"] = np.asarray(r["ev_off"]), np.asarray(r["ev_cells"])
    tables = O.light_tables_from_reference(lay["links"]["lights"], lay["links"]["ctrl"], r["groups"])
    tapes = without_routes(r)
    rain = case.get("rain_rect") is not None
    traffic = OracleTrafficBackend(r["W"], r["H"], tables, tapes, r["n_ticks"], rain_enabled=rain, route_capacity=1 << 21)
    maps = lay["maps"]
    planner = OraclePlannerBackend(r["W"], r["H"], maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"])
    sim = PlannedTraffic(traffic, planner, r["W"], r["H"], maps["intersection_map"], tapes, rain_enabled=rain)
    n = check_against_fixture(r, sim, r["n_ticks"])
    assert n > 200
    if case.get("sideswipe_p"):
        assert r["sideswipes_fired"] >= 3 and (r["vflags"] & 32).any(), r["sideswipes_fired"]


def test_tick_mirror_writes_the_planner_state_of_the_live_reference():
    """The tick side of the seam with planning (adaptor.GpuTickMirror over a PlannedTraffic): after n ticks every VehicleAgent of a
    duck-typed model holds the route AND the planner's private state (cooldown, overtake / detour flags, timers, saved routes) of the
    same vehicle in the live reference model."""
    import types
    from fake_model import FakeModel
    from oracle.refharness import ticks
    from trafficsimulation_b200.adaptor import GpuTickMirror
    from trafficsimulation_b200.replan import PlannedTraffic
    case = dict(seed=5, n_ticks=90, spawns_per_tick=8, malfunction_p=0.03)
    r = ticks.run_ticks(**case)
    lay = r["layout"]
    ref_model = lay["model"]
    tables = O.light_tables_from_reference(lay["links"]["lights"], lay["links"]["ctrl"], r["groups"])
    tapes = without_routes(r)
    maps = lay["maps"]
    sim = PlannedTraffic(OracleTrafficBackend(r["W"], r["H"], tables, tapes, r["n_ticks"], route_capacity=1 << 21),
                         OraclePlannerBackend(r["W"], r["H"], maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"]),
                         r["W"], r["H"], maps["intersection_map"], tapes)
    m = FakeModel(width=r["W"], height=r["H"])
    mirror = GpuTickMirror(m, sim, vehicle_factory=lambda v: types.SimpleNamespace(attempt=v, pos=None))
    mirror.gpu_step(r["n_ticks"])
    mirror.sync_to_model()
    ref = {ag._tsim_idx: ag for ag in ref_model.active_vehicle_agents}
    assert sorted(ref) == sorted(mirror.vehicles) and len(ref) > 100
    names = ("path", "path_retry_cooldown", "is_overtaking", "is_in_stuck_detour", "overtaking_duration", "stuck_detour_duration")
    busy = 0
    for v, want in ref.items():
        got = mirror.vehicles[v]
        assert got.pos == tuple(want.pos)
        for n in names:
            a, b = getattr(got, n), getattr(want, n)
            assert (list(map(tuple, a)) if n == "path" else a) == (list(map(tuple, b)) if n == "path" else b), (v, n, a, b)
        for n, flag in (("overtake_path", "is_overtaking"), ("pre_overtake_path", "is_overtaking"),
                        ("stuck_detour_path", "is_in_stuck_detour"), ("pre_stuck_detour_path", "is_in_stuck_detour")):
            if getattr(want, flag):   # the saved routes only mean something while the manoeuvre lasts
                assert list(map(tuple, getattr(got, n))) == list(map(tuple, getattr(want, n))), (v, n)
                busy += 1
    assert busy > 0
