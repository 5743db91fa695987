"""tapes.synth_planned_trips (the route-less tapes of bench.py's planned_trips leg) and the planning loop on them: CPU only."""
import numpy as np

from oracle import oracle as O
from golden_util import tick_fixtures, load_ticks
from planning_backends import OracleTrafficBackend, OraclePlannerBackend
from test_replan_golden import fixture_maps


def test_planned_trips_run_between_entrances_and_exits_and_arrive():
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.replan import PlannedTraffic
    r = load_ticks(tick_fixtures()[0])
    W, H = r["W"], r["H"]
    cfgd = dict(r["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads", False)
    oc = O.OracleCity(O.make_cfg(**cfgd), r["hbands"], r["vbands"])
    oc.run_all(r["tape_zone"], r["tape_carve"], r["tape_entrance"], carve=carve)
    T = oc.planes()["cell_type"].reshape(-1)
    n_ticks, per_tick = 70, 6
    tp = tapes.synth_planned_trips(3, W, H, T, per_tick, n_ticks, malfunction_p=0.001)
    assert len(tp["origin"]) == n_ticks * per_tick and len(tp["ev_tick"]) == 0
    assert np.isin(T[tp["origin"]], (19, 13)).all() and np.isin(T[tp["target"]], (19, 14)).all() and (tp["origin"] != tp["target"]).all()
    assert np.array_equal(tp["spawn_tick"], np.repeat(np.arange(n_ticks), per_tick))
    for t in (0, n_ticks - 1):   # the activation ranks of a tick are distinct
        assert len(np.unique(tp["rank"][t])) == tp["rank"].shape[1]
    maps = fixture_maps(r)
    tables = O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])
    sim = PlannedTraffic(OracleTrafficBackend(W, H, tables, tp, n_ticks), OraclePlannerBackend(W, H, maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"]),
                         W, H, maps["intersection_map"], tp, record_events=False)
    sim.step(n_ticks)
    st = sim.state_host()
    spawned = len(sim.veh)
    arrived = int(sim.traffic.sim.a["steps"][(st["pos"] < 0)].astype(bool).sum())
    assert sim.events == [] and sim.routes_planned > n_ticks * per_tick // 2
    assert arrived > 20 and spawned > 50, (arrived, spawned)     # vehicles plan, drive and reach their exits
    # every live vehicle stands on the cell its route continues from
    for v, s in sim.veh.items():
        assert s.pos == st["pos"][v] and (not s.path or abs(s.path[0] - s.pos) in (1, W))


def _default_city():
    r = load_ticks(tick_fixtures()[0])
    return r, fixture_maps(r), O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])


def test_planning_loop_edge_cases():
    """No trips at all; a trip whose target cannot be reached (no route: the vehicle waits on its entrance, later attempts on that
    cell are dropped); a trip that spawns on the last tick (its route is planned, nothing is left to hand it to)."""
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.replan import PlannedTraffic
    r, maps, tables = _default_city()
    W, H = r["W"], r["H"]
    planner = lambda: OraclePlannerBackend(W, H, maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"])
    # (a) an empty spawn tape
    tp = tapes.synth_planned_trips(1, W, H, np.where(np.arange(W * H) < 2, 19, 6), 0, 5)
    sim = PlannedTraffic(OracleTrafficBackend(W, H, tables, tp, 5), planner(), W, H, maps["intersection_map"], tp)
    sim.step(5)
    assert sim.events == [] and sim.searches == 0 and not sim.veh and len(sim.state_host()["occ"]) == 0
    # (b) + (c)
    full = tapes.synth_planned_trips(5, W, H, _entrance_plane(r), 1, 12)
    wall = int(np.flatnonzero(maps["is_road_map"].reshape(-1) == 0)[0])      # a wall cell: no route leads there
    full["target"][0] = wall
    full["origin"][3] = full["origin"][0]                                     # a later attempt on the blocked entrance
    sim = PlannedTraffic(OracleTrafficBackend(W, H, tables, full, 12), planner(), W, H, maps["intersection_map"], full)
    sim.step(12)
    st = sim.state_host()
    assert sim.veh[0].path == [] and st["pos"][0] == full["origin"][0] and st["pos"][3] == -1
    routes = {v: c for _, v, c in sim.events}
    assert routes[0] == [] and 11 in routes and len(routes[11]) > 0 and st["pos"][11] == full["origin"][11]


def _entrance_plane(r):
    cfgd = dict(r["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads", False)
    oc = O.OracleCity(O.make_cfg(**cfgd), r["hbands"], r["vbands"])
    oc.run_all(r["tape_zone"], r["tape_carve"], r["tape_entrance"], carve=carve)
    return oc.planes()["cell_type"].reshape(-1)
