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
