"""GPU parity of the tick (tsim_tick_run through the C ABI) vs the reference fixtures and the C oracle."""
import os

import numpy as np
import pytest

from golden_util import tick_fixtures, load_ticks, compare_tick

pytestmark = pytest.mark.gpu


def build_city(cfgd, hb, vb, tz, tc, te):
    from trafficsimulation_b200.layout import GpuCityLayout
    cfgd = dict(cfgd)
    carve = cfgd.pop("carve_subblock_roads", False)
    city = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
    city.set_bands(hb, vb)
    city.generate(tz, tc, te)
    return city


@pytest.mark.parametrize("live_list", [True, False], ids=["live_list", "vehicle_indexed"])
@pytest.mark.parametrize("path", tick_fixtures(), ids=lambda p: os.path.basename(p)[6:-4])
def test_gpu_ticks_match_reference_fixture(path, live_list):
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    from trafficsimulation_b200.light_groups import groups_as_cell_lists
    r = load_ticks(path)
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    green = r["algo"] == "NEIGHBOR_GREEN_WAVE"   # neighbour links: marched here in the order the reference created its groups in
    tabs = light_tables_from_layout(city, links=green, creation_order=np.argsort(r["g_creation_rank"]) if green else None)
    if green:
        assert np.array_equal(tabs["g_nbr"], r["g_nbr"]) and (r["g_nbr"] >= 0).any()
    lights = city.light_links_host()["lights"]
    mine = groups_as_cell_lists(tabs, lights)
    assert len(mine) == len(r["groups"])
    for i, (a, b) in enumerate(zip(mine, r["groups"])):
        for k in a:
            if k in b:   # the out-lane lists are only in the fixtures of the controllers that read them
                assert np.array_equal(a[k], b[k]), ("group table", i, k)
    sim = GpuTraffic(r["W"], r["H"], tabs, r, r["n_ticks"], algo=r["algo"], rain_enabled=r["meta"]["rain_enabled"], live_list=live_list)
    if not live_list and (r["malfunction"] & 2).any():
        # sideswipe draws that fire are settled by the live-list kernel only; the vehicle-indexed one (shards) refuses such a tape
        from trafficsimulation_b200._lib import TsimError
        with pytest.raises(TsimError, match="flag 34"):
            sim.step(r["n_ticks"])
        return
    if (r["malfunction"] & 2).any():
        assert (r["vflags"] & 32).any()   # the fixture does hold collisions
    for t in range(r["n_ticks"]):
        sim.step(1)
        compare_tick(t, sim.state_host(), r)
    c = sim.counters()
    assert c["tick"] == r["n_ticks"] and c["vehicle_updates"] == int((r["pos"][:-1] >= 0).sum())


def test_gpu_ticks_multi_tick_launch_equals_single_ticks():
    """n ticks in one persistent launch == n launches of one tick."""
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    r = load_ticks(tick_fixtures()[0])
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    sim = GpuTraffic(r["W"], r["H"], tabs, r, r["n_ticks"], rain_enabled=r["meta"]["rain_enabled"])
    sim.step(60)
    compare_tick(59, sim.state_host(), r)
    sim.step(r["n_ticks"] - 60)
    compare_tick(r["n_ticks"] - 1, sim.state_host(), r)


@pytest.mark.parametrize("size,nveh,algo,live_list", [(512, 20000, "QUEUE_ACTUATED", True), (768, 60000, "FIXED_TIME", True),
                                                       (512, 20000, "FIXED_TIME", False), (768, 60000, "QUEUE_ACTUATED", "sorted"),
                                                       (512, 20000, "PRESSURE_CONTROL", True), (512, 20000, "PRESSURE_CONTROL", False),
                                                       (384, 12000, "NEIGHBOR_GREEN_WAVE", True), (384, 12000, "NEIGHBOR_GREEN_WAVE", False)])
def test_gpu_ticks_match_oracle_synthetic(size, nveh, algo, live_list, monkeypatch):
    """Dense synthetic traffic on a city the reference cannot plan routes for: CUDA vs the pinned C oracle."""
    if live_list == "sorted":
        monkeypatch.setenv("TSIM_TICK_SORT", "1")
        live_list = True
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    seed = size
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = build_city(dict(width=size, height=size), hb, vb, tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    tabs = light_tables_from_layout(city, links=algo == "NEIGHBOR_GREEN_WAVE")
    planes = city.planes_host()
    n_ticks = 50
    tp = tapes.synth_traffic(seed, size, size, planes["cell_type"], planes["dirs"], nveh, n_ticks, route_len=120, spawn_ticks=5,
                             malfunction_p=0.001)
    sim = GpuTraffic(size, size, tabs, tp, n_ticks, algo=algo, live_list=live_list)
    ora = O.OracleTicks(size, size, tabs, tp, n_ticks, algo={"QUEUE_ACTUATED": 0, "FIXED_TIME": 1, "PRESSURE_CONTROL": 2, "NEIGHBOR_GREEN_WAVE": 3}[algo])
    moved = 0
    prev = None
    for t in range(n_ticks):
        sim.step(1)
        ora.run(1)
        got, want = sim.state_host(), ora.state()
        for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"):
            assert np.array_equal(got[k], want[k]), (t, k)
        if prev is not None:
            moved += int(((got["pos"] != prev) & (prev >= 0)).sum())
        prev = got["pos"]
    assert moved > nveh   # traffic actually flows
    assert sim.counters()["fixed_point_iterations"] >= n_ticks


def test_gpu_ticks_without_vehicles_cycle_the_lights():
    """Edge case: an empty spawn tape.  The light groups still run their controllers; everything equals the oracle."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    size, nt = 256, 40
    hb, vb = tapes.synth_bands(9, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = GpuCityLayout(width=size, height=size)
    city.set_bands(hb, vb)
    city.generate(tapes.synth_zone_tape(9, cap), None, None)
    tabs = light_tables_from_layout(city)
    pl = city.planes_host()
    tp = tapes.synth_traffic(9, size, size, pl["cell_type"], pl["dirs"], 0, nt)
    assert len(tp["origin"]) == 0
    for algo in ("QUEUE_ACTUATED", "FIXED_TIME"):
        sim = GpuTraffic(size, size, tabs, tp, nt, algo=algo)
        ora = O.OracleTicks(size, size, tabs, tp, nt, algo=0 if algo == "QUEUE_ACTUATED" else 1)
        sim.step(nt)
        ora.run(nt)
        got, want = sim.state_host(), ora.state()
        for k in ("stop", "occ", "groups"):
            assert np.array_equal(got[k], want[k]), (algo, k)
        assert sim.counters()["vehicle_updates"] == 0


@pytest.mark.parametrize("tile_sorted", ["0", "1"], ids=["plain_append", "tile_sorted_append"])
def test_gpu_ticks_trips_injected_every_tick(tile_sorted, monkeypatch):
    """Trips injected every tick over a long run (BASELINE.json configs[3] in miniature): the live list grows and shrinks all the
    time, attempts pile up on occupied origins and are dropped; the whole state equals the oracle's every few ticks.  Both
    forms of the live list: survivors appended as they come, and appended tile by tile (what fleets of >= 200 k vehicles get)."""
    monkeypatch.setenv("TSIM_TICK_SORT", tile_sorted)
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    size, seed, n_ticks = 256, 21, 160
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = build_city(dict(width=size, height=size), hb, vb, tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    tabs = light_tables_from_layout(city)
    planes = city.planes_host()
    tp = tapes.synth_trips(seed, size, size, planes["cell_type"], planes["dirs"], trips_per_tick=40, n_ticks=n_ticks, route_len=60)
    sim = GpuTraffic(size, size, tabs, tp, n_ticks)
    ora = O.OracleTicks(size, size, tabs, tp, n_ticks)
    done = 0
    for step in (1, 1, 1, 7, 10, 20, 40, 80):
        sim.step(step)
        ora.run(step)
        done += step
        got, want = sim.state_host(), ora.state()
        for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"):
            assert np.array_equal(got[k], want[k]), (done, k)
    assert done == n_ticks and int((want["pos"] >= 0).sum()) > 0


def test_gpu_ticks_sideswipes_match_oracle():
    """Sideswipe draws that fire, in dense synthetic traffic: hundreds of candidates per tick, partners earlier and later in the
    list, chains (a vehicle hit as a partner is no candidate any more); CUDA vs the C oracle, which is pinned against the live
    reference on this rule (tests/test_ticks_vs_reference.py, sideswipe case)."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    size, nveh, n_ticks, seed = 512, 20000, 60, 77
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = build_city(dict(width=size, height=size), hb, vb, tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    tabs = light_tables_from_layout(city)
    planes = city.planes_host()
    tp = tapes.synth_traffic(seed, size, size, planes["cell_type"], planes["dirs"], nveh, n_ticks, route_len=120, spawn_ticks=5,
                             malfunction_p=0.001, sideswipe_p=0.05)
    sim = GpuTraffic(size, size, tabs, tp, n_ticks)
    ora = O.OracleTicks(size, size, tabs, tp, n_ticks)
    for t in range(n_ticks):
        sim.step(1)
        ora.run(1)
        got, want = sim.state_host(), ora.state()
        for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"):
            assert np.array_equal(got[k], want[k]), (t, k)
    assert int(((want["vflags"] & 32) != 0).sum()) >= 20   # collisions did happen
