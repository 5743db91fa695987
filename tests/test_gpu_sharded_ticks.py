"""Row-band shards of the tick on ONE GPU: N shards == the reference fixtures == the C oracle, tick by tick.
(The per-rank deployment runs the same code over NCCL; its host logic is covered by tests/test_sharded_tick_tables.py.)"""
import os

import numpy as np
import pytest

from golden_util import tick_fixtures, load_ticks, compare_tick
from test_gpu_ticks import build_city

pytestmark = pytest.mark.gpu

KEYS = ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups")


@pytest.mark.parametrize("n_shards,halo", [(2, 100), (3, 64)])
def test_sharded_ticks_match_reference_fixture(n_shards, halo):
    from trafficsimulation_b200.traffic import light_tables_from_layout
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    r = load_ticks(tick_fixtures()[0])
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    sim = ShardedTraffic(r["W"], r["H"], tabs, r, r["n_ticks"], n_shards, halo=halo, rain_enabled=r["meta"]["rain_enabled"])
    for t in range(r["n_ticks"]):
        sim.step(1)
        compare_tick(t, sim.state_host(), r)
    assert sim.counters()["vehicle_updates"] == int((r["pos"][:-1] >= 0).sum())


@pytest.mark.parametrize("n_shards,algo", [(2, "QUEUE_ACTUATED"), (3, "FIXED_TIME"), (4, "QUEUE_ACTUATED")])
def test_sharded_ticks_match_oracle_synthetic(n_shards, algo):
    """Dense traffic crossing the cuts all the time (1024 rows, halo 128: the windows are real sub-grids)."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.traffic import light_tables_from_layout
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    W, H, nveh, n_ticks, seed = 512, 1024, 50000, 40, 77
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = build_city(dict(width=W, height=H), hb, vb, tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    tabs = light_tables_from_layout(city)
    planes = city.planes_host()
    tp = tapes.synth_traffic(seed, W, H, planes["cell_type"], planes["dirs"], nveh, n_ticks, route_len=150, spawn_ticks=6, malfunction_p=0.001)
    sim = ShardedTraffic(W, H, tabs, tp, n_ticks, n_shards, halo=128, algo=algo)
    ora = O.OracleTicks(W, H, tabs, tp, n_ticks, algo=0 if algo == "QUEUE_ACTUATED" else 1)
    crossed = 0
    prev = None
    updates = 0
    for t in range(n_ticks):
        updates += int(ora.a["alive"].sum())
        sim.step(1)
        ora.run(1)
        got, want = sim.state_host(), ora.state()
        for k in KEYS:
            assert np.array_equal(got[k], want[k]), (t, k)
        if prev is not None:
            cut_rows = np.asarray(sim.plan.own_lo[1:])
            a, b = prev // W, got["pos"] // W
            live = (prev >= 0) & (got["pos"] >= 0)
            crossed += int((live[:, None] & ((a[:, None] < cut_rows) != (b[:, None] < cut_rows))).sum())
        prev = got["pos"]
    assert crossed > 50 * (n_shards - 1)            # vehicles really migrate between shards
    assert sim.counters()["vehicle_updates"] == updates


def test_sharded_ticks_reject_a_halo_smaller_than_a_light_group():
    from trafficsimulation_b200.traffic import light_tables_from_layout
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    r = load_ticks(tick_fixtures()[0])
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    with pytest.raises(ValueError, match="halo"):
        ShardedTraffic(r["W"], r["H"], tabs, r, r["n_ticks"], 2, halo=16)


def test_sharded_ticks_flag_a_ghost_that_differs_from_its_owner():
    """The halo refresh is a check, not only a copy: a ghost row that is not what the owner sends raises."""
    from trafficsimulation_b200 import _lib
    from trafficsimulation_b200.traffic import light_tables_from_layout
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    r = load_ticks(tick_fixtures()[0])
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    sim = ShardedTraffic(r["W"], r["H"], tabs, r, r["n_ticks"], 2, halo=100)
    sim.step(5)
    s0 = sim.sims[0]
    row = sim.plan.own_hi[0] + 3 - s0.win_y0          # a halo row of shard 0 inside its verify band
    s0.occupancy_map[row, 0] = 1                       # a wall cell: nothing will ever clear it
    with pytest.raises(_lib.TsimError, match="ghost diverged from its owner"):
        sim.step(1)
