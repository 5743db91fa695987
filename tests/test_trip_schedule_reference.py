"""the spawn tape made from the reference's pending trips (tapes.spawn_tape_from_generator) must name, tick by
tick and in order, exactly the trips the LIVE generator hands to `_spawn` (agents/dynamic_traffic_generator.py:151-189)."""
import numpy as np
import pytest

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("seed,n_ticks", [(12345, 400), (5, 250)])
def test_spawn_tape_matches_live_generator(seed, n_ticks):
    from oracle.refharness import harness as H
    from trafficsimulation_b200 import tapes
    ref = H.load_reference()
    Defaults = ref.Defaults
    saved = (Defaults.TOTAL_SERVICE_VEHICLES_FOOD, Defaults.TOTAL_SERVICE_VEHICLES_WASTE)
    Defaults.TOTAL_SERVICE_VEHICLES_FOOD = Defaults.TOTAL_SERVICE_VEHICLES_WASTE = 0
    try:
        lay = H.run_layout(seed, enable_traffic=True, enable_rain=False, keep_model=True)
        model = lay["model"]
        gen = model.dynamic_traffic_generator
        W = model.width
        tape, trips = tapes.spawn_tape_from_generator(gen, W, n_ticks)
        assert len(tape["spawn_tick"]) > 50
        seen = []          # (tick, origin cell, destination cell) of every _spawn call of the live generator
        state = {"tick": 0}
        cell = lambda a: a.position[1] * W + a.position[0]
        gen._spawn = lambda trip: seen.append((state["tick"], cell(trip.origin), cell(trip.destination)))   # no vehicles: the schedule alone
        with H._in_tmpdir():
            for t in range(n_ticks):
                state["tick"] = t
                gen.step()
    finally:
        Defaults.TOTAL_SERVICE_VEHICLES_FOOD, Defaults.TOTAL_SERVICE_VEHICLES_WASTE = saved
        Defaults.ENABLE_TRAFFIC = False
    want = np.array(seen, np.int64).reshape(-1, 3)
    got = np.stack([tape["spawn_tick"], tape["origin"], tape["target"]], 1).astype(np.int64)
    assert np.array_equal(got, want), (len(got), len(want))
    assert np.all(np.diff(tape["spawn_tick"]) >= 0)


def test_spawn_tape_clock_is_accumulated():
    """0.1 added ten times is not 1.0: the tick of a trip follows the reference's accumulated float clock."""
    from trafficsimulation_b200 import tapes
    e, edges = 0.0, []
    for _ in range(30):
        e += 0.1
        edges.append(e)
    d = np.array([edges[9], np.nextafter(edges[9], 9.0), 0.0, 3.5])
    t = tapes.spawn_tape_from_trips(d, [1, 2, 3, 4], [5, 6, 7, 8], 0.1, 30)
    assert t["trip"].tolist() == [0, 1] and t["spawn_tick"].tolist() == [9, 10]   # depart 0.0 is never spawned, 3.5 lies beyond the horizon
