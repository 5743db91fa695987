"""Tapes for config 4 in its literal form (trips injected every tick, tapes.synth_trips): valid for the tick oracle -- sorted
spawns, adjacent route steps along allowed arrows, injective activation ranks, no tape-contract violation over a whole run --
and the traffic they produce is live (vehicles spawn, move, arrive; some attempts fail on occupied cells).  CPU only."""
import numpy as np

from oracle import oracle as O
from golden_util import tick_fixtures, load_ticks
from trafficsimulation_b200 import tapes


def test_synth_trips_drive_the_tick_oracle():
    r = load_ticks(tick_fixtures()[0])                       # the reference's default city: planes + light tables
    lay = np.load(tick_fixtures()[0].replace("ticks_", "layout_"), allow_pickle=True)
    W, H = r["W"], r["H"]
    n_ticks, per_tick = 150, 40
    tp = tapes.synth_trips(3, W, H, lay["cell_type"], lay["dirs"], per_tick, n_ticks, route_len=60)
    nv = len(tp["origin"])
    assert nv > 0.9 * n_ticks * per_tick
    assert np.all(np.diff(tp["spawn_tick"]) >= 0) and tp["spawn_tick"].max() == n_ticks - 1
    assert tp["rank"].shape == (n_ticks, nv)
    for t in (0, n_ticks // 2, n_ticks - 1):
        assert len(np.unique(tp["rank"][t])) == nv           # an activation ORDER: no ties
    D = lay["dirs"].reshape(-1).astype(np.int64)
    step = np.array([W, 1, -W, -1])
    for v in (0, nv // 3, nv - 1):                           # route steps follow an arrow of the cell they leave
        cells = np.concatenate([[tp["origin"][v]], tp["ev_cells"][tp["ev_off"][v]:tp["ev_off"][v + 1]]])
        for a, b in zip(cells[:-1], cells[1:]):
            d = int(np.flatnonzero(step == b - a)[0])
            assert D[a] & (1 << d)
    tables = O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])
    sim = O.OracleTicks(W, H, tables, tp, n_ticks)
    live, arrived_any, prev_alive = [], False, None
    for t in range(n_ticks):
        sim.run(1)                                           # raises on a tape-contract violation
        alive = sim.a["alive"].astype(bool)
        live.append(int(alive.sum()))
        if prev_alive is not None and (prev_alive & ~alive).any():
            arrived_any = True
        prev_alive = alive
    spawned = int((sim.a["pos"] >= 0).sum())
    assert max(live) > 10 * per_tick and arrived_any
    assert 0 < spawned <= nv                                 # attempts on occupied cells fail, the others spawn
