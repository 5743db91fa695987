"""Route planning joined to the tick on the GPU (replan.PlannedTraffic.on_gpu: tsim_tick_run + tsim_astar_batch + tsim_density_map
through the C ABI) against the reference fixtures WITHOUT their route events: every route the unmodified reference planned and every
per-tick state must come out of the device path.  (The file sorts after the other GPU tests: it is the newest path.)"""
import os

import numpy as np
import pytest

from golden_util import tick_fixtures, load_ticks
from planning_backends import without_routes
from test_replan_golden import check_against_fixture
from test_gpu_ticks import build_city

pytestmark = pytest.mark.gpu

CASES = [(p, n) for p in tick_fixtures() for name, n in (("s14_carve", 100), ("s31_fixed_time", 70), ("s5_stranded", 90), ("s21_sideswipe", 80)) if name in p]


@pytest.mark.parametrize("path,n_ticks", CASES, ids=lambda v: os.path.basename(v)[6:-4] if isinstance(v, str) else str(v))
def test_gpu_planned_traffic_reproduces_reference_routes(path, n_ticks):
    from trafficsimulation_b200.replan import PlannedTraffic
    from trafficsimulation_b200.traffic import light_tables_from_layout
    r = load_ticks(path)
    city = build_city(r["meta"]["cfg"], r["hbands"], r["vbands"], r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    tapes = without_routes(r)
    # a route buffer far smaller than the run needs: the compaction path runs too
    sim = PlannedTraffic.on_gpu(r["W"], r["H"], tabs, tapes, r["n_ticks"], city.maps_host(), algo=r["algo"],
                                rain_enabled=r["meta"]["rain_enabled"], route_cells=80000 if r["W"] < 200 else 400000)
    n = check_against_fixture(r, sim, n_ticks)
    assert n > 100 and sim.searches >= n // 2 and sim.compactions >= 1
    if (r["malfunction"] & 2).any():
        assert (r["vflags"][:n_ticks] & 32).any()   # sideswipe collisions happen inside the ticks that ran
