"""The drop-in seam end to end on the GPU box: `build_layout_on_gpu(model, tapes)` fills a duck-typed
model (tests/fake_model.py: the surface of the reference's CityModel the adaptor touches) through libtsim.so, and what the
model then holds must be the reference fixture; `GpuTickMirror` mirrors the device tick state into the model's maps and
vehicle objects, which must equal the reference's per-tick fixture."""
import os
import types

import numpy as np
import pytest

from fake_model import FakeModel, extract_links, extract_planes, fake_defaults
from golden_util import layout_fixtures, load, load_ticks, tick_fixtures, PLANES

pytestmark = pytest.mark.gpu

FIX = [p for p in layout_fixtures() if any(k in p for k in ("default12345", "s7_carve", "s14_150x110_carve", "s22_fwd_inrange"))]


@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[7:-4])
def test_build_layout_on_gpu_fills_the_model_like_the_reference(path):
    from trafficsimulation_b200.adaptor import build_layout_on_gpu
    g = load(path)
    m = FakeModel(**g["meta"]["cfg"])
    tapes = {k: g[k] for k in ("hbands", "vbands", "tape_zone", "tape_carve", "tape_entrance")}
    city = build_layout_on_gpu(m, tapes=tapes, defaults=fake_defaults())
    got = extract_planes(m)
    for f in PLANES:
        assert np.array_equal(got[f], g[f]), f
    gl = extract_links(m)
    for k in ("lights", "ctrl", "incoming", "outgoing"):
        assert np.array_equal(gl[k], g["links_" + k]), k
    assert len(m._blocks_data) == g["meta"]["n_blocks"]
    assert m.horizontal_bands and m.vertical_bands and m.horizontal_bands[0][2] in ("R1", "R2", "R3")
    assert city.n_blocks == g["meta"]["n_blocks"]


def test_build_layout_on_gpu_draws_its_own_decisions():
    """No tapes: bands, carve pivots and zones are drawn with `random` in the reference's order; the result is a complete city
    that equals the oracle run on the decisions that were drawn."""
    import random
    from oracle import oracle as O
    from trafficsimulation_b200.adaptor import build_layout_on_gpu
    m = FakeModel(width=160, height=140, carve_subblock_roads=True)
    for k, v in dict(r1_chance_mean=0.15, r1_chance_std=0.03, r2_chance_mean=0.70, r2_chance_std=0.05, min_r1_bands=2, min_block_spacing=6,
                     max_block_spacing=18, highway_offset_from_edges=7, subblock_chance=0.5).items():
        setattr(m, k, v)
    random.seed(99)
    city = build_layout_on_gpu(m, defaults=fake_defaults())
    got = extract_planes(m)
    assert int((got["cell_type"] == 6).sum()) == 0 and int((got["cell_type"] == 15).sum()) > 0   # zoned, with lights
    oc = O.OracleCity(O.make_cfg(width=160, height=140, fast_reach=1), city.hbands, city.vbands)
    oc.run_all(city._zone_tape.cpu().numpy(), city._carve_tape.view(-1, 8).cpu().numpy(), None, carve=True)
    for f in PLANES:
        assert np.array_equal(got[f], oc.planes()[f]), f


def test_tick_mirror_matches_reference_fixture():
    from trafficsimulation_b200.adaptor import GpuTickMirror
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    r = load_ticks([p for p in tick_fixtures() if "default12345" in p][0])
    cfgd = dict(r["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads", False)
    city = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
    city.set_bands(r["hbands"], r["vbands"])
    city.generate(r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    sim = GpuTraffic(r["W"], r["H"], light_tables_from_layout(city), r, r["n_ticks"], rain_enabled=r["meta"]["rain_enabled"])
    m = FakeModel(**r["meta"]["cfg"])
    mirror = GpuTickMirror(m, sim, vehicle_factory=lambda v: types.SimpleNamespace(attempt=v, pos=None))
    W = r["W"]
    for t_end in (1, 25, 60, 120):
        mirror.gpu_step(t_end - m.step_count)
        assert m.step_count == t_end
        mirror.sync_to_model()
        t = t_end - 1
        want_pos = r["pos"][t]
        live = np.flatnonzero(want_pos >= 0)
        assert sorted(mirror.vehicles) == live.tolist() and len(m.active_vehicle_agents) == len(live)
        for v in live.tolist():
            ag = mirror.vehicles[v]
            assert ag.pos == (int(want_pos[v] % W), int(want_pos[v] // W)) and m.grid.where[id(ag)] == ag.pos
            f = int(r["vflags"][t][v])
            assert (ag.base_speed, ag.stuck_ticks, ag.is_stuck, ag.is_in_malfunction) == (int(r["base_speed"][t][v]), int(r["stuck_ticks"][t][v]), bool(f & 1), bool(f & 2))
        for name, key in (("occupancy_map", "occ"), ("stop_map", "stop"), ("stuck_map", "stuckmap")):
            want = r[key + "_cells"][r[key + "_off"][t]: r[key + "_off"][t + 1]]
            assert np.array_equal(np.flatnonzero(getattr(m, name).reshape(-1)), want), (t, name)
