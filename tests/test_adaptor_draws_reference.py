"""The decisions `build_layout_on_gpu` draws when no tape is given (`_draw_carve_tape`, `_draw_zone_tape`) against the LIVE
reference: with `random` in the state the reference had when it entered `_carve_subblock_roads` /
`_flood_fill_blocks_storing_data` (city_model.py:563, :742), the adaptor must make the same draws -- same tapes as the harness
recorded, and the same generator state afterwards, so that everything drawn later stays aligned."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.reference

CASES = [(7, {"carve_subblock_roads": True}), (14, {"width": 150, "height": 110, "carve_subblock_roads": True}),
         (16, {"width": 64, "height": 64, "carve_subblock_roads": True}), (12345, {})]


@pytest.mark.parametrize("seed,kw", CASES, ids=[f"s{s}" for s, _ in CASES])
def test_adaptor_draws_equal_reference_draws(seed, kw):
    from oracle import oracle as O
    from oracle.refharness import harness as Hn
    from trafficsimulation_b200 import adaptor
    ref = Hn.load_reference()
    cm = ref.cm
    states = {}
    originals = {n: getattr(cm.CityModel, n) for n in ("_carve_subblock_roads", "_flood_fill_blocks_storing_data")}

    def wrap(name):
        def f(self, *a, **k):
            states[name + ":in"] = random.getstate()
            out = originals[name](self, *a, **k)
            states[name + ":out"] = random.getstate()
            return out
        return f
    try:
        for n in originals:
            setattr(cm.CityModel, n, wrap(n))
        want = Hn.run_layout(seed, keep_model=True, snapshots=("_build_roads_and_sidewalks", "_carve_subblock_roads"), **kw)
    finally:
        for n, f in originals.items():
            setattr(cm.CityModel, n, f)
    model = want["model"]
    cfg = O.make_cfg(**{k: getattr(model, k) for k in ("width", "height", "wall_thickness", "sidewalk_ring_width", "ring_road_type",
                                                        "optimized_intersections", "subblock_roads_have_intersections",
                                                        "subblock_road_type", "min_subblock_spacing", "traffic_light_range")})
    carve = bool(kw.get("carve_subblock_roads"))
    oc = O.OracleCity(cfg, want["hbands"], want["vbands"])
    oc.frame(); oc.roads()
    if carve:
        random.setstate(states["_carve_subblock_roads:in"])
        tc = adaptor._draw_carve_tape(model, oc.nothing_blobs(), ref.Defaults)
        assert random.getstate() == states["_carve_subblock_roads:out"], "generator state after the carve draws"
        assert np.array_equal(tc, want["tape_carve"])
        oc.carve(tc)
    random.setstate(states["_flood_fill_blocks_storing_data:in"])
    tz = adaptor._draw_zone_tape(oc.nothing_blobs(), ref.Defaults)
    assert random.getstate() == states["_flood_fill_blocks_storing_data:out"], "generator state after the zone draws"
    zones = list(ref.Defaults.AVAILABLE_CITY_BLOCKS)
    want_zone = [zones.index(b["block_type"]) if b["block_type"] in zones else 0 for b in model._blocks_data]
    assert tz[: len(want_zone)].tolist() == want_zone
