"""Pins the C oracle (oracle/city_oracle.c) against the LIVE reference CityModel, pass by pass.

Runs only where /root/reference exists (the build container).  On the GPU box the same
guarantee travels as the fixtures in tests/golden/ (see test_oracle_golden.py).
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.reference

PASS_ORDER = ["_clear_interior", "_build_roads_and_sidewalks", "_carve_subblock_roads",
              "_flood_fill_blocks_storing_data", "_eliminate_dead_ends", "_upgrade_r2_to_intersections",
              "_final_place_block_entrances", "_remove_invalid_intersection_directions",
              "_add_entrance_directions", "_add_traffic_lights"]

CASES = [
    (12345, {}),
    (7, {}),
    (7, {"carve_subblock_roads": True}),
    (3, {"carve_subblock_roads": True, "subblock_roads_have_intersections": False}),
    (11, {"ring_road_type": "R1"}),
    (12, {"ring_road_type": "R3", "carve_subblock_roads": True}),
    (13, {"optimized_intersections": False}),
    (14, {"width": 150, "height": 110, "carve_subblock_roads": True}),
    (15, {"width": 96, "height": 128, "wall_thickness": 6, "sidewalk_ring_width": 1}),
    (16, {"width": 64, "height": 64, "wall_thickness": 3, "sidewalk_ring_width": 2,
          "carve_subblock_roads": True, "subblock_chance": 0.9}),
    (17, {"traffic_light_range": 3, "min_block_spacing": 4, "max_block_spacing": 9}),
    (18, {"carve_subblock_roads": True, "subblock_chance": 1.0, "subblock_road_type": "R2"}),
    (21, {"forward_traffic_light_range": True}),
    (22, {"forward_traffic_light_range": True, "forward_traffic_light_range_intersections": "Include in Range"}),
    (23, {"forward_traffic_light_range": True, "forward_traffic_light_range_intersections": "Include as Extra",
          "carve_subblock_roads": True}),
    (24, {"width": 400, "height": 300, "carve_subblock_roads": True}),
    (25, {"ring_road_type": "R1", "optimized_intersections": False, "carve_subblock_roads": True,
          "subblock_chance": 1.0}),
    (26, {"min_r1_bands": 4, "highway_offset_from_edges": 0}),
    (27, {"r1_chance_mean": 0.5, "r2_chance_mean": 0.2, "carve_subblock_roads": True}),
    (28, {"r1_chance_mean": 0.0, "r2_chance_mean": 0.1, "min_r1_bands": 0}),
]


def cfg_kwargs(model):
    return dict(width=model.width, height=model.height, wall_thickness=model.wall_thickness,
                sidewalk_ring_width=model.sidewalk_ring_width, ring_road_type=model.ring_road_type,
                optimized_intersections=model.optimized_intersections,
                subblock_roads_have_intersections=model.subblock_roads_have_intersections,
                subblock_road_type=model.subblock_road_type,
                min_subblock_spacing=model.min_subblock_spacing,
                traffic_light_range=model.traffic_light_range,
                forward_traffic_light_range=model.forward_traffic_light_range,
                forward_traffic_light_range_intersections=model.forward_traffic_light_range_intersections)


def dense_tapes(ref_out):
    """per-block-id zone / entrance-run tapes from the reference run."""
    model = ref_out["model"]
    zones = ["Residential", "Office", "Market", "Leisure", "Other"]
    n = len(model._blocks_data)
    zone = np.zeros(n, np.uint8)
    for info in model._blocks_data:
        if info["block_type"] in zones:
            zone[info["block_id"] - 1] = zones.index(info["block_type"])
    run = np.zeros(n, np.int32)
    for i, be in enumerate(model.block_entrances):
        run[be.block_id - 1] = ref_out["tape_entrance"][i]
    return zone, run


@pytest.mark.parametrize("seed,kw", CASES, ids=[f"s{s}-{'-'.join(k) or 'default'}" for s, k in CASES])
def test_oracle_matches_reference_per_pass(seed, kw):
    from oracle.refharness import harness as h
    ref = h.run_layout(seed, snapshots=PASS_ORDER, keep_model=True, **kw)
    model = ref["model"]
    assert model is not None, ref["crashed"]
    cfg = O.make_cfg(**cfg_kwargs(model))
    zone, run = dense_tapes(ref)
    oc = O.OracleCity(cfg, ref["hbands"], ref["vbands"])

    def check(name, fields=("cell_type", "dirs", "aux", "block_id")):
        want = ref["snaps"][name]
        got = oc.planes()
        for f in fields:
            w, g = want[f], got[f]
            if f == "aux":  # has-light / orig bits only meaningful after lights
                w = w & 0x60 if name != "_add_traffic_lights" else w
                g = g & 0x60 if name != "_add_traffic_lights" else g
            bad = np.argwhere(w != g)
            assert len(bad) == 0, (name, f, len(bad), bad[:5].tolist(),
                                   [(int(w[y, x]), int(g[y, x])) for y, x in bad[:5]])

    oc.frame()
    check("_clear_interior", ("cell_type", "dirs"))
    oc.roads()
    check("_build_roads_and_sidewalks", ("cell_type", "dirs", "aux"))
    if model.carve_subblock_roads:
        blobs = oc.nothing_blobs()
        assert len(blobs) == len(ref["tape_carve"])
        oc.carve(ref["tape_carve"])
        check("_carve_subblock_roads", ("cell_type", "dirs", "aux"))
    n = oc.zones(zone)
    assert n == len(model._blocks_data)
    check("_flood_fill_blocks_storing_data")
    oc.dead_ends()
    check("_eliminate_dead_ends")
    oc.upgrade_r2()
    check("_upgrade_r2_to_intersections")
    oc.entrances_pass(run)
    check("_final_place_block_entrances")
    oc.validate_dirs()
    check("_remove_invalid_intersection_directions")
    oc.entrance_dirs()
    check("_add_entrance_directions")
    links = oc.lights()
    check("_add_traffic_lights")
    for k in ("lights", "ctrl", "incoming", "outgoing"):
        assert np.array_equal(links[k], ref["links"][k]), k
    maps = oc.simple_maps()
    for k, v in ref["maps"].items():
        assert np.array_equal(maps[k], v), k
    # SCC-accelerated reachability must give the same answers as the literal BFS
    cfg2 = O.make_cfg(**{**cfg_kwargs(model), "fast_reach": 1})
    oc2 = O.OracleCity(cfg2, ref["hbands"], ref["vbands"])
    oc2.run_all(zone, ref["tape_carve"], run, carve=model.carve_subblock_roads)
    for f, v in oc.planes().items():
        assert np.array_equal(v, oc2.planes()[f]), f
    for k in ("ctrl", "incoming", "outgoing"):
        assert np.array_equal(oc2.links[k], links[k])
