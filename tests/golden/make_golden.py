"""Generates tests/golden/layout_*.npz from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Each fixture holds the inputs a parity test needs (constructor kwargs, band lists, dense tapes)
and the reference's outputs (final planes after `_add_traffic_lights`, light link tables, the
four derived maps, and a sha256 digest of the planes after every pass).  The reference itself
cannot travel to the GPU box; these files are how its answers do.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.refharness import harness as h  # noqa: E402

PASS_ORDER = ["_clear_interior", "_build_roads_and_sidewalks", "_carve_subblock_roads",
              "_flood_fill_blocks_storing_data", "_eliminate_dead_ends", "_upgrade_r2_to_intersections",
              "_final_place_block_entrances", "_remove_invalid_intersection_directions",
              "_add_entrance_directions", "_add_traffic_lights"]

CASES = {
    "default12345": (12345, {}),
    "s7": (7, {}),
    "s7_carve": (7, {"carve_subblock_roads": True}),
    "s3_carve_noint": (3, {"carve_subblock_roads": True, "subblock_roads_have_intersections": False}),
    "s11_ringR1": (11, {"ring_road_type": "R1"}),
    "s12_ringR3_carve": (12, {"ring_road_type": "R3", "carve_subblock_roads": True}),
    "s13_noopt": (13, {"optimized_intersections": False}),
    "s14_150x110_carve": (14, {"width": 150, "height": 110, "carve_subblock_roads": True}),
    "s15_96x128": (15, {"width": 96, "height": 128, "wall_thickness": 6, "sidewalk_ring_width": 1}),
    "s16_64_carve": (16, {"width": 64, "height": 64, "wall_thickness": 3, "sidewalk_ring_width": 2,
                          "carve_subblock_roads": True, "subblock_chance": 0.9}),
    "s17_tl3": (17, {"traffic_light_range": 3, "min_block_spacing": 4, "max_block_spacing": 9}),
    "s18_carveR2": (18, {"carve_subblock_roads": True, "subblock_chance": 1.0, "subblock_road_type": "R2"}),
    "s22_fwd_inrange": (22, {"forward_traffic_light_range": True,
                             "forward_traffic_light_range_intersections": "Include in Range"}),
    "s23_fwd_extra_carve": (23, {"forward_traffic_light_range": True,
                                 "forward_traffic_light_range_intersections": "Include as Extra",
                                 "carve_subblock_roads": True}),
    "s24_400x300_carve": (24, {"width": 400, "height": 300, "carve_subblock_roads": True}),
    "s26_hw4": (26, {"min_r1_bands": 4, "highway_offset_from_edges": 0}),
}


def model_cfg(model):
    return dict(width=model.width, height=model.height, wall_thickness=model.wall_thickness,
                sidewalk_ring_width=model.sidewalk_ring_width, ring_road_type=model.ring_road_type,
                optimized_intersections=bool(model.optimized_intersections),
                carve_subblock_roads=bool(model.carve_subblock_roads),
                subblock_roads_have_intersections=bool(model.subblock_roads_have_intersections),
                subblock_road_type=model.subblock_road_type,
                min_subblock_spacing=model.min_subblock_spacing,
                traffic_light_range=model.traffic_light_range,
                forward_traffic_light_range=bool(model.forward_traffic_light_range),
                forward_traffic_light_range_intersections=model.forward_traffic_light_range_intersections)


def dense_tapes(out):
    model = out["model"]
    zones = ["Residential", "Office", "Market", "Leisure", "Other"]
    n = len(model._blocks_data)
    zone = np.zeros(n, np.uint8)
    for info in model._blocks_data:
        if info["block_type"] in zones:
            zone[info["block_id"] - 1] = zones.index(info["block_type"])
    run = np.zeros(n, np.int32)
    for i, be in enumerate(model.block_entrances):
        run[be.block_id - 1] = out["tape_entrance"][i]
    return zone, run


def digest(planes, fields):
    hsh = hashlib.sha256()
    for f in fields:
        hsh.update(np.ascontiguousarray(planes[f]).tobytes())
    return hsh.hexdigest()


def main():
    for name, (seed, kw) in CASES.items():
        out = h.run_layout(seed, snapshots=PASS_ORDER, keep_model=True, **kw)
        model = out["model"]
        assert model is not None, (name, out["crashed"])
        zone, run = dense_tapes(out)
        digests = {}
        for p, planes in out["snaps"].items():
            fields = ("cell_type", "dirs") if p != "_add_traffic_lights" else ("cell_type", "dirs", "aux", "block_id")
            digests[p] = digest(planes, fields)
        meta = dict(seed=seed, kwargs=kw, cfg=model_cfg(model), digests=digests,
                    n_blocks=len(model._blocks_data))
        path = os.path.join(HERE, f"layout_{name}.npz")
        np.savez_compressed(
            path, meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
            hbands=out["hbands"], vbands=out["vbands"],
            tape_zone=zone, tape_carve=out["tape_carve"], tape_entrance=run,
            cell_type=out["final"]["cell_type"], dirs=out["final"]["dirs"], aux=out["final"]["aux"],
            block_id=out["final"]["block_id"],
            links_lights=out["links"]["lights"], links_ctrl=out["links"]["ctrl"],
            links_incoming=out["links"]["incoming"], links_outgoing=out["links"]["outgoing"],
            **out["maps"])
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
