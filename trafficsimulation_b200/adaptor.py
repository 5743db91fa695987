"""Drop-in seam for the reference ``CityModel``: run the layout passes on the GPU and
fill THE SAME grid the Python passes would have filled.

    # Simulation/city_model.py, where the pass methods are called (:124-139)
    if os.environ.get("TSIM_BACKEND") == "b200":
        from trafficsimulation_b200.adaptor import build_layout_on_gpu
        build_layout_on_gpu(self)
    else:
        self._place_thick_wall(); ...                       # the Python passes

Two layers:

* ``fill_model_from_planes(model, planes, links)`` -- pure host code, no GPU: turns the packed planes and the
  light link tables into the reference's own objects: one ``CellAgent`` per cell through ``model.place_cell``
  (city_model.py:1864-1870) with ``directions``, ``road_type``, ``block_id`` / ``block_type``, ``light``,
  ``controlled_blocks``, ``assigned_incoming_road_blocks``, ``highway_orientation`` / ``highway_id``, and the
  trackers ``_blocks_data``, ``block_entrances``, ``highway_entrances`` / ``highway_exits``,
  ``controlled_roads``, ``traffic_lights``, ``_intersection_cells``, ``_ring_road_cells``, ``_road_cells``
  (:96-107).  Everything downstream of the passes in ``CityModel.__init__``
  (``_create_intersection_light_groups`` :141, ``_instantiate_city_blocks`` :142, ``_build_simple_maps`` :148)
  then runs unchanged on the filled grid.  tests/test_adaptor_reference.py checks this against a model the
  reference built itself.
* ``build_layout_on_gpu(model, tapes=None)`` -- band lists from the model's own kwargs (bands.py restates the
  generator draw for draw, so ``random`` is consumed exactly as the reference would), decisions either replayed
  from ``tapes`` or drawn with ``random`` in the reference's pass order, all passes through libtsim.so, then
  ``fill_model_from_planes``.

What is NOT reproduced: the identity of ``cell.light`` when several lights claimed the same lane cell (only its
None-ness is ever read: cell.py:281,307,313), the order of the tracker lists where the reference iterates a
Python set (exported sorted), and -- when decisions are drawn here rather than replayed -- WHICH of several
equally long entrance runs the reference's ``random.choice`` over a set-ordered list would have picked
(SURVEY.md §8c).
"""
from __future__ import annotations

import random

import numpy as np

from .encoding import DIR_NAMES

AUX_ORIG, AUX_RING, AUX_EVER, AUX_LIGHT = 0x1F, 0x20, 0x40, 0x80


def _decode_dirs(code: int):
    n = (code >> 12) & 7
    return [DIR_NAMES[(code >> (4 + 2 * i)) & 3] for i in range(n)]


def bands_from_array(arr):
    """int32 [n,4] (start, end, type 1..3, dir 0..3 or -1) -> the reference's band tuples (start, end, "R1".., "N".. or "")."""
    return [(int(s), int(e), f"R{int(t)}", DIR_NAMES[int(d)] if int(d) >= 0 else "") for s, e, t, d in np.asarray(arr).reshape(-1, 4)]


def fill_model_from_planes(model, planes, links, hbands=None, vbands=None):
    """planes: numpy [H,W] ``cell_type`` u8, ``dirs`` u16, ``aux`` u8, ``block_id`` i32; links: ``lights`` [n] cell
    indices, ``ctrl`` / ``incoming`` [m,2] (light cell, cell) pairs (``GpuCityLayout.light_links_host()``);
    hbands / vbands: the band lists (int32 [n,4]) the city was built from -- the light groups read them
    (intersection_light_group.py:190)."""
    if hbands is not None:
        model.horizontal_bands = bands_from_array(hbands)
    if vbands is not None:
        model.vertical_bands = bands_from_array(vbands)
    from Simulation.config import Defaults   # the reference's own vocabulary (this runs inside its process)
    zones = list(Defaults.ZONES)
    road_like = set(Defaults.ROAD_LIKE_TYPES)
    T, D, A, B = planes["cell_type"], planes["dirs"], planes["aux"], planes["block_id"]
    H, W = T.shape
    zone_names = set(Defaults.AVAILABLE_CITY_BLOCKS) | {"Empty"}
    cells = {}
    blocks = {}
    model._intersection_cells = set()
    model._ring_road_cells = set()
    model._road_cells = set()
    for y in range(H):
        for x in range(W):
            name = zones[int(T[y, x])]
            model.place_cell(x, y, name, f"{name}_{x}_{y}")
            c = model.get_cell_contents(x, y)[0]
            cells[(x, y)] = c
            code = int(D[y, x])
            if code:
                c.directions = _decode_dirs(code)
            a = int(A[y, x])
            if name == "ControlledRoad":
                c.road_type = zones[a & AUX_ORIG]
                c.base_color = Defaults.ZONE_COLORS.get(c.road_type)
                model.controlled_roads.append(c)
            elif name == "TrafficLight":
                model.traffic_lights.append(c)
            elif name in ("HighwayEntrance", "HighwayExit"):
                c.highway_orientation = "horizontal" if y in (0, H - 1) else "vertical"
                c.highway_id = f"highway_{c.highway_orientation}"
                (model.highway_entrances if name == "HighwayEntrance" else model.highway_exits).append(c)
            if a & AUX_EVER:
                model._intersection_cells.add((x, y))
            if a & AUX_RING:
                model._ring_road_cells.add((x, y))
            if name in road_like or name == "ControlledRoad":
                model._road_cells.add((x, y))
            b = int(B[y, x])
            if b > 0 and name in zone_names:
                info = blocks.setdefault(b, {"block_id": b, "block_type": name, "region": [], "ring": set()})
                info["region"].append((x, y))
    # blocks in id order; ring = 4-neighbours of the region outside it (:795-800)
    model._blocks_data = []
    for b in sorted(blocks):
        info = blocks[b]
        reg = set(info["region"])
        for (x, y) in info["region"]:
            for nx, ny in ((x + 1, y), (x - 1, y), (x, y + 1), (x, y - 1)):
                if 0 <= nx < W and 0 <= ny < H and (nx, ny) not in reg:
                    info["ring"].add((nx, ny))
        info["ring"] = sorted(info["ring"])
        model._blocks_data.append(info)
    btype = {b: blocks[b]["block_type"] for b in blocks}
    for (x, y), c in cells.items():
        if c.cell_type == "BlockEntrance":
            b = int(B[y, x])
            c.block_id = b
            c.block_type = btype.get(b)
            model.block_entrances.append(c)
    # lights: controlled blocks, assigned lane cells, cell.light
    lights = np.asarray(links["lights"]).reshape(-1)
    for li in lights:
        model.stop_map[int(li) // W, int(li) % W] = 0
    for li, ci in np.asarray(links["ctrl"]).reshape(-1, 2):
        tl, road = cells[(int(li) % W, int(li) // W)], cells[(int(ci) % W, int(ci) // W)]
        tl.controlled_blocks.append(road)
        road.light = tl
    for li, ci in np.asarray(links["incoming"]).reshape(-1, 2):
        tl, lane = cells[(int(li) % W, int(li) // W)], cells[(int(ci) % W, int(ci) // W)]
        tl.assigned_incoming_road_blocks.append(lane)
        if int(A[int(ci) // W, int(ci) % W]) & AUX_LIGHT:
            lane.light = tl
    return cells


# ---------------------------------------------------------------------------------------------------------------
def _draw_carve_tape(model, table):
    """The draws of _carve_subblock_roads (city_model.py:649-682) for the blobs of `table` (minx, miny, maxx, maxy, size, root),
    in discovery order, with the global `random` module exactly as the reference calls it."""
    from Simulation.config import Defaults
    ms = model.min_subblock_spacing
    chance = getattr(model, "subblock_chance", Defaults.SUBBLOCK_CHANGE)
    rows = np.zeros((len(table), 8), np.int32)
    code = {"N": 0, "E": 1, "S": 2, "W": 3}
    for i, (minx, miny, maxx, maxy, size, _root) in enumerate(table.tolist()):
        if random.random() > chance:                      # :649
            continue
        rows[i, 0] = 1
        w, h = maxx - minx + 1, maxy - miny + 1
        if w < 2 * ms + 1 or h < 2 * ms + 1:              # :655
            continue
        for attempt in range(20):                         # :659-675
            px, py = random.randint(minx + ms, maxx - ms), random.randint(miny + ms, maxy - ms)
            hd, vd = random.choice(["W", "E"]), random.choice(["N", "S"])
            rows[i, 7] = attempt + 1
            rows[i, 2:6] = (px, py, code[hd], code[vd])
            break                                         # a pivot inside [min+ms, max-ms] always passes the side tests
        leg = random.choice([("horizontal", "vertical"), ("vertical", "horizontal")])   # :682
        rows[i, 1] = 1
        rows[i, 6] = 1 if leg[0] == "horizontal" else 0
    return rows


def build_layout_on_gpu(model, tapes=None, device="cuda:0"):
    """Run every layout pass of ``CityModel.__init__`` (:125-139) on the GPU and fill ``model``'s grid.

    ``tapes``: optional dict with ``hbands``, ``vbands``, ``tape_zone``, ``tape_carve``, ``tape_entrance`` (a recorded
    reference run, oracle/refharness) -- replayed verbatim.  Without it the decisions are drawn here.
    """
    from Simulation.config import Defaults
    from .bands import BandParams, bands_to_array, make_city_bands
    from .layout import GpuCityLayout
    kw = dict(width=model.width, height=model.height, wall_thickness=model.wall_thickness,
              sidewalk_ring_width=model.sidewalk_ring_width, ring_road_type=model.ring_road_type,
              optimized_intersections=model.optimized_intersections, carve_subblock_roads=model.carve_subblock_roads,
              subblock_roads_have_intersections=model.subblock_roads_have_intersections, subblock_road_type=model.subblock_road_type,
              min_subblock_spacing=model.min_subblock_spacing, traffic_light_range=model.traffic_light_range,
              forward_traffic_light_range=model.forward_traffic_light_range,
              forward_traffic_light_range_intersections=model.forward_traffic_light_range_intersections,
              block_entrance_road_level=getattr(Defaults, "BLOCK_ENTRANCE_ROAD_LEVEL", 0))
    city = GpuCityLayout(device=device, **kw)
    if tapes is not None:
        hb, vb = tapes["hbands"], tapes["vbands"]
    else:
        bp = BandParams(width=model.width, height=model.height, wall_thickness=model.wall_thickness, sidewalk_ring_width=model.sidewalk_ring_width,
                        ring_road_type=model.ring_road_type, r1_chance_mean=model.r1_chance_mean, r1_chance_std=model.r1_chance_std,
                        r2_chance_mean=model.r2_chance_mean, r2_chance_std=model.r2_chance_std, min_r1_bands=model.min_r1_bands,
                        min_block_spacing=model.min_block_spacing, max_block_spacing=model.max_block_spacing,
                        highway_offset_from_edges=model.highway_offset_from_edges)
        hbl, vbl = make_city_bands(bp)                     # consumes `random` exactly like city_model.py:380-394
        model.horizontal_bands, model.vertical_bands = hbl, vbl
        hb, vb = bands_to_array(hbl), bands_to_array(vbl)
    city.set_bands(hb, vb)
    city._place_thick_wall(); city._place_sidewalk_inner_ring(); city._clear_interior()
    city._build_roads_and_sidewalks()
    if model.carve_subblock_roads:
        if tapes is not None:
            tc = tapes["tape_carve"]
        else:
            _, table = city.label_nothing()
            tc = _draw_carve_tape(model, table.cpu().numpy())
        city._carve_subblock_roads(tc)
    if tapes is not None:
        tz, te = tapes["tape_zone"], tapes["tape_entrance"]
    else:
        # one random.choices per block with a bounding box of at least 3 x 3, in block-id order (:781)
        n, table = city.label_nothing()
        tab = table.cpu().numpy()
        types = Defaults.AVAILABLE_CITY_BLOCKS
        weights = [Defaults.CITY_BLOCK_CHANCE[bt] for bt in types]
        tz = np.zeros(max(n, 1), np.uint8)
        for i in range(n):
            if tab[i, 2] - tab[i, 0] + 1 >= 3 and tab[i, 3] - tab[i, 1] + 1 >= 3:
                tz[i] = types.index(random.choices(types, weights=weights, k=1)[0])
        te = None                                          # first of the longest runs (see the module docstring)
    city._flood_fill_blocks_storing_data(tz)
    city._eliminate_dead_ends()
    city._upgrade_r2_to_intersections()
    city._final_place_block_entrances(te)
    city._remove_invalid_intersection_directions(); city._add_entrance_directions()
    city._add_traffic_lights()
    fill_model_from_planes(model, city.planes_host(), city.light_links_host(), hb, vb)
    return city
