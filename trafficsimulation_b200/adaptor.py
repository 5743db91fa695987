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


def _defaults(defaults=None):
    """The reference's own vocabulary: ``Simulation.config.Defaults`` of the process this runs in, unless the caller hands one in."""
    if defaults is not None:
        return defaults
    from Simulation.config import Defaults
    return Defaults


def fill_model_from_planes(model, planes, links, hbands=None, vbands=None, window=None, defaults=None):
    """planes: numpy [H,W] ``cell_type`` u8, ``dirs`` u16, ``aux`` u8, ``block_id`` i32; links: ``lights`` [n] cell
    indices, ``ctrl`` / ``incoming`` (/ ``outgoing`` with forward_traffic_light_range) [m,2] (light cell, cell) pairs
    (``GpuCityLayout.light_links_host()``); hbands / vbands: the band lists (int32 [n,4]) the city was built from -- the
    light groups read them (intersection_light_group.py:190).

    ``window = (x0, y0, x1, y1)``: materialise ``CellAgent`` objects only for that rectangle (a front-end's viewport on a
    city too large to hold one Python object per cell); the trackers and ``_blocks_data`` always cover the whole city.

    Everything that can be decided per plane is decided with numpy on the planes (type names and direction lists through
    look-up tables, the sparse sets -- controlled roads, lights, highway ends, tracker sets, block regions and rings -- by
    ``flatnonzero`` / sorting); the only per-cell Python work left is the reference's own ``place_cell`` call
    (city_model.py:1864-1870), which is what creates the Mesa agent."""
    if hbands is not None:
        model.horizontal_bands = bands_from_array(hbands)
    if vbands is not None:
        model.vertical_bands = bands_from_array(vbands)
    Defaults = _defaults(defaults)
    zones = list(Defaults.ZONES)
    T, D, A, B = (np.asarray(planes[k]) for k in ("cell_type", "dirs", "aux", "block_id"))
    H, W = T.shape
    x0, y0, x1, y1 = (0, 0, W, H) if window is None else window
    code_of = {z: i for i, z in enumerate(zones)}
    road_like = np.zeros(len(zones), bool)
    for z in list(Defaults.ROAD_LIKE_TYPES) + ["ControlledRoad"]:
        road_like[code_of[z]] = True
    is_zone = np.zeros(len(zones), bool)
    for z in list(Defaults.AVAILABLE_CITY_BLOCKS) + ["Empty"]:
        is_zone[code_of[z]] = True
    # ---- cells of the window: one place_cell each, attributes from look-up tables
    dir_lut = {}
    cells = {}
    place, get = model.place_cell, model.get_cell_contents
    for y in range(y0, y1):
        trow, drow = T[y].tolist(), D[y].tolist()
        for x in range(x0, x1):
            name = zones[trow[x]]
            place(x, y, name, f"{name}_{x}_{y}")
            c = get(x, y)[0]
            cells[(x, y)] = c
            code = drow[x]
            if code:
                d = dir_lut.get(code)
                if d is None:
                    d = dir_lut[code] = _decode_dirs(code)
                c.directions = list(d)
    cell_at = cells.get

    def each(mask):
        ys, xs = np.nonzero(mask)
        return zip(xs.tolist(), ys.tolist())
    # ---- sparse per-type attributes and tracker lists (raster order)
    for x, y in each(T == code_of["ControlledRoad"]):
        c = cell_at((x, y))
        if c is not None:
            c.road_type = zones[int(A[y, x]) & AUX_ORIG]
            c.base_color = Defaults.ZONE_COLORS.get(c.road_type)
            model.controlled_roads.append(c)
    for x, y in each(T == code_of["TrafficLight"]):
        c = cell_at((x, y))
        if c is not None:
            model.traffic_lights.append(c)
    for name, tracker in (("HighwayEntrance", model.highway_entrances), ("HighwayExit", model.highway_exits)):
        for x, y in each(T == code_of[name]):
            c = cell_at((x, y))
            if c is not None:
                c.highway_orientation = "horizontal" if y in (0, H - 1) else "vertical"
                c.highway_id = f"highway_{c.highway_orientation}"
                tracker.append(c)
    model._intersection_cells = set(each((A & AUX_EVER) != 0))
    model._ring_road_cells = set(each((A & AUX_RING) != 0))
    model._road_cells = set(each(road_like[T]))
    # ---- blocks in id order: region = the zone cells carrying the id (raster order), ring = their 4-neighbours outside (:795-800)
    zone_cell = is_zone[T] & (B > 0)
    flat_b = np.where(zone_cell, B, 0).reshape(-1)
    order = np.argsort(flat_b, kind="stable")
    order = order[np.searchsorted(flat_b[order], 1):]
    ids, first = np.unique(flat_b[order], return_index=True)
    bounds = list(first) + [len(order)]
    ring_pairs = []
    Bz = np.where(zone_cell, B, 0)
    for dy, dx in ((0, 1), (0, -1), (1, 0), (-1, 0)):
        src = Bz[max(0, -dy): H - max(0, dy), max(0, -dx): W - max(0, dx)]
        dst = Bz[max(0, dy): H - max(0, -dy), max(0, dx): W - max(0, -dx)]
        ys, xs = np.nonzero((src > 0) & (dst != src))
        ring_pairs.append(np.stack([src[ys, xs].astype(np.int64), ys + max(0, dy), xs + max(0, dx)], 1))
    ring_all = np.unique(np.concatenate(ring_pairs, 0), axis=0) if ring_pairs else np.zeros((0, 3), np.int64)
    ring_first = np.searchsorted(ring_all[:, 0], ids) if len(ring_all) else np.zeros(len(ids), np.int64)
    ring_last = np.searchsorted(ring_all[:, 0], ids, side="right") if len(ring_all) else np.zeros(len(ids), np.int64)
    model._blocks_data = []
    btype = {}
    for k, b in enumerate(ids.tolist()):
        cell_idx = order[bounds[k]: bounds[k + 1]]
        ys, xs = np.divmod(cell_idx, W)
        name = zones[int(T[ys[0], xs[0]])]
        btype[b] = name
        ring = ring_all[ring_first[k]: ring_last[k]]
        model._blocks_data.append({"block_id": b, "block_type": name, "region": list(zip(xs.tolist(), ys.tolist())),
                                   "ring": sorted(zip(ring[:, 2].tolist(), ring[:, 1].tolist()))})
    for x, y in each(T == code_of["BlockEntrance"]):
        c = cell_at((x, y))
        if c is not None:
            b = int(B[y, x])
            c.block_id = b
            c.block_type = btype.get(b)
            model.block_entrances.append(c)
    # ---- lights: controlled blocks, assigned lane cells, cell.light
    lights = np.asarray(links["lights"]).reshape(-1)
    if len(lights):
        model.stop_map[lights // W, lights % W] = 0
    for li, ci in np.asarray(links["ctrl"]).reshape(-1, 2).tolist():
        tl, road = cell_at((li % W, li // W)), cell_at((ci % W, ci // W))
        if tl is not None and road is not None:
            tl.controlled_blocks.append(road)
            road.light = tl
    inc = np.asarray(links["incoming"]).reshape(-1, 2)
    has_light = (A.reshape(-1)[inc[:, 1]] & AUX_LIGHT) != 0 if len(inc) else np.zeros(0, bool)
    for (li, ci), keep in zip(inc.tolist(), has_light.tolist()):
        tl, lane = cell_at((li % W, li // W)), cell_at((ci % W, ci // W))
        if tl is not None and lane is not None:
            tl.assigned_incoming_road_blocks.append(lane)
            if keep:
                lane.light = tl
    # forward_traffic_light_range (city_model.py:1550-1584): assigned outgoing cells; `.light` wherever the plane says so
    out = np.asarray(links["outgoing"]).reshape(-1, 2) if links.get("outgoing") is not None else np.zeros((0, 2), np.int64)
    out_light = (A.reshape(-1)[out[:, 1]] & AUX_LIGHT) != 0 if len(out) else np.zeros(0, bool)
    for (li, ci), keep in zip(out.tolist(), out_light.tolist()):
        tl, lane = cell_at((li % W, li // W)), cell_at((ci % W, ci // W))
        if tl is not None and lane is not None:
            tl.assigned_outgoing_road_blocks.append(lane)
            if keep and lane.light is None:
                lane.light = tl
    return cells


# ---------------------------------------------------------------------------------------------------------------
def _draw_carve_tape(model, table, defaults=None):
    """The draws of _carve_subblock_roads (city_model.py:649-682) for the blobs of `table` (minx, miny, maxx, maxy, size, root),
    in discovery order, with the global `random` module exactly as the reference calls it (same calls, same arguments,
    same order, including the up-to-20 pivot attempts and their side tests)."""
    ms = model.min_subblock_spacing
    chance = getattr(model, "subblock_chance", None)
    if chance is None:
        chance = _defaults(defaults).SUBBLOCK_CHANGE
    rows = np.zeros((len(table), 8), np.int32)
    code = {"N": 0, "E": 1, "S": 2, "W": 3}
    for i, (minx, miny, maxx, maxy, _size, _root) in enumerate(np.asarray(table).tolist()):
        if random.random() > chance:                      # :649
            continue
        rows[i, 0] = 1
        w, h = maxx - minx + 1, maxy - miny + 1
        if w < 2 * ms + 1 or h < 2 * ms + 1:              # :655
            continue
        accepted = False
        for attempt in range(20):                         # :659-675
            px, py = random.randint(minx + ms, maxx - ms), random.randint(miny + ms, maxy - ms)
            hd, vd = random.choice(["W", "E"]), random.choice(["N", "S"])
            rows[i, 7] = attempt + 1
            small_w = (px - minx) if hd == "W" else (maxx - px)
            small_h = (py - miny) if vd == "S" else (maxy - py)
            if small_w >= ms and small_h >= ms:
                rows[i, 2:6] = (px, py, code[hd], code[vd])
                accepted = True
                break
        if not accepted:
            continue
        leg = random.choice([("horizontal", "vertical"), ("vertical", "horizontal")])   # :682
        rows[i, 1] = 1
        rows[i, 6] = 1 if leg[0] == "horizontal" else 0
    return rows


def _draw_zone_tape(table, defaults=None):
    """One ``random.choices`` per block with a bounding box of at least 3 x 3, in block-id order (city_model.py:775-781);
    `table` = the component table of the labelling right before the zoning pass."""
    Defaults = _defaults(defaults)
    types = Defaults.AVAILABLE_CITY_BLOCKS
    weights = [Defaults.CITY_BLOCK_CHANCE[bt] for bt in types]
    tab = np.asarray(table)
    tz = np.zeros(max(len(tab), 1), np.uint8)
    for i in range(len(tab)):
        if tab[i, 2] - tab[i, 0] + 1 >= 3 and tab[i, 3] - tab[i, 1] + 1 >= 3:
            tz[i] = types.index(random.choices(types, weights=weights, k=1)[0])
    return tz


def build_layout_on_gpu(model, tapes=None, device="cuda:0", window=None, defaults=None):
    """Run every layout pass of ``CityModel.__init__`` (:125-139) on the GPU and fill ``model``'s grid.

    ``tapes``: optional dict with ``hbands``, ``vbands``, ``tape_zone``, ``tape_carve``, ``tape_entrance`` (a recorded
    reference run, oracle/refharness) -- replayed verbatim.  Without it the decisions are drawn here.
    ``window``: see ``fill_model_from_planes``.  ``defaults``: the reference's ``Defaults`` (imported from
    ``Simulation.config`` when omitted).
    """
    Defaults = _defaults(defaults)
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
            tc = _draw_carve_tape(model, table.cpu().numpy(), Defaults)
        city._carve_subblock_roads(tc)
    if tapes is not None:
        tz, te = tapes["tape_zone"], tapes["tape_entrance"]
    else:
        _, table = city.label_nothing()
        tz = _draw_zone_tape(table.cpu().numpy(), Defaults)
        te = None                                          # first of the longest runs (see the module docstring)
    city._flood_fill_blocks_storing_data(tz)
    city._eliminate_dead_ends()
    city._upgrade_r2_to_intersections()
    city._final_place_block_entrances(te)
    city._remove_invalid_intersection_directions(); city._add_entrance_directions()
    city._add_traffic_lights()
    fill_model_from_planes(model, city.planes_host(), city.light_links_host(), hb, vb, window=window, defaults=Defaults)
    return city


# ---------------------------------------------------------------------------------------------------------------
class GpuTickMirror:
    """The tick side of the seam: ``CityModel.step()`` (city_model.py:1831-1860) with the vehicle CA and the light groups on
    the device.

        mirror = GpuTickMirror(model, sim, vehicles)      # sim: GpuTraffic / ShardedTraffic built from the model's tapes, or a
                                                          # replan.PlannedTraffic (no route tape: the vehicles plan their routes)
        mirror.gpu_step(n)                                # instead of model.schedule.step() n times
        mirror.sync_to_model()                            # only when a front-end / statistic reads the Python objects

    ``gpu_step`` advances the device state and ``model.step_count`` and nothing else: no Python object is touched per tick.
    ``sync_to_model`` writes the device state back into the reference's own containers, the way its own movement code leaves
    them (city_model.py:1897-1963, vehicle_base.py:521-532): ``occupancy_map`` / ``stop_map`` / ``stuck_map`` (numpy [H, W]),
    and per vehicle ``pos`` (through ``model.move_vehicle`` bookkeeping: ``grid.move_agent``), ``current_speed``, ``base_speed``,
    ``is_stuck``, ``stuck_ticks``, ``direction``, ``is_in_malfunction`` -- with a ``PlannedTraffic`` also ``path`` and the planner's
    fields (``path_retry_cooldown``, ``is_overtaking``, ``overtake_path``, ``pre_overtake_path``, ``overtaking_duration`` and their
    stuck-detour twins, vehicle_base.py:43-56); vehicles that arrived are removed through
    ``model.remove_vehicle``, vehicles the device spawned are placed through ``model.place_vehicle`` (the caller supplies them
    through ``vehicles``: spawn-attempt index -> VehicleAgent, or a factory called with the attempt index).
    """

    DIRS = ("N", "E", "S", "W")

    def __init__(self, model, sim, vehicles=None, vehicle_factory=None):
        self.model, self.sim = model, sim
        self.vehicles = dict(vehicles or {})       # attempt index -> VehicleAgent currently on the model's grid
        self.factory = vehicle_factory
        self._placed = set(self.vehicles)

    def gpu_step(self, n=1, check=True):
        self.sim.step(n, check=check)
        self.model.step_count = getattr(self.model, "step_count", 0) + n

    def sync_to_model(self):
        m, st = self.model, self.sim.state_host()
        W = self.sim.W
        H = len(m.occupancy_map)
        for name, key in (("occupancy_map", "occ"), ("stop_map", "stop"), ("stuck_map", "stuckmap")):
            plane = getattr(m, name)
            plane[...] = 0
            cells = st[key]
            plane[cells // W, cells % W] = 1
        pos = st["pos"]
        live = np.flatnonzero(pos >= 0)
        for v in sorted(self._placed - set(live.tolist())):          # arrived (on_target_reached vehicle_base.py:755-775)
            ag = self.vehicles.pop(v)
            self._placed.discard(v)
            self._untrack(ag)
        flags = st["vflags"]
        planner_fields = getattr(self.sim, "planner_fields", None)
        for v in live.tolist():
            xy = (int(pos[v] % W), int(pos[v] // W))
            ag = self.vehicles.get(v)
            if ag is None:
                if self.factory is None:
                    continue
                ag = self.vehicles[v] = self.factory(v)
            if v not in self._placed:                                # place_vehicle city_model.py:1897-1908, maps already mirrored
                m.active_vehicle_agents.append(ag)
                m.grid.place_agent(ag, xy)
                if hasattr(m, "schedule"):
                    m.schedule.add(ag)
                self._placed.add(v)
            elif tuple(ag.pos) != xy:                                # move_vehicle :1945-1963
                m.grid.move_agent(ag, xy)
            ag.pos = xy
            f = int(flags[v])
            ag.base_speed = int(st["base_speed"][v])
            ag.is_stuck = bool(f & 1)
            ag.is_in_malfunction = bool(f & 2)
            ag.stuck_ticks = int(st["stuck_ticks"][v])
            d = (f >> 2) - 1
            ag.direction = self.DIRS[d] if d >= 0 else None
            if planner_fields is not None:                           # replan.PlannedTraffic: the route and the planner's own state
                for name, value in planner_fields(v).items():
                    setattr(ag, name, value)
        assert H * W == m.occupancy_map.size
        return st

    def _untrack(self, ag):
        m = self.model
        if ag in m.active_vehicle_agents:
            m.active_vehicle_agents.remove(ag)
        m.grid.remove_agent(ag)
        if hasattr(m, "schedule") and ag in getattr(m.schedule, "agents", ()):
            m.schedule.remove(ag)


def gpu_step(model, n=1):
    """``CityModel.step`` with the backend switch on: ``model._tsim_mirror`` is the ``GpuTickMirror`` the
    constructor hook installed."""
    model._tsim_mirror.gpu_step(n)
