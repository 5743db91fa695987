"""ctypes front-end of the CPU oracle (oracle/city_oracle.c, oracle/vehicle_oracle.c, oracle/astar_oracle.c, oracle/density_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


class OCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "W", "H", "wall", "ring_w", "ring_type", "optimized", "sub_int", "sub_type", "min_sub",
        "tl_range", "fwd", "fwd_mode", "entr_level", "fast_reach")]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("city_oracle.c", "vehicle_oracle.c", "astar_oracle.c", "density_oracle.c")]
    if force or not os.path.exists(_LIB) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_city_sizeof.restype = C.c_size_t
        _lib.oracle_light_links.restype = C.c_int64
    return _lib


ROAD_CODE = {None: 0, "R1": 1, "R2": 2, "R3": 3}
FWD_MODES = ["Skip", "Include in Range", "Include as Extra"]


def make_cfg(width=200, height=200, wall_thickness=15, sidewalk_ring_width=2, ring_road_type="R2",
             optimized_intersections=True, subblock_roads_have_intersections=True,
             subblock_road_type="R3", min_subblock_spacing=5, traffic_light_range=10,
             forward_traffic_light_range=False, forward_traffic_light_range_intersections="Skip",
             block_entrance_road_level=0, fast_reach=0, **_ignored):
    return OCfg(width, height, wall_thickness, sidewalk_ring_width, ROAD_CODE[ring_road_type],
                int(optimized_intersections), int(subblock_roads_have_intersections),
                ROAD_CODE[subblock_road_type], min_subblock_spacing, traffic_light_range,
                int(forward_traffic_light_range), FWD_MODES.index(forward_traffic_light_range_intersections),
                block_entrance_road_level, fast_reach)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleCity:
    """Holds planes + band lists and runs the restated passes one by one."""

    def __init__(self, cfg: OCfg, hbands, vbands):
        self.cfg = cfg
        W, H = cfg.W, cfg.H
        self.W, self.H = W, H
        self.cell_type = np.zeros((H, W), np.uint8)
        self.dirs = np.zeros((H, W), np.uint16)
        self.aux = np.zeros((H, W), np.uint8)
        self.block_id = np.zeros((H, W), np.int32)
        self.hb = np.ascontiguousarray(hbands, np.int32).reshape(-1, 4)
        self.vb = np.ascontiguousarray(vbands, np.int32).reshape(-1, 4)
        self._city = C.create_string_buffer(lib().oracle_city_sizeof())
        lib().oracle_bind(self._city, C.byref(cfg), _p(self.cell_type, C.c_uint8), _p(self.dirs, C.c_uint16),
                          _p(self.aux, C.c_uint8), _p(self.block_id, C.c_int32),
                          _p(self.hb, C.c_int32), len(self.hb), _p(self.vb, C.c_int32), len(self.vb))
        self.n_blocks = 0
        self.entrances = None
        self.sweeps = 0

    def planes(self):
        return {"cell_type": self.cell_type, "dirs": self.dirs, "aux": self.aux, "block_id": self.block_id}

    def frame(self):
        lib().oracle_frame(self._city)

    def roads(self):
        lib().oracle_roads(self._city)

    def nothing_blobs(self):
        cap = max(16, self.W * self.H // 4)
        out = np.zeros((cap, 6), np.int32)
        n = lib().oracle_nothing_blobs(self._city, _p(out, C.c_int32), cap)
        return out[:n].copy()

    def carve(self, tape):
        tape = np.ascontiguousarray(tape, np.int32).reshape(-1, 8)
        n = lib().oracle_carve(self._city, _p(tape, C.c_int32), len(tape))
        if n < 0:
            raise ValueError("carve tape too short")
        return n

    def zones(self, zone_by_block):
        z = np.ascontiguousarray(zone_by_block, np.uint8)
        n = lib().oracle_zones(self._city, _p(z, C.c_uint8), len(z))
        if n < 0:
            raise ValueError("zone tape too short")
        self.n_blocks = n
        return n

    def dead_ends(self):
        self.sweeps = lib().oracle_dead_ends(self._city)
        return self.sweeps

    def upgrade_r2(self):
        lib().oracle_upgrade_r2(self._city)

    def entrances_pass(self, run_by_block=None):
        n = self.n_blocks
        run = np.zeros(max(n, 1), np.int32) if run_by_block is None else np.ascontiguousarray(run_by_block, np.int32)
        assert len(run) >= n
        out = np.full(max(n, 1), -1, np.int32)
        rc = lib().oracle_entrances(self._city, n, _p(run, C.c_int32), _p(out, C.c_int32))
        if rc < 0:
            raise ValueError("entrance tape index out of range")
        self.entrances = out[:n]
        return self.entrances

    def validate_dirs(self):
        lib().oracle_validate_dirs(self._city)

    def entrance_dirs(self):
        lib().oracle_entrance_dirs(self._city)

    def lights(self):
        lib().oracle_lights(self._city)
        out = {}
        for which, name in enumerate(("ctrl", "incoming", "outgoing")):
            n = lib().oracle_light_links(which, None)
            a = np.zeros((n, 2), np.int32)
            if n:
                lib().oracle_light_links(which, _p(a, C.c_int32))
            out[name] = a
        t = self.cell_type.reshape(-1)
        out["lights"] = np.flatnonzero(t == 15).astype(np.int32)
        self.links = out
        return out

    def simple_maps(self):
        maps = {k: np.zeros((self.H, self.W), np.uint8)
                for k in ("is_road_map", "road_type_map", "intersection_map", "allowed_dirs_map")}
        lib().oracle_simple_maps(self._city, _p(maps["is_road_map"], C.c_uint8), _p(maps["road_type_map"], C.c_uint8),
                                 _p(maps["intersection_map"], C.c_uint8), _p(maps["allowed_dirs_map"], C.c_uint8))
        return maps

    def run_all(self, tape_zone, tape_carve=None, tape_entrance=None, carve=False):
        self.frame()
        self.roads()
        if carve:
            self.carve(tape_carve)
        self.zones(tape_zone)
        self.dead_ends()
        self.upgrade_r2()
        self.entrances_pass(tape_entrance)
        self.validate_dirs()
        self.entrance_dirs()
        self.lights()
        return self.planes()


# ------------------------------------------------------------------------------------------------
# tick oracle (oracle/vehicle_oracle.c)
# ------------------------------------------------------------------------------------------------
class VSim(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("W", "H", "n_vehicles", "n_groups", "n_lights", "algo", "rain_enabled", "tick")] +
                [(n, C.c_void_p) for n in (
                    "occ", "stop", "stuckmap", "rain", "spawn_tick", "origin", "target", "speed", "malf", "rank",
                    "ev_first", "ev_vehicle", "ev_off", "ev_cells",
                    "pos", "path_off", "path_len", "steps",
                    "alive", "base_speed", "cur_speed", "max_steps", "early", "is_stuck", "prev_valid", "malfunction", "direction",
                    "stuck_ticks", "stranded",
                    "tl_off", "tl_cells", "g_all_off", "g_all", "g_ns_off", "g_ns", "g_ew_off", "g_ew",
                    "g_nsin_off", "g_nsin", "g_ewin_off", "g_ewin", "g_cl_off", "g_cl",
                    "g_cur", "g_pend", "g_qt", "g_gap", "g_last", "g_ft_phase", "g_ft_timer", "collision", "veh_at",
                    "g_nsout_off", "g_nsout", "g_ewout_off", "g_ewout", "g_nsp", "g_ewp", "g_nbr")])


def csr(lists, dtype=np.int32):
    off = np.zeros(len(lists) + 1, np.int32)
    off[1:] = np.cumsum([len(a) for a in lists])
    flat = np.concatenate([np.asarray(a, dtype) for a in lists]) if len(lists) and off[-1] else np.zeros(0, dtype)
    return off, np.ascontiguousarray(flat, dtype)


def light_tables_from_reference(lights, ctrl_pairs, groups):
    """CSR tables of the tick oracle from reference-extracted data.

    lights: sorted light cells; ctrl_pairs: (light cell, controlled cell) pairs; groups: list of dicts with
    cluster / lights / ns_lights / ew_lights / ns_in / ew_in arrays of cell indices (canonical order).
    """
    lights = np.asarray(lights, np.int32)
    lidx = {int(c): i for i, c in enumerate(lights)}
    per_light = [[int(c)] for c in lights]
    for l, c in np.asarray(ctrl_pairs).reshape(-1, 2):
        per_light[lidx[int(l)]].append(int(c))
    t = {}
    t["tl_off"], t["tl_cells"] = csr(per_light)
    for key, src in (("g_all", "lights"), ("g_ns", "ns_lights"), ("g_ew", "ew_lights")):
        t[key + "_off"], t[key] = csr([[lidx[int(c)] for c in g[src]] for g in groups])
    for key, src in (("g_nsin", "ns_in"), ("g_ewin", "ew_in"), ("g_cl", "cluster")):
        t[key + "_off"], t[key] = csr([g[src] for g in groups])
    for key, src in (("g_nsout", "ns_out"), ("g_ewout", "ew_out")):   # newer harness runs / fixtures only (pressure controllers)
        if all(src in g for g in groups):
            t[key + "_off"], t[key] = csr([g[src] for g in groups])
    if len(groups) and all("nbr" in g for g in groups):                # neighbour links, canonical indices (N, S, E, W)
        t["g_nbr"] = np.stack([g["nbr"] for g in groups]).astype(np.int32)
    t["n_lights"], t["n_groups"] = len(lights), len(groups)
    return t


class OracleTicks:
    """State + tapes of the tick oracle.  `tapes` needs: spawn_tick, origin, target, speed, malfunction, rank,
    ev_tick, ev_vehicle, ev_off, ev_cells (route events sorted by tick), rain_map."""

    def __init__(self, W, H, tables, tapes, n_ticks, algo=0, rain_enabled=False):
        self.W, self.H, self.n_ticks = W, H, n_ticks
        nv = len(tapes["spawn_tick"])
        self.nv = nv
        a = {}
        a["occ"] = np.zeros(W * H, np.uint8); a["stop"] = np.zeros(W * H, np.uint8); a["stuckmap"] = np.zeros(W * H, np.uint8)
        a["rain"] = np.ascontiguousarray(tapes["rain_map"], np.uint8).reshape(-1)
        for k in ("spawn_tick", "origin", "target"):
            a[k] = np.ascontiguousarray(tapes[k], np.int32)
        a["speed"] = np.ascontiguousarray(tapes["speed"], np.uint8); a["malf"] = np.ascontiguousarray(tapes["malfunction"], np.uint8)
        a["rank"] = np.ascontiguousarray(tapes["rank"], np.int32)
        ev_tick = np.asarray(tapes["ev_tick"], np.int32)
        assert np.all(np.diff(ev_tick) >= 0), "route events must be sorted by tick"
        a["ev_first"] = np.searchsorted(ev_tick, np.arange(n_ticks + 1)).astype(np.int32)
        a["ev_vehicle"] = np.ascontiguousarray(tapes["ev_vehicle"], np.int32)
        a["ev_off"] = np.ascontiguousarray(tapes["ev_off"], np.int64)
        a["ev_cells"] = np.ascontiguousarray(np.append(tapes["ev_cells"], 0), np.int32)
        for k in ("pos", "path_off", "path_len", "steps", "stranded"):
            a[k] = np.zeros(nv, np.int32)
        a["pos"][:] = -1
        for k in ("alive", "base_speed", "cur_speed", "max_steps", "early", "is_stuck", "prev_valid", "malfunction"):
            a[k] = np.zeros(nv, np.int8)
        a["direction"] = np.full(nv, -1, np.int8)
        a["stuck_ticks"] = np.zeros(nv, np.int16)
        a["collision"] = np.zeros(nv, np.int8)
        a["veh_at"] = np.full(W * H, -1, np.int32)
        for k in ("tl_off", "tl_cells", "g_all_off", "g_all", "g_ns_off", "g_ns", "g_ew_off", "g_ew",
                  "g_nsin_off", "g_nsin", "g_ewin_off", "g_ewin", "g_cl_off", "g_cl"):
            a[k] = np.ascontiguousarray(tables[k], np.int32)
        ng = tables["n_groups"]
        for k in ("g_nsout", "g_ewout"):
            if k in tables:
                a[k + "_off"], a[k] = np.ascontiguousarray(tables[k + "_off"], np.int32), np.ascontiguousarray(tables[k], np.int32)
            elif algo == 2:
                raise ValueError("PRESSURE_CONTROL needs the g_nsout / g_ewout lane tables")
            else:
                a[k + "_off"], a[k] = np.zeros(ng + 1, np.int32), np.zeros(0, np.int32)
        a["g_nsp"] = np.zeros(ng, np.int32); a["g_ewp"] = np.zeros(ng, np.int32)
        if "g_nbr" in tables:
            a["g_nbr"] = np.ascontiguousarray(tables["g_nbr"], np.int32).reshape(-1)
        elif algo == 3:
            raise ValueError("NEIGHBOR_GREEN_WAVE needs the g_nbr link table")
        else:
            a["g_nbr"] = np.full(4 * max(ng, 1), -1, np.int32)
        if algo == 2:
            # run_pressure_control hands compute_max_pressure `to_int32(occ_map)` (intersection_light_group.py:443-453): the [H, W] map
            # reshaped to (-1, 2).  Numba does not check bounds, so `occupancy_map[y, x]` reads flat element 2 * y + x of the map
            # (always inside the buffer) -- the controller counts the vehicles on THOSE cells.  Restated as it runs.
            for k in ("g_nsin", "g_ewin", "g_nsout", "g_ewout"):
                c = a[k].astype(np.int64)
                a[k] = np.ascontiguousarray(2 * (c // W) + c % W, np.int32)
        a["g_cur"] = np.full(ng, -1, np.int32); a["g_pend"] = np.zeros(ng, np.int32)   # apply_phase(0) in __init__ (:115-116)
        for k in ("g_qt", "g_gap", "g_last", "g_ft_phase", "g_ft_timer"):
            a[k] = np.zeros(ng, np.int32)
        self.a = a
        self.sim = VSim(W, H, nv, ng, tables["n_lights"], algo, int(rain_enabled), 0)
        for name, _ in VSim._fields_[8:]:
            setattr(self.sim, name, a[name].ctypes.data)
        lib().oracle_ticks_run.restype = C.c_int

    def run(self, n=1):
        rc = lib().oracle_ticks_run(C.byref(self.sim), n)
        if rc < 0:
            raise ValueError(f"tape contract violated at tick {-rc - 1} (vehicle at its target in phase A)")

    def state(self):
        a = self.a
        alive = a["alive"].astype(bool)
        pos = np.where(alive, a["pos"], -1)
        flags = (a["is_stuck"].astype(np.uint8) & 1) | ((a["malfunction"].astype(np.uint8) & 1) << 1) | ((a["direction"] + 1).astype(np.uint8) << 2) | \
                ((a["collision"].astype(np.uint8) & 1) << 5)
        return dict(pos=pos, base_speed=np.where(alive, a["base_speed"], 0), stuck_ticks=np.where(alive, a["stuck_ticks"], 0),
                    vflags=np.where(alive, flags, 0).astype(np.uint8),
                    occ=np.flatnonzero(a["occ"]).astype(np.int32), stop=np.flatnonzero(a["stop"]).astype(np.int32),
                    stuckmap=np.flatnonzero(a["stuckmap"]).astype(np.int32),
                    groups=np.stack([a["g_cur"], a["g_pend"], a["g_qt"], a["g_gap"], a["g_last"]], 1),
                    groups_ext=np.stack([a["g_ft_timer"], a["g_ft_phase"], a["g_nsp"], a["g_ewp"]], 1))


# ------------------------------------------------------------------------------------------------
# route planner oracle (oracle/astar_oracle.c)
# ------------------------------------------------------------------------------------------------
class AstarMaps(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32)] + [(n, C.c_void_p) for n in
                                                          ("occupancy", "stop_map", "is_road", "road_type", "allowed_dirs", "density")]


class OracleAstar:
    """`astar_numba(width, height, sx, sy, gx, gy, maps..., flags)` of the reference (astar_numba.py:240-281) on fixed maps.
    Maps are [H][W]; `query` returns the path as (x, y) tuples, first step first, like the reference."""

    def __init__(self, occupancy, stop_map, is_road_map, road_type_map, allowed_dirs_map, density_map=None):
        self.H, self.W = np.asarray(is_road_map).shape
        u8 = lambda a: np.ascontiguousarray(np.asarray(a).astype(np.uint8))
        self._keep = [u8(occupancy), u8(stop_map), u8(is_road_map), u8(road_type_map), u8(allowed_dirs_map),
                      None if density_map is None else np.ascontiguousarray(density_map, np.float64)]
        self.maps = AstarMaps(self.W, self.H, *[(a.ctypes.data if a is not None else None) for a in self._keep])
        self._out = np.zeros(self.W * self.H, np.int32)
        lib().oracle_astar.restype = C.c_int

    def query(self, sx, sy, gx, gy, respect_awareness=False, awareness_range=10, soft_obstacles=False, ignore_flow=False,
              maximum_steps=0x7FFFFFFF):
        n = lib().oracle_astar(C.byref(self.maps), int(sx), int(sy), int(gx), int(gy), int(respect_awareness), int(awareness_range),
                               int(soft_obstacles), int(ignore_flow), int(maximum_steps), _p(self._out, C.c_int32), len(self._out))
        if n < 0:
            raise ValueError(f"astar oracle error {n}")
        cells = self._out[:n]
        return [(int(c % self.W), int(c // self.W)) for c in cells]


def density_map(occupancy_map, is_road_map):
    """CityModel._update_density_map (city_model.py:1764-1778): float32 [H][W]."""
    occ = np.ascontiguousarray(np.asarray(occupancy_map) != 0, np.uint8)
    road = np.ascontiguousarray(np.asarray(is_road_map) != 0, np.uint8)
    H, W = occ.shape
    out = np.zeros((H, W), np.float32)
    lib().oracle_density_map(W, H, _p(occ, C.c_uint8), _p(road, C.c_uint8), _p(out, C.c_float))
    return out
