"""ctypes binding of libtsim.so -- the C ABI declared in include/tsim.h.

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsim.so")

STATUS_NAMES = {0: "TSIM_OK", 1: "TSIM_ERR_CONFIG", 2: "TSIM_ERR_WORKSPACE", 3: "TSIM_ERR_CUDA",
                4: "TSIM_ERR_TAPE", 5: "TSIM_ERR_UNSUPPORTED", 6: "TSIM_ERR_CAPACITY"}

# every symbol include/tsim.h declares (tests check the library exports all of them)
SYMBOLS = [
    "tsim_version", "tsim_last_error", "tsim_launch_count", "tsim_build_line_table", "tsim_build_class_tables", "tsim_build_row_patterns", "tsim_workspace_bytes",
    "tsim_layout_frame_roads", "tsim_layout_label_nothing", "tsim_shard_counts", "tsim_rows_digest", "tsim_debug_write_probe", "tsim_layout_carve", "tsim_layout_zones",
    "tsim_layout_dead_ends", "tsim_layout_upgrade_r2", "tsim_layout_entrances", "tsim_layout_fix_dirs",
    "tsim_layout_lights", "tsim_lights_prepare", "tsim_lights_seed", "tsim_lights_reach", "tsim_lights_reach_planes",
    "tsim_lights_finish", "tsim_lights_eval", "tsim_lights_links", "tsim_maps", "tsim_tick_init", "tsim_tick_run", "tsim_tick_export", "tsim_tick_tiles", "tsim_tick_group_ws_bytes", "tsim_tick_probe_bytes", "tsim_debug_tick_phases", "tsim_tick_message_words", "tsim_tick_pack", "tsim_tick_unpack", "tsim_astar_scratch_bytes", "tsim_astar_batch", "tsim_density_map", "tsim_rain_discs", "tsim_label_mask",
]


class TsimError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "wall_thickness", "sidewalk_ring_width", "ring_road_type",
        "optimized_intersections", "subblock_roads_have_intersections", "subblock_road_type",
        "min_subblock_spacing", "traffic_light_range", "forward_traffic_light_range",
        "forward_intersections_mode", "block_entrance_road_level", "win_y0", "win_rows", "win_halo")]


class Planes(C.Structure):
    _fields_ = [("cell_type", C.c_void_p), ("dirs", C.c_void_p), ("aux", C.c_void_p), ("block_id", C.c_void_p)]


class Blobs(C.Structure):
    _fields_ = [("table", C.c_void_p), ("cap", C.c_int32), ("count", C.c_void_p), ("id_base", C.c_void_p)]


class Lines(C.Structure):
    _fields_ = [("row", C.c_void_p), ("col", C.c_void_p), ("row_class", C.c_void_p), ("col_class", C.c_void_p), ("lut", C.c_void_p),
                ("n_row_classes", C.c_int32), ("n_col_classes", C.c_int32),
                ("pat_type", C.c_void_p), ("pat_dirs", C.c_void_p), ("pat_aux", C.c_void_p)]


class LightLinks(C.Structure):
    _fields_ = [("n_lights", C.c_void_p), ("light_cell", C.c_void_p), ("ctrl_off", C.c_void_p),
                ("ctrl_cell", C.c_void_p), ("inc_off", C.c_void_p), ("inc_cell", C.c_void_p),
                ("cap_lights", C.c_int32), ("cap_ctrl", C.c_int32), ("cap_inc", C.c_int32),
                ("out_off", C.c_void_p), ("out_cell", C.c_void_p), ("cap_out", C.c_int32)]


class LightTables(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_lights", C.c_int32)] + [(n, C.c_void_p) for n in (
        "tl_off", "tl_cells", "g_all_off", "g_all", "g_ns_off", "g_ns", "g_ew_off", "g_ew",
        "g_nsin_off", "g_nsin", "g_ewin_off", "g_ewin", "g_cl_off", "g_cl", "g_nsout_off", "g_nsout", "g_ewout_off", "g_ewout", "g_nbr")]


class TickTapes(C.Structure):
    _fields_ = [("n_ticks", C.c_int32), ("n_vehicles", C.c_int32)] + [(n, C.c_void_p) for n in (
        "spawn_first", "origin", "target", "speed", "malfunction", "rank", "ev_first", "ev_vehicle", "ev_off", "ev_cells", "rain_map")]


class TickState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "occupancy", "stop_map", "stuck_map", "claim", "stopw",
        "pos", "path_len", "steps", "stranded", "path_off", "stuck_ticks",
        "alive", "base_speed", "cur_speed", "max_steps", "early", "is_stuck", "prev_valid", "malfunction", "direction", "moved",
        "g_cur", "g_pend", "g_qt", "g_gap", "g_last", "g_ft_phase", "g_ft_timer", "g_plan", "scalars")] + \
        [("own_row_lo", C.c_int32), ("own_row_hi", C.c_int32)] + \
        [(n, C.c_void_p) for n in ("probe", "recs", "plans", "ev_stamp", "ev_plen", "ev_poff", "live_idx", "sort_keys", "tile_ws", "group_ws", "g_wave")]


class TickStrips(C.Structure):
    _fields_ = [(n, C.c_int32 * 2) for n in ("send_lo", "send_hi", "halo_lo", "halo_hi", "verify_lo", "verify_hi")] + \
               [("send_msg", C.c_void_p * 2), ("recv_msg", C.c_void_p * 2), ("cap", C.c_int32),
                ("g_send", C.c_void_p * 2), ("n_g_send", C.c_int32 * 2),
                ("g_recv", C.c_void_p * 2), ("g_verify", C.c_void_p * 2), ("n_g_recv", C.c_int32 * 2)]


class AstarMaps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map", "density_map", "spawn_rank")]


ASTAR_QUERY_WORDS = 8   # tsim_astar_query: sx, sy, gx, gy, flags, awareness_range, maximum_steps, spawn_rank_limit (int32 each)

TICK_REC_WORDS, TICK_REC_HEADER, CELL_OUTSIDE = 12, 16, -2

ABI_VERSION = 5   # TSIM_ABI_VERSION of include/tsim.h these bindings were written against

_lib = None


def load():
    """Load libtsim.so (built by `__graft_entry__.build()` / `make -C trafficsimulation_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "trafficsimulation_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.tsim_last_error.restype = C.c_char_p
    lib.tsim_launch_count.restype = C.c_longlong
    lib.tsim_tick_message_words.restype = C.c_longlong
    for name in SYMBOLS:
        getattr(lib, name)   # AttributeError if the library does not export a declared symbol
    if lib.tsim_version() != ABI_VERSION:   # the struct layouts below are those of include/tsim.h at this version
        raise ImportError(f"{LIB_PATH} was built for TSIM_ABI_VERSION {lib.tsim_version()}, this package binds version {ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise TsimError(status, load().tsim_last_error().decode())
