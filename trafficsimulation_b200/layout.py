"""Host-side mirror of the reference's layout interface, backed by libtsim.so (sm_100a kernels).

``GpuCityLayout`` takes the same constructor kwargs as the reference ``CityModel``
(Simulation/city_model.py:27-53) and exposes one method per reference pass, with the reference's
names, so parity tests read like the reference's own build sequence (city_model.py:124-148):

    city = GpuCityLayout(width=..., height=..., ...)
    city.set_bands(hbands, vbands)                  # L1a: host band lists (bands.make_city_bands)
    city._place_thick_wall(); city._place_sidewalk_inner_ring(); city._clear_interior()
    city._build_roads_and_sidewalks()               # the four above are ONE fused kernel
    city._carve_subblock_roads(tape_carve)
    city._flood_fill_blocks_storing_data(tape_zone)
    city._eliminate_dead_ends(); city._upgrade_r2_to_intersections()
    city._final_place_block_entrances(tape_entrance)
    city._remove_invalid_intersection_directions(); city._add_entrance_directions()
    city._add_traffic_lights()
    city._build_simple_maps()

PyTorch is used for device buffers and streams only.  There is no CPU path: without a CUDA device
or without libtsim.so every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .encoding import FORWARD_MODES, ROAD_CODE

BLOB_STRIDE = 6


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class GpuCityLayout:
    def __init__(self, width=200, height=200, wall_thickness=15, sidewalk_ring_width=2, ring_road_type="R2",
                 optimized_intersections=True, carve_subblock_roads=False, subblock_roads_have_intersections=True,
                 subblock_road_type="R3", min_subblock_spacing=5, traffic_light_range=10,
                 forward_traffic_light_range=False, forward_traffic_light_range_intersections="Skip",
                 block_entrance_road_level=0, device="cuda:0", win_y0=0, win_rows=None, win_halo=0, frame_tables="rows", **_unused_reference_kwargs):
        """``win_y0`` / ``win_rows``: this object holds only the global rows [win_y0, win_y0 + win_rows) of the
        width x height city (a row-band shard window, see sharded.py); default = the whole grid."""
        if not torch.cuda.is_available():
            raise RuntimeError("trafficsimulation_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.width, self.height = int(width), int(height)
        self.win_y0 = int(win_y0)
        self.win_rows = int(self.height - self.win_y0 if win_rows is None else win_rows)
        self.carve_subblock_roads = bool(carve_subblock_roads)
        self.frame_tables = frame_tables   # "rows" (pattern rows), "lut" (class look-up), "none" (closed form per cell): same result, three kernels
        self.cfg = _lib.Cfg(self.width, self.height, wall_thickness, sidewalk_ring_width, ROAD_CODE[ring_road_type],
                            int(optimized_intersections), int(subblock_roads_have_intersections),
                            ROAD_CODE[subblock_road_type], min_subblock_spacing, traffic_light_range,
                            int(forward_traffic_light_range), FORWARD_MODES.index(forward_traffic_light_range_intersections),
                            block_entrance_road_level, self.win_y0, self.win_rows, int(win_halo))
        n = self.width * self.win_rows
        dev = self.device
        self.cell_type = torch.empty(n, dtype=torch.uint8, device=dev)
        self.dirs = torch.empty(n, dtype=torch.int16, device=dev)      # u16 bit patterns
        self.aux = torch.empty(n, dtype=torch.uint8, device=dev)
        self.block_id = torch.zeros(n, dtype=torch.int32, device=dev)   # written as a whole by the zoning pass
        self._planes = _lib.Planes(self.cell_type.data_ptr(), self.dirs.data_ptr(), self.aux.data_ptr(), self.block_id.data_ptr())
        ws = C.c_size_t(0)
        _lib.check(self.lib.tsim_workspace_bytes(C.byref(self.cfg), C.byref(ws)))
        self.workspace = torch.empty(ws.value, dtype=torch.uint8, device=dev)
        # [0] err flag, [1] sweeps, [2] n_blobs, [3] n_lights, [4] id_base, [8..9] pivot candidates, [10] pivot, [11] reach changed
        self.flags = torch.zeros(16, dtype=torch.int32, device=dev)
        self.blobs = None
        self.n_blocks = 0
        self.entrances = None
        self.links = None
        self.maps = None
        self._lines = None
        self._frame_done = False

    # ------------------------------------------------------------------ helpers
    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _flag_ptr(self, i):
        return C.c_void_p(self.flags.data_ptr() + 4 * i)

    def _check_flag(self, what):
        v = int(self.flags[0].item())
        if v:
            self.flags[0] = 0
            raise _lib.TsimError(4 if v < 10 else 6, f"{what}: device error flag {v} (DESIGN.md §9 lists the codes)")

    def set_bands(self, hbands, vbands):
        """Band lists as int32 [n,4] (start, end, type 1..3, dir 0..3 or -1); city_model.py:380-394."""
        hb = np.ascontiguousarray(hbands, np.int32).reshape(-1, 4)
        vb = np.ascontiguousarray(vbands, np.int32).reshape(-1, 4)
        row = np.zeros(self.height, np.uint32)
        col = np.zeros(self.width, np.uint32)
        _lib.check(self.lib.tsim_build_line_table(hb.ctypes.data_as(C.c_void_p), len(hb), self.height, row.ctypes.data_as(C.c_void_p)))
        _lib.check(self.lib.tsim_build_line_table(vb.ctypes.data_as(C.c_void_p), len(vb), self.width, col.ctypes.data_as(C.c_void_p)))
        self.hbands, self.vbands = hb, vb
        self.row_table = torch.from_numpy(row.view(np.int32)).to(self.device)
        self.col_table = torch.from_numpy(col.view(np.int32)).to(self.device)
        # class tables for the bulk path of the frame pass (optional: too many classes -> closed form everywhere)
        rc, cc = np.zeros(self.height, np.uint8), np.zeros(self.width, np.uint8)
        lut = np.zeros(256 * 256, np.uint32)
        nr, nc = C.c_int32(0), C.c_int32(0)
        st = self.lib.tsim_build_class_tables(C.byref(self.cfg), row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                                              rc.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p), lut.ctypes.data_as(C.c_void_p),
                                              len(lut), C.byref(nr), C.byref(nc))
        if self.frame_tables == "none":
            st = 6
        if st == 0:
            self.row_class = torch.from_numpy(rc).to(self.device)
            self.col_class = torch.from_numpy(cc).to(self.device)
            self.class_lut = torch.from_numpy(lut[: nr.value * nc.value].view(np.int32)).to(self.device)
            # the bulk part of every row class as a ready-made row (a few MB, stays in L2)
            pt = np.zeros((nr.value, self.width), np.uint8)
            pd = np.zeros((nr.value, self.width), np.uint16)
            pa = np.zeros((nr.value, self.width), np.uint8)
            _lib.check(self.lib.tsim_build_row_patterns(C.byref(self.cfg), row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p),
                                                        rc.ctypes.data_as(C.c_void_p), nr.value, pt.ctypes.data_as(C.c_void_p),
                                                        pd.ctypes.data_as(C.c_void_p), pa.ctypes.data_as(C.c_void_p)))
            self._patterns = (torch.from_numpy(pt).to(self.device), torch.from_numpy(pd.view(np.int16)).to(self.device), torch.from_numpy(pa).to(self.device))
            pats = [t.data_ptr() for t in self._patterns] if self.frame_tables == "rows" else [0, 0, 0]
            self._lines = _lib.Lines(self.row_table.data_ptr(), self.col_table.data_ptr(), self.row_class.data_ptr(), self.col_class.data_ptr(),
                                     self.class_lut.data_ptr(), nr.value, nc.value, *pats)
        elif st == 6:   # TSIM_ERR_CAPACITY: not an error, just no tables
            self._lines = _lib.Lines(self.row_table.data_ptr(), self.col_table.data_ptr(), 0, 0, 0, 0, 0, 0, 0, 0)
        else:
            _lib.check(st)
        # capacity of the component tables: every rectangle of the band grid can split in at most 3
        cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
        if self.win_rows < self.height:   # a window sees its share of the band grid (plus slack for cut components)
            cap = int(cap * min(1.0, 1.5 * self.win_rows / self.height)) + 4 * (len(vb) + 2) + 64
        self.blob_cap = cap
        self.blobs = torch.zeros(cap * BLOB_STRIDE, dtype=torch.int32, device=self.device)
        self._blobs = _lib.Blobs(self.blobs.data_ptr(), cap, self.flags.data_ptr() + 4 * 2, self.flags.data_ptr() + 4 * 4)   # count = flags[2], id_base = flags[4]
        self._frame_done = False

    # ------------------------------------------------------------------ passes (reference names)
    def _place_thick_wall(self):            # city_model.py:315 -- fused into _build_roads_and_sidewalks
        self._frame_done = True

    def _place_sidewalk_inner_ring(self):   # city_model.py:329 -- fused
        self._frame_done = True

    def _clear_interior(self):              # city_model.py:366 -- fused
        self._frame_done = True

    def _build_roads_and_sidewalks(self):   # city_model.py:375
        if self._lines is None:
            raise RuntimeError("set_bands() first")
        _lib.check(self.lib.tsim_layout_frame_roads(C.byref(self.cfg), C.byref(self._planes), C.byref(self._lines), self._stream))

    def label_nothing(self):
        """Label the current `Nothing` components; returns (n, table[n,6]) on the host (synchronises)."""
        self._label_async()
        n = int(self.flags[2].item())
        self._check_flag("label_nothing")
        if n > self.blob_cap:
            raise _lib.TsimError(6, f"{n} components exceed the table capacity {self.blob_cap}")
        return n, self.blobs[: n * BLOB_STRIDE].view(n, BLOB_STRIDE)

    def _label_async(self):
        _lib.check(self.lib.tsim_layout_label_nothing(C.byref(self.cfg), C.byref(self._planes), C.byref(self._blobs), self._flag_ptr(0),
                                                      _ptr(self.workspace), C.c_size_t(self.workspace.numel()), self._stream))

    def _carve_subblock_roads(self, tape_carve, check=True, relabel=True):   # city_model.py:563
        tape = tape_carve if isinstance(tape_carve, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(tape_carve, np.int32)).to(self.device)
        tape = tape.contiguous().view(-1)
        if relabel:
            self._label_async()
        self._carve_tape = tape
        _lib.check(self.lib.tsim_layout_carve(C.byref(self.cfg), C.byref(self._planes), C.byref(self._lines), C.byref(self._blobs),
                                              _ptr(tape), tape.numel() // 8, self._flag_ptr(0), self._stream))
        if check:
            self._check_flag("_carve_subblock_roads")

    def _flood_fill_blocks_storing_data(self, tape_zone, check=True, relabel=True):   # city_model.py:742
        z = tape_zone if isinstance(tape_zone, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(tape_zone, np.uint8)).to(self.device)
        self._zone_tape = z
        if relabel:
            self._label_async()
        _lib.check(self.lib.tsim_layout_zones(C.byref(self.cfg), C.byref(self._planes), C.byref(self._blobs), _ptr(z), z.numel(),
                                              self._flag_ptr(0), _ptr(self.workspace), C.c_size_t(self.workspace.numel()), self._stream))
        if check:
            self.n_blocks = int(self.flags[2].item())
            self._check_flag("_flood_fill_blocks_storing_data")

    def _eliminate_dead_ends(self):   # city_model.py:811
        _lib.check(self.lib.tsim_layout_dead_ends(C.byref(self.cfg), C.byref(self._planes), self._flag_ptr(1), _ptr(self.workspace),
                                                  C.c_size_t(self.workspace.numel()), self._stream))

    def _upgrade_r2_to_intersections(self, check=True):   # city_model.py:842
        _lib.check(self.lib.tsim_layout_upgrade_r2(C.byref(self.cfg), C.byref(self._planes), C.byref(self._lines), self._flag_ptr(0), self._stream))
        if check:
            self._check_flag("_upgrade_r2_to_intersections")

    def _final_place_block_entrances(self, tape_entrance=None, check=True):   # city_model.py:884
        n_tape = self.blob_cap if tape_entrance is None else len(tape_entrance)
        if tape_entrance is None:
            run = torch.zeros(n_tape, dtype=torch.int32, device=self.device)
        elif isinstance(tape_entrance, torch.Tensor):
            run = tape_entrance
        else:
            run = torch.from_numpy(np.ascontiguousarray(tape_entrance, np.int32)).to(self.device)
        self._run_tape = run
        if self.entrances is None or self.entrances.numel() != self.blob_cap:   # rows < n_blocks are rewritten by every call
            self.entrances = torch.full((self.blob_cap,), -1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.tsim_layout_entrances(C.byref(self.cfg), C.byref(self._planes), C.byref(self._blobs), _ptr(run),
                                                  n_tape, _ptr(self.entrances), self._flag_ptr(0), _ptr(self.workspace),
                                                  C.c_size_t(self.workspace.numel()), self._stream))
        if check:
            self._check_flag("_final_place_block_entrances")

    def _remove_invalid_intersection_directions(self):   # city_model.py:969 -- runs together with the next pass
        self._fix_pending = True

    def _add_entrance_directions(self):   # city_model.py:1035
        _lib.check(self.lib.tsim_layout_fix_dirs(C.byref(self.cfg), C.byref(self._planes), self._stream))
        self._fix_pending = False

    def _lights_prepare(self):
        _lib.check(self.lib.tsim_lights_prepare(C.byref(self.cfg), C.byref(self._planes), self._flag_ptr(8), self._flag_ptr(0), _ptr(self.workspace),
                                                C.c_size_t(self.workspace.numel()), self._stream))

    def _lights_seed(self):   # flags[10] = window cell of the agreed pivot (negative: outside)
        _lib.check(self.lib.tsim_lights_seed(C.byref(self.cfg), self._flag_ptr(10), _ptr(self.workspace), C.c_size_t(self.workspace.numel()), self._stream))

    def _lights_reach(self, edge_rows=-1):  # flags[11] is set when a bit was added
        _lib.check(self.lib.tsim_lights_reach(C.byref(self.cfg), C.c_int32(edge_rows), self._flag_ptr(11), self._flag_ptr(0), _ptr(self.workspace),
                                              C.c_size_t(self.workspace.numel()), self._stream))

    def reach_planes(self):
        """The two reachability bit-planes inside the workspace as int64 tensors [win_rows, words_per_row]."""
        fw, bw, wp = C.c_size_t(0), C.c_size_t(0), C.c_int32(0)
        _lib.check(self.lib.tsim_lights_reach_planes(C.byref(self.cfg), C.c_size_t(self.workspace.numel()), C.byref(fw), C.byref(bw), C.byref(wp)))
        nb = self.win_rows * wp.value * 8
        return (self.workspace[fw.value: fw.value + nb].view(torch.int64).view(self.win_rows, wp.value),
                self.workspace[bw.value: bw.value + nb].view(torch.int64).view(self.win_rows, wp.value))

    def _lights_finish(self, check=True):
        lk = self._links_struct()
        _lib.check(self.lib.tsim_lights_finish(C.byref(self.cfg), C.byref(self._planes), C.byref(lk), self._flag_ptr(0), _ptr(self.workspace),
                                               C.c_size_t(self.workspace.numel()), self._stream))
        if check:
            self._check_flag("_add_traffic_lights")

    def _lights_eval(self):    # staged form of _add_traffic_lights: _lights_prepare, _lights_eval, _lights_links
        lk = self._links_struct()
        _lib.check(self.lib.tsim_lights_eval(C.byref(self.cfg), C.byref(self._planes), C.byref(lk), self._flag_ptr(0), _ptr(self.workspace),
                                             C.c_size_t(self.workspace.numel()), self._stream))

    def _lights_links(self, check=True):
        lk = self._links_struct()
        _lib.check(self.lib.tsim_lights_links(C.byref(self.cfg), C.byref(self._planes), C.byref(lk), self._flag_ptr(0), _ptr(self.workspace),
                                              C.c_size_t(self.workspace.numel()), self._stream))
        if check:
            self._check_flag("_add_traffic_lights")

    def _links_struct(self):
        n = self.width * self.win_rows
        cap_l, cap_c, cap_i = max(1024, n // 16), max(1024, n // 8), max(4096, n // 2)
        dev = self.device
        t = getattr(self, "_link_tensors", None)
        if t is None:   # allocated once, reused by every call
            t = dict(n_lights=self.flags[3:4], light_cell=torch.empty(cap_l, dtype=torch.int32, device=dev),
                     ctrl_off=torch.empty(cap_l + 1, dtype=torch.int32, device=dev), ctrl_cell=torch.empty(cap_c, dtype=torch.int32, device=dev),
                     inc_off=torch.empty(cap_l + 1, dtype=torch.int32, device=dev), inc_cell=torch.empty(cap_i, dtype=torch.int32, device=dev))
            if self.cfg.forward_traffic_light_range:
                t["out_off"] = torch.empty(cap_l + 1, dtype=torch.int32, device=dev)
                t["out_cell"] = torch.empty(cap_i, dtype=torch.int32, device=dev)
            self._link_tensors = t
        fwd = "out_off" in t
        return _lib.LightLinks(self._flag_ptr(3), t["light_cell"].data_ptr(), t["ctrl_off"].data_ptr(), t["ctrl_cell"].data_ptr(),
                               t["inc_off"].data_ptr(), t["inc_cell"].data_ptr(), cap_l, cap_c, cap_i,
                               t["out_off"].data_ptr() if fwd else 0, t["out_cell"].data_ptr() if fwd else 0, cap_i if fwd else 0)

    def _add_traffic_lights(self, check=True):   # city_model.py:1422
        lk = self._links_struct()
        _lib.check(self.lib.tsim_layout_lights(C.byref(self.cfg), C.byref(self._planes), C.byref(lk), self._flag_ptr(0), _ptr(self.workspace),
                                               C.c_size_t(self.workspace.numel()), self._stream))
        if check:
            self._check_flag("_add_traffic_lights")

    def lights_undecided(self):
        """Candidates whose `leads_to` queries the search stages of the last `_add_traffic_lights` left undecided: after the Z
        witnesses (they went to the window closures) and after those (they made the pass close the reachability planes)."""
        sc = self.workspace[: 64 * 4].view(torch.int32)[13:15].cpu().numpy()   # scalars at the head of the lights workspace
        return {"after_z_witness": int(sc[0]), "after_window": int(sc[1])}

    def _build_simple_maps(self):   # city_model.py:2151
        n = self.width * self.win_rows
        if self.maps is None:
            self.maps = {k: torch.empty(n, dtype=torch.uint8, device=self.device)
                         for k in ("is_road_map", "road_type_map", "intersection_map", "allowed_dirs_map")}
        m = self.maps
        _lib.check(self.lib.tsim_maps(C.byref(self.cfg), C.byref(self._planes), _ptr(m["is_road_map"]), _ptr(m["road_type_map"]),
                                      _ptr(m["intersection_map"]), _ptr(m["allowed_dirs_map"]), self._stream))

    # ------------------------------------------------------------------ whole pipeline
    def generate(self, tape_zone, tape_carve=None, tape_entrance=None, check=True, lights=True, maps=True):
        """All layout passes in the reference's order (city_model.py:125-139, 148)."""
        self._place_thick_wall(); self._place_sidewalk_inner_ring(); self._clear_interior()
        self._build_roads_and_sidewalks()
        if self.carve_subblock_roads:
            self._carve_subblock_roads(tape_carve, check=check)
        self._flood_fill_blocks_storing_data(tape_zone, check=check)
        self._eliminate_dead_ends()
        self._upgrade_r2_to_intersections(check=check)
        self._final_place_block_entrances(tape_entrance, check=check)
        self._remove_invalid_intersection_directions()
        self._add_entrance_directions()
        if lights:
            self._add_traffic_lights(check=check)
        if maps:
            self._build_simple_maps()

    def capture(self, fn):
        """CUDA graph of one call of ``fn()`` (any sequence of this object's passes with device-resident tapes and ``check=False``):
        every launch of the pipeline is static -- grids, pointers, the cooperative kernels included -- so a city is ONE graph launch
        instead of ~90 kernel launches.  Returns (graph, launches): ``graph.replay()`` regenerates the city from whatever the tape
        tensors hold at that moment.  ``fn`` is run once eagerly first (allocations, lazy initialisation happen outside the capture)."""
        fn()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        n0 = self.lib.tsim_launch_count()
        with torch.cuda.graph(g):
            fn()
        return g, int(self.lib.tsim_launch_count() - n0)

    # ------------------------------------------------------------------ read-back
    def planes_host(self):
        H, W = self.win_rows, self.width
        return {"cell_type": self.cell_type.cpu().numpy().reshape(H, W),
                "dirs": self.dirs.cpu().numpy().view(np.uint16).reshape(H, W),
                "aux": self.aux.cpu().numpy().reshape(H, W),
                "block_id": self.block_id.cpu().numpy().reshape(H, W)}

    def maps_host(self):
        H, W = self.win_rows, self.width
        return {k: v.cpu().numpy().reshape(H, W) for k, v in self.maps.items()}

    def light_links_host(self):
        """Link tables as sorted (light, cell) pair multisets, the form the oracle / fixtures use."""
        t = self._link_tensors
        n = int(self.flags[3].item())
        lights = t["light_cell"][:n].cpu().numpy()
        out = {"lights": lights}
        tables = [("ctrl", "ctrl_off", "ctrl_cell"), ("incoming", "inc_off", "inc_cell")]
        if "out_off" in t:
            tables.append(("outgoing", "out_off", "out_cell"))
        for name, off, cell in tables:
            o = t[off][: n + 1].cpu().numpy()
            c = t[cell][: int(o[-1])].cpu().numpy() if n else np.zeros(0, np.int32)
            owner = np.repeat(lights, np.diff(o)) if n else np.zeros(0, np.int32)
            pairs = np.stack([owner, c], 1).astype(np.int32) if n else np.zeros((0, 2), np.int32)
            order = np.lexsort((pairs[:, 1], pairs[:, 0]))
            out[name] = pairs[order]
        return out

    def sweeps(self):
        return int(self.flags[1].item())
