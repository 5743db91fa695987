"""Host side of the tick: vehicle SoA, light-group tables and tapes on the device, advanced by
``tsim_tick_run`` (one persistent cooperative CUDA kernel for any number of ticks).

Mirrors the reference's ``CityModel.step()`` (city_model.py:1831-1860) for the parts in scope: the
public maps ``occupancy_map``, ``stop_map``, ``stuck_map`` (city_model.py:109-115) live here as device
planes; vehicles are rows of a structure of arrays indexed by spawn-attempt number.

Tape contract: see DESIGN.md §5 and oracle/refharness/ticks.py (activation order = light groups, then
vehicles by ``rank``, then the spawner; route events replay every path the planner produced).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .light_groups import build_light_tables, pressure_cells

_I8 = ("alive", "base_speed", "cur_speed", "max_steps", "early", "is_stuck", "prev_valid", "malfunction", "direction", "moved")
_I32 = ("pos", "path_len", "steps", "stranded")
_G32 = ("g_cur", "g_pend", "g_qt", "g_gap", "g_last", "g_ft_phase", "g_ft_timer", "g_plan")
_LT = ("tl_off", "tl_cells", "g_all_off", "g_all", "g_ns_off", "g_ns", "g_ew_off", "g_ew",
       "g_nsin_off", "g_nsin", "g_ewin_off", "g_ewin", "g_cl_off", "g_cl")
_LT_OUT = ("g_nsout_off", "g_nsout", "g_ewout_off", "g_ewout")   # PRESSURE_CONTROL only
ALGOS = {"QUEUE_ACTUATED": 0, "FIXED_TIME": 1, "PRESSURE_CONTROL": 2, "NEIGHBOR_GREEN_WAVE": 3}   # Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM (config.py:341)


def light_tables_from_layout(city, links=False, creation_order=None):
    """Build the light-group tables for a generated ``GpuCityLayout`` (device labelling + host table work).
    links: also the neighbour-link table ``g_nbr`` NEIGHBOR_GREEN_WAVE reads (``light_groups.neighbor_links``: host-side ray marches,
    reference-sized cities); creation_order: the reference's group creation order if it is known (see there)."""
    W, H = city.width, city.height
    dev = city.device
    mask = ((city.aux & 0x40) != 0).to(torch.uint8)
    labels = torch.empty(W * H, dtype=torch.int32, device=dev)
    cap = max(1024, (W * H) // 16)
    blobs = torch.zeros(cap * 6, dtype=torch.int32, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    bl = _lib.Blobs(blobs.data_ptr(), cap, n.data_ptr(), 0)
    _lib.check(city.lib.tsim_label_mask(C.byref(city.cfg), C.c_void_p(mask.data_ptr()), C.c_void_p(labels.data_ptr()), C.byref(bl),
                                        city._flag_ptr(0), C.c_void_p(city.workspace.data_ptr()), C.c_size_t(city.workspace.numel()),
                                        city._stream))
    nc = int(n.item())
    city._check_flag("tsim_label_mask")
    if nc > cap:
        raise _lib.TsimError(6, f"{nc} intersection clusters exceed the table capacity {cap}")
    t = city._link_tensors
    nl = int(city.flags[3].item())
    planes = city.planes_host()
    ctrl_off = t["ctrl_off"][: nl + 1].cpu().numpy()
    inc_off = t["inc_off"][: nl + 1].cpu().numpy()
    out_off = t["out_off"][: nl + 1].cpu().numpy() if "out_off" in t else None
    tabs = build_light_tables(
        W, H, planes["cell_type"], planes["dirs"], labels.cpu().numpy().reshape(H, W), blobs[: nc * 6].view(nc, 6).cpu().numpy(),
        t["light_cell"][:nl].cpu().numpy(), ctrl_off, t["ctrl_cell"][: int(ctrl_off[-1]) if nl else 0].cpu().numpy(),
        inc_off, t["inc_cell"][: int(inc_off[-1]) if nl else 0].cpu().numpy(),
        out_off, t["out_cell"][: int(out_off[-1]) if nl else 0].cpu().numpy() if out_off is not None else None)
    if links:
        from .light_groups import neighbor_links
        lk = neighbor_links(W, H, planes["cell_type"], labels.cpu().numpy().reshape(H, W), tabs, t["light_cell"][:nl].cpu().numpy(),
                            city.hbands, city.vbands, creation_order=creation_order)
        tabs["g_nbr"], tabs["links_order_dependent"] = lk["nbr"], lk["order_dependent"]
    return tabs


class GpuTraffic:
    """Vehicle CA + traffic lights on the device.

    tapes: dict with spawn_tick (sorted), origin, target, speed [T,V] u8, malfunction [T,V] u8, rank [T,V] i32,
    ev_tick (sorted), ev_vehicle, ev_off (int64, n_events+1), ev_cells, optional rain_map [H,W] u8.
    """

    def __init__(self, width, height, light_tables, tapes, n_ticks, algo="QUEUE_ACTUATED", rain_enabled=False, device="cuda:0",
                 window=None, own_rows=None, live_list=None, route_capacity=None):
        """route_capacity: number of route cells the event buffer holds; the tapes then carry NO route events (`ev_*` ignored) and the
        caller hands every tick's routes over with ``push_route_events`` before it runs the tick (``replan.PlannedTraffic``).
        window = (win_y0, win_rows, win_halo): this object is one row-band shard (``ShardedTraffic`` builds those);
        every cell index in `light_tables` / `tapes` is then local to the window, cells outside it are -2;
        own_rows = (lo, hi): local rows the shard owns (the update counter skips the ghosts on the other rows).
        live_list: run the live-list kernel (k_tick2.cu: compacted vehicle records, one probe word per cell); default: yes
        on a whole city, no on a shard window (the halo exchange works on the vehicle-indexed arrays)."""
        if not torch.cuda.is_available():
            raise RuntimeError("trafficsimulation_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = dev = torch.device(device)
        self.W, self.H, self.n_ticks = int(width), int(height), int(n_ticks)
        if algo not in ALGOS:
            raise ValueError(f"light controller {algo!r}: the device runs {sorted(ALGOS)}")
        self.algo = ALGOS[algo]
        if self.algo >= 2 and window is not None:
            raise NotImplementedError(f"{algo} on a row-band shard: the cells / neighbour groups that controller reads lie outside the shard's window")
        self.win_y0, self.win_rows, self.win_halo = (0, self.H, 0) if window is None else (int(v) for v in window)
        self.cfg = _lib.Cfg(self.W, self.H, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, self.win_y0, self.win_rows, self.win_halo)
        n = self.W * self.win_rows

        def up(a, dt):   # host array -> device; a tensor already on the device is shared, not copied
            if isinstance(a, torch.Tensor):
                return a.to(device=dev, dtype=getattr(torch, np.dtype(dt).name)).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
        # ---- light tables
        if self.algo == 2:
            if any(k not in light_tables for k in _LT_OUT):
                raise ValueError("PRESSURE_CONTROL needs the g_nsout / g_ewout lane tables (light_groups.build_light_tables)")
            light_tables = pressure_cells(light_tables, self.W)   # the cells the reference's controller really reads
        self.lt_t = {k: up(light_tables[k], np.int32) for k in _LT + (_LT_OUT if self.algo == 2 else ())}
        self.n_groups, self.n_lights = int(light_tables["n_groups"]), int(light_tables["n_lights"])
        if self.algo == 3:
            if "g_nbr" not in light_tables:
                raise ValueError("NEIGHBOR_GREEN_WAVE needs the g_nbr link table (light_tables_from_layout(city, links=True))")
            self.lt_t["g_nbr"] = up(np.asarray(light_tables["g_nbr"]).reshape(-1), np.int32)
        self.lt = _lib.LightTables(self.n_groups, self.n_lights, *[self.lt_t[k].data_ptr() for k in _LT],
                                   *[(self.lt_t[k].data_ptr() if self.algo == 2 else 0) for k in _LT_OUT],
                                   self.lt_t["g_nbr"].data_ptr() if self.algo == 3 else 0)
        # ---- tapes
        spawn_tick = np.asarray(tapes["spawn_tick"], np.int32)
        self.route_capacity = None if route_capacity is None else int(route_capacity)
        if self.route_capacity is not None:
            if window is not None:
                raise NotImplementedError("pushed route events on a row-band shard")
            ne = len(spawn_tick)   # at most one event per vehicle and tick; the cell buffer is append-only (live routes point into it)
            tapes = dict(tapes, ev_tick=np.zeros(0, np.int32), ev_vehicle=np.zeros(0, np.int32), ev_off=np.zeros(1, np.int64),
                         ev_cells=np.zeros(0, np.int32))
        ev_tick = np.asarray(tapes["ev_tick"], np.int32)
        if np.any(np.diff(spawn_tick) < 0) or np.any(np.diff(ev_tick) < 0):
            raise ValueError("spawn attempts and route events must be sorted by tick")
        self.nv = nv = len(spawn_tick)
        self._validate_tapes(tapes, n_ticks, nv, len(ev_tick), n)
        tt = {}
        tt["spawn_first"] = up(np.searchsorted(spawn_tick, np.arange(n_ticks + 1)), np.int32)
        tt["origin"], tt["target"] = up(tapes["origin"], np.int32), up(tapes["target"], np.int32)
        tt["speed"], tt["malfunction"] = up(tapes["speed"], np.uint8), up(tapes["malfunction"], np.uint8)
        tt["rank"] = up(tapes["rank"], np.int32)
        tt["ev_first"] = up(np.searchsorted(ev_tick, np.arange(n_ticks + 1)), np.int32)
        tt["ev_vehicle"] = up(tapes["ev_vehicle"], np.int32)
        tt["ev_off"] = up(tapes["ev_off"], np.int64)
        ev = tapes["ev_cells"]   # one spare entry at the end: the kernel may form the address of path[len]
        tt["ev_cells"] = up(ev, np.int32) if isinstance(ev, torch.Tensor) else up(np.append(np.asarray(ev, np.int32), 0), np.int32)
        if self.route_capacity is not None:   # room for the pushed events: every tick's events are written at the front of ev_vehicle / ev_off
            tt["ev_first"] = torch.zeros(n_ticks + 2, dtype=torch.int32, device=dev)
            tt["ev_vehicle"] = torch.zeros(max(ne, 1), dtype=torch.int32, device=dev)
            tt["ev_off"] = torch.zeros(max(ne, 1) + 1, dtype=torch.int64, device=dev)
            tt["ev_cells"] = torch.zeros(self.route_capacity + 1, dtype=torch.int32, device=dev)
            self._route_used = 0
        rain = tapes.get("rain_map") if rain_enabled else None
        tt["rain_map"] = up(np.asarray(rain).reshape(-1), np.uint8) if rain is not None else None
        self.tt = tt
        self.tp = _lib.TickTapes(n_ticks, nv, *[(tt[k].data_ptr() if tt[k] is not None else 0) for k in
                                                ("spawn_first", "origin", "target", "speed", "malfunction", "rank", "ev_first", "ev_vehicle",
                                                 "ev_off", "ev_cells", "rain_map")])
        # ---- state
        z = lambda cnt, dt: torch.zeros(max(int(cnt), 1), dtype=dt, device=dev)
        s = {"occupancy": z(n, torch.uint8), "stop_map": z(n, torch.uint8), "stuck_map": z(n, torch.uint8),
             "claim": z(2 * n, torch.int64), "stopw": z(n, torch.int32)}
        for k in _I32:
            s[k] = z(nv, torch.int32)
        s["path_off"] = z(nv, torch.int64)
        s["stuck_ticks"] = z(nv, torch.int16)
        for k in _I8:
            s[k] = z(nv, torch.int8)
        self.gstate = torch.zeros(len(_G32), max(self.n_groups, 1), dtype=torch.int32, device=dev)   # one row per field: shards exchange columns
        for i, k in enumerate(_G32):
            s[k] = self.gstate[i]
        s["scalars"] = z(16, torch.int32)
        self.live_list = (window is None and own_rows is None) if live_list is None else bool(live_list)
        if self.live_list:
            nb = C.c_longlong(0)
            _lib.check(self.lib.tsim_tick_probe_bytes(C.byref(self.cfg), C.byref(nb)))
            s["probe"] = z(nb.value // 8, torch.int64)   # bit planes over 8 x 8-cell tiles
            nt = C.c_int32(0)
            _lib.check(self.lib.tsim_tick_tiles(C.byref(self.cfg), C.byref(nt)))
            s["recs"] = torch.empty(max(3 * nv * 48, 16), dtype=torch.uint8, device=dev)   # two halves of the live list + the sort's staging copy
            s["sort_keys"], s["tile_ws"] = z(2 * nv, torch.int32), z(2 * nt.value, torch.int32)
            s["plans"] = torch.empty(max(nv * 32, 16), dtype=torch.uint8, device=dev)
            s["ev_stamp"], s["ev_plen"], s["ev_poff"] = z(nv, torch.int32), z(nv, torch.int32), z(nv, torch.int64)
            nb = C.c_longlong(0)
            _lib.check(self.lib.tsim_tick_group_ws_bytes(C.byref(self.cfg), C.byref(self.lt), C.byref(nb)))
            s["group_ws"] = z((nb.value + 7) // 8, torch.int64)   # occupancy bit tiles + (tile, mask) lists of the light groups
        if self.algo == 3:
            s["g_wave"] = z(3 * self.n_groups + 4, torch.int32)   # scratch of the green-wave fixed point
        if window is not None and not self.live_list:
            s["live_idx"] = z(nv, torch.int32)   # a shard iterates the vehicles of its own window (rebuilt after every halo refresh)
        self.s = s
        v1 = [f[0] for f in _lib.TickState._fields_ if f[1] is C.c_void_p][:30]
        v2 = ("probe", "recs", "plans", "ev_stamp", "ev_plen", "ev_poff")
        self.st = _lib.TickState(*[s[k].data_ptr() for k in v1], *(own_rows or (0, 0)), *[(s[k].data_ptr() if self.live_list else 0) for k in v2],
                                 s["live_idx"].data_ptr() if "live_idx" in s else 0,
                                 s["sort_keys"].data_ptr() if self.live_list else 0, s["tile_ws"].data_ptr() if self.live_list else 0,
                                 s["group_ws"].data_ptr() if self.live_list else 0, s["g_wave"].data_ptr() if "g_wave" in s else 0)
        _lib.check(self.lib.tsim_tick_init(C.byref(self.cfg), C.byref(self.lt), C.byref(self.tp), C.byref(self.st), self._stream))
        self._ticks_run, self._exported_at = 0, -1

    @staticmethod
    def _validate_tapes(tapes, n_ticks, nv, n_events, n_cells):
        """The kernels index the tapes without bounds checks: everything they can reach is checked here, once."""
        size = lambda a: a.numel() if isinstance(a, torch.Tensor) else np.asarray(a).size
        for k in ("speed", "malfunction", "rank"):
            if size(tapes[k]) < n_ticks * nv:
                raise ValueError(f"tape '{k}' holds {size(tapes[k])} entries, {n_ticks} ticks x {nv} vehicles need {n_ticks * nv}")
        for k in ("origin", "target"):
            if size(tapes[k]) != nv:
                raise ValueError(f"tape '{k}' must have one entry per spawn attempt ({nv})")
        if size(tapes["ev_vehicle"]) != n_events or size(tapes["ev_off"]) != n_events + 1:
            raise ValueError("ev_vehicle needs one entry per route event and ev_off one more")
        ev_off, ev_cells = tapes["ev_off"], tapes["ev_cells"]
        if n_events:
            last = int(ev_off[-1]) if not isinstance(ev_off, torch.Tensor) else int(ev_off[-1].item())
            if last > size(ev_cells):
                raise ValueError(f"ev_off ends at {last}, ev_cells holds {size(ev_cells)} cells")
        for k in ("origin", "target", "ev_cells", "ev_vehicle"):
            a = tapes[k]
            if size(a) == 0:
                continue
            lo, hi = (int(a.min().item()), int(a.max().item())) if isinstance(a, torch.Tensor) else (int(np.min(a)), int(np.max(a)))
            top = nv if k == "ev_vehicle" else n_cells
            floor = 0 if k == "ev_vehicle" else _lib.CELL_OUTSIDE
            if lo < floor or hi >= top or (k != "ev_vehicle" and lo == -1):
                raise ValueError(f"tape '{k}' holds values outside [{floor}, {top})")

    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def route_room(self):
        """Free cells of the route buffer (routes are appended: a live vehicle's route stays where it was written)."""
        return self.route_capacity - self._route_used

    def push_route_events(self, vehicles, paths, compact=False):
        """Routes for the NEXT tick to run: ``vehicles[i]`` follows ``paths[i]`` (cell indices, first step first; may be empty) from
        phase A of that tick on -- a live vehicle's re-plan, or the first route of a vehicle that spawned in the tick before.
        compact: the caller hands over the remaining route of EVERY live vehicle, so the buffer starts again at its front."""
        if self.route_capacity is None:
            raise RuntimeError("push_route_events needs GpuTraffic(route_capacity=...)")
        if compact:
            self._route_used = 0
        t = self._ticks_run
        if t >= self.n_ticks:
            raise ValueError(f"tick {t} is past the end of the tapes ({self.n_ticks} ticks)")
        n = len(vehicles)
        if n != len(paths) or n > self.nv or len(set(int(v) for v in vehicles)) != n:
            raise ValueError("one route per vehicle and tick")
        lens = np.array([len(p) for p in paths], np.int64)
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum(lens)
        total = int(off[-1])
        if self._route_used + total > self.route_capacity:
            raise _lib.TsimError(6, f"route buffer full: {self._route_used} + {total} cells > route_capacity {self.route_capacity}")
        if n:
            veh = np.asarray(vehicles, np.int32)
            cells = np.concatenate([np.asarray(p, np.int32) for p in paths]) if total else np.zeros(0, np.int32)
            if veh.min() < 0 or veh.max() >= self.nv or (total and (cells.min() < 0 or cells.max() >= self.W * self.win_rows)):
                raise ValueError("route event outside the fleet / the grid")
            self.tt["ev_vehicle"][:n] = torch.from_numpy(veh).to(self.device)
            self.tt["ev_off"][: n + 1] = torch.from_numpy(off + self._route_used).to(self.device)
            if total:
                self.tt["ev_cells"][self._route_used: self._route_used + total] = torch.from_numpy(cells).to(self.device)
            self._route_used += total
        first = self.tt["ev_first"]
        first[t] = 0
        first[t + 1:] = n        # the events of tick t are entries [0, n); nothing is queued for the ticks after it yet

    def plan_snapshot(self):
        """What the route planner reads between two ticks (host arrays): the public maps and, per vehicle, alive / pos / path_len /
        stuck_ticks / stranded (ticks left) / malfunction_flag / collision_flag / base_speed / cur_speed / is_stuck / direction."""
        self.export()
        s = self.s
        g = lambda k: s[k][: max(self.nv, 1)].cpu().numpy()[: self.nv]
        return dict(occupancy=s["occupancy"].cpu().numpy(), stop_map=s["stop_map"].cpu().numpy(), alive=g("alive") == 1, pos=g("pos"),
                    path_len=g("path_len"), stuck_ticks=g("stuck_ticks"), stranded=g("stranded"), malfunction_flag=(g("malfunction") & 1) != 0,
                    collision_flag=(g("malfunction") & 2) != 0, base_speed=g("base_speed"), cur_speed=g("cur_speed"), is_stuck=g("is_stuck"),
                    direction=g("direction"))

    def export(self):
        """Live-list kernel: bring the public maps and the vehicle SoA up to date (a tick itself only moves records and probe bytes)."""
        if self.live_list and self._exported_at != self._ticks_run:
            _lib.check(self.lib.tsim_tick_export(C.byref(self.cfg), C.byref(self.tp), C.byref(self.st), self._stream))
            self._exported_at = self._ticks_run

    # public maps of the reference model (city_model.py:109-115)
    @property
    def occupancy_map(self):
        self.export()
        return self.s["occupancy"].view(self.win_rows, self.W)

    @property
    def stop_map(self):
        self.export()
        return self.s["stop_map"].view(self.win_rows, self.W)

    @property
    def stuck_map(self):
        self.export()
        return self.s["stuck_map"].view(self.win_rows, self.W)

    def step(self, n=1, check=True):
        """Advance n ticks (CityModel.step, city_model.py:1831)."""
        _lib.check(self.lib.tsim_tick_run(C.byref(self.cfg), C.byref(self.lt), C.byref(self.tp), C.byref(self.st), int(n), self.algo, self._stream))
        self._ticks_run += int(n)
        if check:
            err = int(self.s["scalars"][1].item())
            if err:
                why = {30: "a live vehicle stands on its target in phase A (tape contract)", 31: "ran past the end of the tapes",
                       32: "the claim fixed point did not settle", 33: "a tape speed above 5",
                       34: "a firing sideswipe draw (tape bit 1) needs the live-list kernel (live_list=True)",
                       35: "an occupied cell without a live vehicle on it"}.get(err, "")
                raise _lib.TsimError(4 if err in (30, 33, 34) else 6, f"tick kernel error flag {err}: {why}")

    @property
    def tick(self):
        return int(self.s["scalars"][0].item())

    def counters(self):
        sc = self.s["scalars"].cpu().numpy()
        return {"tick": int(sc[0]), "fixed_point_iterations": int(sc[2]), "vehicle_updates": int(sc[6:8].view(np.int64)[0])}

    def state_host(self):
        """Same dict as oracle.OracleTicks.state() / the reference fixtures."""
        self.export()   # live list: the maps and the vehicle SoA are only written on demand
        s = {k: v.cpu().numpy() for k, v in self.s.items() if k not in ("claim", "stopw", "probe", "recs", "plans", "ev_stamp", "ev_plen", "ev_poff", "live_idx", "sort_keys", "tile_ws", "group_ws")}
        alive = s["alive"][: self.nv] == 1
        cut = lambda a: a[: self.nv]
        # bit 0 is_stuck, bit 1 is_in_malfunction, bits 2-4 direction + 1, bit 5 is_in_collision (the live-list kernel exports the
        # collision flag as bit 1 of the malfunction byte)
        malf = cut(s["malfunction"]).astype(np.uint8)
        flags = (cut(s["is_stuck"]).astype(np.uint8) & 1) | ((malf & 1) << 1) | ((cut(s["direction"]) + 1).astype(np.uint8) << 2) | (((malf >> 1) & 1) << 5)
        ng = self.n_groups
        return dict(pos=np.where(alive, cut(s["pos"]), -1), base_speed=np.where(alive, cut(s["base_speed"]), 0),
                    stuck_ticks=np.where(alive, cut(s["stuck_ticks"]), 0), vflags=np.where(alive, flags, 0).astype(np.uint8),
                    occ=np.flatnonzero(s["occupancy"]).astype(np.int32), stop=np.flatnonzero(s["stop_map"]).astype(np.int32),
                    stuckmap=np.flatnonzero(s["stuck_map"]).astype(np.int32),
                    groups=np.stack([s["g_cur"][:ng], s["g_pend"][:ng], s["g_qt"][:ng], s["g_gap"][:ng], s["g_last"][:ng]], 1))
