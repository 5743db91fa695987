"""Row-band shards of the tick across the GPUs of one box (SURVEY.md §8e "Vehicle step").

Shard ``s`` owns the vehicles, light groups and map rows of its own row band and simulates a WINDOW = own rows
plus ``halo`` rows of each neighbour with the single-GPU tick kernel (``GpuTraffic`` on a window).  Vehicles and
groups on the halo rows are GHOSTS: copies of what the neighbour owns, simulated redundantly so that a tick
needs no communication while it runs.  A vehicle moves at most 5 cells and looks at most 5 cells ahead per tick,
a light group reads and writes cells of its own intersection only, so whatever happens on a shard's own rows
depends on state a few rows away -- except through chains of vehicles that block one another, which are short
but not bounded a priori.  After EVERY tick the owners refresh the neighbours' halos:

* the three map planes (occupancy, stop, stuck) row-wise,
* the vehicles on the ``halo`` own rows next to a cut as fixed-size records,
* the state of the light groups both shards simulate,

all in ONE message per neighbour, built by one kernel (``tsim_tick_pack``) and installed by another (``tsim_tick_unpack``).

Exactness is checked, not assumed: on the half of a halo next to the cut the receiving shard compares its own
ghost simulation with what the owner sent (rows, vehicle records, group state).  An error that starts at the
window edge can only reach a shard's own rows through vehicles / groups whose state is wrong on the way, one
link (<= ``group extent + 10`` rows, asserted against ``halo // 2``) at a time, so it would show up there first and
raises ``TSIM_ERR_CAPACITY`` ("halo too small") instead of a silently different city.

Every shard keeps the vehicle arrays at their global length (a vehicle is the same index everywhere; 10 M
vehicles are ~0.4 GB) and the replicated tapes; cell indices are translated to the window on the way in.
Deployments: one process per GPU under ``torch.distributed`` (NCCL), or one process holding all shards (how
the ``-m gpu`` tests check N shards == 1 shard == the oracle on a single GPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .sharded import Comm, ShardPlan
from .traffic import GpuTraffic

_GROUP_LISTS = ("g_all", "g_ns", "g_ew", "g_nsin", "g_ewin", "g_cl")
MAX_LINK_ROWS = 10   # a vehicle reaches 5 rows; the vehicle that blocks it started at most 5 rows from the contested cell


def group_row_extents(tabs, W):
    """Per light group: lowest / highest grid row of any cell it reads or writes, and the row of its first cluster cell."""
    ng = int(tabs["n_groups"])
    lo = np.full(ng, np.iinfo(np.int64).max, np.int64)
    hi = np.full(ng, -1, np.int64)
    tl_off, tl_cells = np.asarray(tabs["tl_off"], np.int64), np.asarray(tabs["tl_cells"], np.int64)
    nl = len(tl_off) - 1
    l_lo = np.full(nl, np.iinfo(np.int64).max, np.int64)
    l_hi = np.full(nl, -1, np.int64)
    if nl and len(tl_cells):
        rows = tl_cells // W
        owner = np.repeat(np.arange(nl), np.diff(tl_off))
        np.minimum.at(l_lo, owner, rows)
        np.maximum.at(l_hi, owner, rows)
    for k in ("g_all", "g_nsin", "g_ewin", "g_cl"):
        off, val = np.asarray(tabs[k + "_off"], np.int64), np.asarray(tabs[k], np.int64)
        owner = np.repeat(np.arange(ng), np.diff(off))
        if k == "g_all":
            np.minimum.at(lo, owner, l_lo[val])
            np.maximum.at(hi, owner, l_hi[val])
        else:
            np.minimum.at(lo, owner, val // W)
            np.maximum.at(hi, owner, val // W)
    cl_off = np.asarray(tabs["g_cl_off"], np.int64)
    first = np.asarray(tabs["g_cl"], np.int64)[cl_off[:-1]] // W if ng else np.zeros(0, np.int64)
    return lo, hi, first


def _csr_take(off, val, sel):
    """rows `sel` (bool mask) of a CSR table"""
    off = np.asarray(off, np.int64)
    n = np.diff(off)[sel]
    start = off[:-1][sel]
    new_off = np.zeros(len(n) + 1, np.int64)
    new_off[1:] = np.cumsum(n)
    idx = np.repeat(start - new_off[:-1], n) + np.arange(int(new_off[-1]))
    return new_off.astype(np.int32), np.asarray(val)[idx]


def shard_light_tables(tabs, W, win_y0, win_rows, sel):
    """The groups `sel` (all of whose cells lie inside the window) with window-local cell indices.  The light table
    keeps every light (groups refer to lights by index); cells of lights outside the window become -2."""
    base, n = win_y0 * W, win_rows * W

    def local(c):
        c = np.asarray(c, np.int64) - base
        return np.where((c >= 0) & (c < n), c, _lib.CELL_OUTSIDE).astype(np.int32)

    out = {"tl_off": np.asarray(tabs["tl_off"], np.int32), "tl_cells": local(tabs["tl_cells"]), "n_lights": int(tabs["n_lights"]),
           "n_groups": int(sel.sum())}
    for k in _GROUP_LISTS:
        off, val = _csr_take(tabs[k + "_off"], tabs[k], sel)
        out[k + "_off"] = off
        out[k] = val.astype(np.int32) if k in ("g_all", "g_ns", "g_ew") else local(val)
        if k not in ("g_all", "g_ns", "g_ew") and len(val) and (out[k] < 0).any():
            raise AssertionError("a selected light group has a cell outside the window")
    return out


class ShardedTraffic:
    """``GpuTraffic`` over ``n_shards`` row bands.  Same tapes, same light tables (global cell indices), same results."""

    def __init__(self, width, height, light_tables, tapes, n_ticks, n_shards, halo=128, algo="QUEUE_ACTUATED", rain_enabled=False,
                 devices=None, distributed=False, group=None):
        self.W, self.H, self.n_ticks = int(width), int(height), int(n_ticks)
        self.plan = plan = ShardPlan(self.H, n_shards, halo if n_shards > 1 else 0)
        self.comm = Comm(n_shards, distributed, group)
        self.n, self.halo = n_shards, plan.halo
        W = self.W
        tabs = light_tables
        self.n_groups_global = ng = int(tabs["n_groups"])
        g_lo, g_hi, g_first = group_row_extents(tabs, W)
        extent = int((g_hi - g_lo + 1).max()) if ng else 0
        if n_shards > 1 and self.halo // 2 < extent + MAX_LINK_ROWS:
            raise ValueError(f"halo {self.halo} too small: needs 2 * (tallest light group {extent} rows + {MAX_LINK_ROWS})")
        own_hi = np.asarray(plan.own_hi)
        self.g_owner = np.searchsorted(own_hi, g_first, side="right")
        self.g_sel = [(g_lo >= plan.win_lo[r]) & (g_hi < plan.win_hi[r]) for r in range(n_shards)]
        for r in range(n_shards):
            if not self.g_sel[r][self.g_owner == r].all():
                raise AssertionError("a light group is not inside its owner's window")
        self.g_local = [np.cumsum(m) - 1 for m in self.g_sel]        # global group -> index in shard r's tables
        nv = len(tapes["origin"])
        self.nv = nv
        cap = 0
        if n_shards > 1:   # a strip of `halo` rows cannot hold more vehicles than road cells; 1/2 of its cells is generous
            cap = int(min(nv, self.halo * W // 2)) + 16
        self.cap = cap
        self.sims, self.strips, self.send, self.recv = {}, {}, {}, {}
        self._keep = []   # device index lists the strips point into
        if devices is None:
            devices = ["cuda:0"] * n_shards
        shared = {}
        for s in self.comm.local:
            dev = torch.device(devices[s] if not distributed else devices[0] if len(devices) == 1 else devices[s])
            y0, rows = plan.win_lo[s], plan.win_hi[s] - plan.win_lo[s]
            base, ncell = y0 * W, rows * W
            lt = shard_light_tables(tabs, W, y0, rows, self.g_sel[s])
            key = str(dev)
            if key not in shared:   # replicated tapes: one device copy per device
                shared[key] = {k: torch.from_numpy(np.ascontiguousarray(tapes[k])).to(dev) for k in ("speed", "malfunction", "rank")}
                shared[key]["cells"] = {k: torch.from_numpy(np.ascontiguousarray(np.asarray(tapes[k]))).to(dev) for k in ("origin", "target")}
                shared[key]["cells"]["ev_cells"] = torch.from_numpy(np.append(np.asarray(tapes["ev_cells"]), 0)).to(dev)
            tl = dict(tapes)
            tl.update({k: shared[key][k] for k in ("speed", "malfunction", "rank")})
            for k, g in shared[key]["cells"].items():
                c = g.to(torch.int64) - base
                tl[k] = torch.where((c >= 0) & (c < ncell), c, torch.full_like(c, _lib.CELL_OUTSIDE)).to(torch.int32)
            if rain_enabled and tapes.get("rain_map") is not None:
                tl["rain_map"] = np.asarray(tapes["rain_map"]).reshape(self.H, W)[y0:y0 + rows]
            sim = GpuTraffic(W, self.H, lt, tl, n_ticks, algo=algo, rain_enabled=rain_enabled, device=dev,
                             window=(y0, rows, self.halo), own_rows=(plan.own_lo[s] - y0, plan.own_hi[s] - y0))
            self.sims[s] = sim
            if n_shards == 1:
                continue
            st = _lib.TickStrips()
            h2 = self.halo // 2
            own0, own1 = plan.own_lo[s] - y0, plan.own_hi[s] - y0
            # [0] neighbour below: my lowest `halo` own rows travel, my lower halo is refreshed, its upper half is verified
            st.send_lo[0], st.send_hi[0] = (own0, own0 + self.halo) if s > 0 else (0, 0)
            st.halo_lo[0], st.halo_hi[0] = (0, own0) if s > 0 else (0, 0)
            st.verify_lo[0], st.verify_hi[0] = (own0 - h2, own0) if s > 0 else (0, 0)
            st.send_lo[1], st.send_hi[1] = (own1 - self.halo, own1) if s + 1 < n_shards else (0, 0)
            st.halo_lo[1], st.halo_hi[1] = (own1, rows) if s + 1 < n_shards else (0, 0)
            st.verify_lo[1], st.verify_hi[1] = (own1, own1 + h2) if s + 1 < n_shards else (0, 0)
            st.cap = cap
            # light groups both neighbours simulate: the owner sends, the other installs (and verifies away from its window edge)
            v_lo = plan.win_lo[s] + (h2 if plan.win_lo[s] > 0 else 0)
            v_hi = plan.win_hi[s] - (h2 if plan.win_hi[s] < self.H else 0)
            self.send[s], self.recv[s], keep = [None, None], [None, None], []
            for d, r in ((0, s - 1), (1, s + 1)):
                if r < 0 or r >= n_shards:
                    continue
                both = self.g_sel[s] & self.g_sel[r]
                mine = np.flatnonzero(both & (self.g_owner == s))
                theirs = np.flatnonzero(both & (self.g_owner == r))
                g_send = torch.from_numpy(self.g_local[s][mine].astype(np.int32)).to(dev)
                g_recv = torch.from_numpy(self.g_local[s][theirs].astype(np.int32)).to(dev)
                g_ver = torch.from_numpy(((g_lo[theirs] >= v_lo) & (g_hi[theirs] < v_hi)).astype(np.uint8)).to(dev)
                keep += [g_send, g_recv, g_ver]
                st.g_send[d], st.n_g_send[d] = g_send.data_ptr(), len(mine)
                st.g_recv[d], st.g_verify[d], st.n_g_recv[d] = g_recv.data_ptr(), g_ver.data_ptr(), len(theirs)
                words = lambda ng_, rows_: int(sim.lib.tsim_tick_message_words(W, cap, ng_, rows_))
                self.send[s][d] = torch.zeros(words(len(mine), st.send_hi[d] - st.send_lo[d]), dtype=torch.int32, device=dev)
                self.recv[s][d] = torch.zeros(words(len(theirs), st.halo_hi[d] - st.halo_lo[d]), dtype=torch.int32, device=dev)
                st.send_msg[d], st.recv_msg[d] = self.send[s][d].data_ptr(), self.recv[s][d].data_ptr()
            self._keep += keep
            self.strips[s] = st

    # ---- exchange plumbing (Comm.exchange talks in global row ranges; the ranges identify the direction)
    def _role(self, s, lo, hi):
        p = self.plan
        if s + 1 < self.n and (lo, hi) == p.up_rows(s):
            return "send", 1
        if s > 0 and (lo, hi) == p.down_rows(s):
            return "send", 0
        if s + 1 < self.n and (lo, hi) == p.down_rows(s + 1):
            return "recv", 1
        if s > 0 and (lo, hi) == p.up_rows(s - 1):
            return "recv", 0
        raise AssertionError((s, lo, hi))

    def _message(self, s, lo, hi):
        """Comm.exchange item: shard s's message buffer for the neighbour the row range belongs to (sent / received in place)."""
        role, d = self._role(s, lo, hi)
        return (self.send if role == "send" else self.recv)[s][d]

    def _strip_call(self, fn, s):
        sim = self.sims[s]
        _lib.check(fn(C.byref(sim.cfg), C.byref(sim.tp), C.byref(sim.st), C.byref(self.strips[s]), sim._stream))

    def step(self, n=1, check=True):
        """Advance n ticks: one persistent launch per tick and shard, then the halo refresh."""
        for _ in range(n):
            for s, sim in self.sims.items():
                with torch.cuda.device(sim.device):
                    sim.step(1, check=False)
                    if self.n > 1:
                        self._strip_call(sim.lib.tsim_tick_pack, s)
            if self.n == 1:
                continue
            self.comm.exchange(self.plan, self._message)
            for s, sim in self.sims.items():
                with torch.cuda.device(sim.device):
                    self._strip_call(sim.lib.tsim_tick_unpack, s)
        if check:
            self.check()

    def check(self):
        """Raise if any shard's kernel flagged an error or a ghost diverged from its owner (host sync).  The diagnosis is made
        from the code reduced over ALL shards, so every rank raises the same error for the same reason."""
        codes, mine = {}, {}
        for s, sim in self.sims.items():
            sc = sim.s["scalars"][:10].cpu().numpy()
            mine[s] = (int(sc[1]), int(sc[9]))
            codes[s] = torch.tensor([max(mine[s])], dtype=torch.int32, device=sim.device)
        worst = self.comm.max(codes)
        if worst:
            halo = 40 <= worst <= 45
            raise _lib.TsimError(6 if halo or worst == 32 else 4,
                                 f"sharded tick: device error flag {worst} (kernel flag, exchange flag per local shard: {mine})"
                                 + (" -- a ghost diverged from its owner: the halo is too small for this traffic" if halo else ""))

    def counters(self):
        """Totals over the LOCAL shards (sum over ranks for the global figure)."""
        out = {"tick": 0, "fixed_point_iterations": 0, "vehicle_updates": 0}
        for sim in self.sims.values():
            c = sim.counters()
            out["tick"] = c["tick"]
            out["fixed_point_iterations"] = max(out["fixed_point_iterations"], c["fixed_point_iterations"])
            out["vehicle_updates"] += c["vehicle_updates"]
        return out

    # ---- results
    def state_parts(self):
        """Per local shard: what it OWNS, in global indices (vehicles on own rows, own map rows, own groups)."""
        parts = {}
        W = self.W
        for s, sim in self.sims.items():
            p, y0 = self.plan, sim.win_y0
            st = {k: v.cpu().numpy() for k, v in sim.s.items() if k not in ("claim", "stopw", "scalars", "live_idx", "sort_keys", "tile_ws", "probe", "recs", "plans", "ev_stamp", "ev_plen", "ev_poff")}
            row = st["pos"][: self.nv] // W + y0
            own = (st["alive"][: self.nv] == 1) & (row >= p.own_lo[s]) & (row < p.own_hi[s])
            ids = np.flatnonzero(own)
            flags = (st["is_stuck"][ids].astype(np.uint8) & 1) | ((st["malfunction"][ids].astype(np.uint8) & 1) << 1) | \
                    ((st["direction"][ids] + 1).astype(np.uint8) << 2)
            a, b = p.own_lo[s] - y0, p.own_hi[s] - y0
            cells = lambda k: (np.flatnonzero(st[k].reshape(sim.win_rows, W)[a:b]) + p.own_lo[s] * W).astype(np.int64)
            gids = np.flatnonzero(self.g_sel[s] & (self.g_owner == s))
            gl = self.g_local[s][gids]
            parts[s] = dict(ids=ids, pos=st["pos"][ids].astype(np.int64) + y0 * W, base_speed=st["base_speed"][ids], stuck_ticks=st["stuck_ticks"][ids],
                            vflags=flags, occ=cells("occupancy"), stop=cells("stop_map"), stuckmap=cells("stuck_map"), gids=gids,
                            groups=np.stack([st[k][gl] for k in ("g_cur", "g_pend", "g_qt", "g_gap", "g_last")], 1) if len(gl) else
                            np.zeros((0, 5), np.int32))
        return parts

    def state_host(self):
        """Same dict as ``GpuTraffic.state_host()`` / ``oracle.OracleTicks.state()``, merged over all shards."""
        parts = self.state_parts()
        if self.comm.dist is not None:
            box = [None] * self.n
            self.comm.dist.all_gather_object(box, parts, group=self.comm.group)
            parts = {k: v for d in box for k, v in d.items()}
        nv, ng = self.nv, self.n_groups_global
        out = dict(pos=np.full(nv, -1, np.int64), base_speed=np.zeros(nv, np.int8), stuck_ticks=np.zeros(nv, np.int16),
                   vflags=np.zeros(nv, np.uint8), groups=np.zeros((ng, 5), np.int32))
        seen = np.zeros(nv, np.int32)
        for s in sorted(parts):
            p = parts[s]
            for k in ("pos", "base_speed", "stuck_ticks", "vflags"):
                out[k][p["ids"]] = p[k]
            seen[p["ids"]] += 1
            out["groups"][p["gids"]] = p["groups"]
        if (seen > 1).any():
            raise AssertionError("a vehicle is owned by two shards")
        for k in ("occ", "stop", "stuckmap"):
            out[k] = np.concatenate([parts[s][k] for s in sorted(parts)])
        return out
