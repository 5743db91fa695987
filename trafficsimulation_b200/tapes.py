"""Synthetic layout tapes for sizes the reference cannot reach.

The tapes are INPUT tensors of the GPU path; for parity runs they are recorded from the reference
(oracle/refharness), for large synthetic cities they are drawn here with numpy generators.
"""
from __future__ import annotations

import random

import numpy as np

from .bands import BandParams, bands_to_array, make_city_bands

ZONE_WEIGHTS = [0.25, 0.25, 0.2, 0.2, 0.1]   # Defaults.CITY_BLOCK_CHANCE, config.py:53-60


def synth_bands(seed: int, **band_kwargs):
    """Band lists from the reference's own generator restated in bands.py, seeded privately."""
    rnd = random.Random(seed)
    hb, vb = make_city_bands(BandParams(**band_kwargs), rnd=rnd)
    return bands_to_array(hb), bands_to_array(vb)


def synth_zone_tape(seed: int, n: int) -> np.ndarray:
    return np.random.default_rng(seed).choice(5, size=n, p=ZONE_WEIGHTS).astype(np.uint8)


def synth_carve_tape(seed: int, blobs: np.ndarray, min_subblock_spacing: int = 5, subblock_chance: float = 0.3) -> np.ndarray:
    """One row per blob: (drawn, carved, px, py, hor_dir, ver_dir, inbound_is_horizontal, tries).

    Follows the acceptance rules of city_model.py:649-682: Bernoulli(subblock_chance); bbox at least
    2*min+1 in both axes; pivot uniform in the legal range (then the first try always passes the
    small-side test); directions and the inbound leg uniform.
    """
    rng = np.random.default_rng(seed)
    n = len(blobs)
    tape = np.zeros((n, 8), np.int32)
    if n == 0:
        return tape
    minx, miny, maxx, maxy = blobs[:, 0], blobs[:, 1], blobs[:, 2], blobs[:, 3]
    ms = min_subblock_spacing
    drawn = rng.random(n) <= subblock_chance
    big = ((maxx - minx + 1) >= 2 * ms + 1) & ((maxy - miny + 1) >= 2 * ms + 1)
    rect = (maxx - minx + 1).astype(np.int64) * (maxy - miny + 1) == blobs[:, 4]
    carved = drawn & big & rect
    px = minx + ms + (rng.random(n) * np.maximum(maxx - ms - (minx + ms) + 1, 1)).astype(np.int32)
    py = miny + ms + (rng.random(n) * np.maximum(maxy - ms - (miny + ms) + 1, 1)).astype(np.int32)
    tape[:, 0] = drawn
    tape[:, 1] = carved
    tape[:, 2] = np.where(carved, px, 0)
    tape[:, 3] = np.where(carved, py, 0)
    tape[:, 4] = np.where(carved, np.where(rng.random(n) < 0.5, 3, 1), 0)   # W / E
    tape[:, 5] = np.where(carved, np.where(rng.random(n) < 0.5, 0, 2), 0)   # N / S
    tape[:, 6] = np.where(carved, rng.random(n) < 0.5, 0)
    tape[:, 7] = carved
    return tape


def synth_carve_uniforms(seed: int, cap: int) -> np.ndarray:
    """Six uniforms per GLOBAL blob id: the shared randomness of ``synth_carve_tape_device`` (replicated on every shard)."""
    return np.random.default_rng(seed).random((cap, 6), dtype=np.float32)


def synth_carve_tape_device(uniforms, table, count, id_base, out, min_subblock_spacing: int = 5, subblock_chance: float = 0.3):
    """Carve tape rows for the blobs of one window, written into the global-id-indexed tape ``out`` [cap, 8] (torch, on device).

    ``table`` [cap_w, 6] / ``count`` / ``id_base`` are the window's component table (tsim_blobs); ``out`` has one
    more row than there are ids (a scratch row).  A blob's row is a
    function of its global id (through ``uniforms``) and of its bounding box only, so every shard that sees the whole
    blob writes the same row; blobs cut by a window edge (partial bounding box) are skipped by the carve kernel.
    Same acceptance rules as ``synth_carve_tape`` (city_model.py:649-682).
    """
    import torch
    ms = min_subblock_spacing
    k = torch.arange(table.shape[0], device=table.device)
    gid = k + id_base.to(torch.int64)
    ok = (k < count) & (gid >= 0) & (gid < out.shape[0] - 1)
    g = gid.clamp(0, out.shape[0] - 2)
    u = uniforms[g]
    minx, miny, maxx, maxy, size = (table[:, i].to(torch.int64) for i in range(5))
    w, h = maxx - minx + 1, maxy - miny + 1
    carved = ok & (u[:, 0] <= subblock_chance) & (w >= 2 * ms + 1) & (h >= 2 * ms + 1) & (w * h == size)
    px = minx + ms + (u[:, 1] * (w - 2 * ms).clamp(min=1)).to(torch.int64)
    py = miny + ms + (u[:, 2] * (h - 2 * ms).clamp(min=1)).to(torch.int64)
    px, py = torch.minimum(px, maxx - ms), torch.minimum(py, maxy - ms)
    rows = torch.stack([(ok & (u[:, 0] <= subblock_chance)).to(torch.int64), carved.to(torch.int64), px, py,
                        torch.where(u[:, 3] < 0.5, 3, 1), torch.where(u[:, 4] < 0.5, 0, 2), (u[:, 5] < 0.5).to(torch.int64),
                        carved.to(torch.int64)], 1)
    rows = torch.where(carved[:, None], rows, torch.zeros_like(rows)).to(torch.int32)
    out[torch.where(ok, g, torch.full_like(g, out.shape[0] - 1))] = rows   # the last row of `out` is a scratch row (no host sync)
    return out


def synth_traffic(seed: int, W: int, H: int, cell_type: np.ndarray, dirs: np.ndarray, n_vehicles: int, n_ticks: int,
                  route_len: int = 200, spawn_ticks: int = 1, malfunction_p: float = 0.0, sideswipe_p: float = 0.0):
    """Synthetic tick tapes for cities too large for the reference's A* (SURVEY.md §8d configs 4/5).

    Vehicles are spawn attempts on distinct random road cells during the first `spawn_ticks` ticks; each
    follows a pre-planned route that obeys the allowed directions of every cell it leaves (a random walk
    over the arrow graph without immediate U-turns, cut at dead ends); the route's last cell is the target.
    Returns the tape dict consumed by `GpuTraffic` (and by the tick oracle).
    """
    rng = np.random.default_rng(seed)
    T = cell_type.reshape(-1)
    D = dirs.reshape(-1).astype(np.int64)
    road = np.flatnonzero((D & 0xF) != 0)
    n_vehicles = min(n_vehicles, len(road))
    origin = rng.choice(road, size=n_vehicles, replace=False).astype(np.int64)
    step = np.array([W, 1, -W, -1], np.int64)
    pos = origin.copy()
    last_dir = np.full(n_vehicles, -1, np.int64)
    alive = np.ones(n_vehicles, bool)
    cols = []
    for _ in range(route_len):
        d = D[pos]
        n = (d >> 12) & 7
        pick = (rng.random(n_vehicles) * np.maximum(n, 1)).astype(np.int64)
        choice = (d >> (4 + 2 * pick)) & 3
        # avoid an immediate U-turn when another arrow exists
        uturn = (choice == (last_dir + 2) % 4) & (last_dir >= 0) & (n > 1)
        alt = (d >> (4 + 2 * ((pick + 1) % np.maximum(n, 1)))) & 3
        choice = np.where(uturn, alt, choice)
        x, y = pos % W, pos // W
        nx, ny = x + np.array([0, 1, 0, -1])[choice], y + np.array([1, 0, -1, 0])[choice]
        ok = alive & (n > 0) & (nx >= 0) & (nx < W) & (ny >= 0) & (ny < H)
        nxt = np.where(ok, ny * W + nx, pos)
        ok &= (D[nxt] & 0xF) != 0          # never step onto an arrow-less cell
        nxt = np.where(ok, nxt, pos)
        alive = ok
        cols.append(np.where(ok, nxt, -1))
        pos = nxt
        last_dir = np.where(ok, choice, last_dir)
    route = np.stack(cols, 1)                                  # [V, route_len], -1 padded at the tail
    length = (route >= 0).sum(1)
    keep = length >= 1
    origin, route, length = origin[keep], route[keep], length[keep]
    nv = len(origin)
    target = route[np.arange(nv), length - 1]
    ok = target != origin
    origin, route, length, target = origin[ok], route[ok], length[ok], target[ok]
    nv = len(origin)
    spawn_tick = np.sort(rng.integers(0, spawn_ticks, size=nv)).astype(np.int32)
    ev_off = np.zeros(nv + 1, np.int64)
    ev_off[1:] = np.cumsum(length)
    ev_cells = route[route >= 0].astype(np.int32)
    tp = dict(
        spawn_tick=spawn_tick, origin=origin.astype(np.int32), target=target.astype(np.int32),
        speed=rng.integers(1, 6, size=(n_ticks, nv), dtype=np.uint8),
        malfunction=(rng.random((n_ticks, nv)) < malfunction_p).astype(np.uint8) if malfunction_p > 0 else np.zeros((n_ticks, nv), np.uint8),
        rank=np.argsort(rng.random((n_ticks, nv)), axis=1).astype(np.int32),
        ev_tick=spawn_tick.copy(), ev_vehicle=np.arange(nv, dtype=np.int32), ev_off=ev_off, ev_cells=ev_cells,
        rain_map=np.zeros((H, W), np.uint8))
    if sideswipe_p > 0:   # bit 1: the sideswipe draw of that (tick, vehicle) fires if it is made; its own stream, the other tapes stay as they are
        tp["malfunction"] |= (np.random.default_rng(seed ^ 0x51DE).random((n_ticks, nv)) < sideswipe_p).astype(np.uint8) << 1
    return tp


def synth_trips(seed: int, W: int, H: int, cell_type: np.ndarray, dirs: np.ndarray, trips_per_tick: int, n_ticks: int, route_len: int = 80):
    """Tick tapes for SURVEY.md §8d config 4 in its literal form: `trips_per_tick` vehicle trips INJECTED EVERY TICK (100 k trips
    over 1000 ticks on the default city), rather than one fleet spawned at tick 0 (`synth_traffic`).

    Every trip is a spawn attempt of its tick on a random road cell (with replacement: an attempt on an occupied cell fails,
    as in the reference, city_model.py:1897-1918) with a pre-planned route like `synth_traffic`'s.  The activation order of a
    tick is an affine permutation `(a_t * v + b_t) mod P` of the vehicle indices (any order is a valid tape; this one costs no
    sort of a 1000 x 100 000 matrix).  Returns the tape dict consumed by `GpuTraffic` and the tick oracle."""
    rng = np.random.default_rng(seed)
    D = dirs.reshape(-1).astype(np.int64)
    road = np.flatnonzero((D & 0xF) != 0)
    nv0 = trips_per_tick * n_ticks
    base = synth_traffic(seed, W, H, cell_type, dirs, 0, 1, route_len=route_len)   # empty tapes with the right keys
    origin = rng.choice(road, size=nv0, replace=True).astype(np.int64)
    step_x, step_y = np.array([0, 1, 0, -1]), np.array([1, 0, -1, 0])
    pos, last_dir, alive = origin.copy(), np.full(nv0, -1, np.int64), np.ones(nv0, bool)
    cols = []
    for _ in range(route_len):
        d = D[pos]
        n = (d >> 12) & 7
        pick = (rng.random(nv0) * np.maximum(n, 1)).astype(np.int64)
        choice = (d >> (4 + 2 * pick)) & 3
        uturn = (choice == (last_dir + 2) % 4) & (last_dir >= 0) & (n > 1)
        choice = np.where(uturn, (d >> (4 + 2 * ((pick + 1) % np.maximum(n, 1)))) & 3, choice)
        nx, ny = pos % W + step_x[choice], pos // W + step_y[choice]
        ok = alive & (n > 0) & (nx >= 0) & (nx < W) & (ny >= 0) & (ny < H)
        nxt = np.where(ok, ny * W + nx, pos)
        ok &= (D[nxt] & 0xF) != 0
        nxt = np.where(ok, nxt, pos)
        alive = ok
        cols.append(np.where(ok, nxt, -1).astype(np.int32))
        pos = nxt
        last_dir = np.where(ok, choice, last_dir)
    route = np.stack(cols, 1)
    length = (route >= 0).sum(1)
    tick_of = np.repeat(np.arange(n_ticks, dtype=np.int32), trips_per_tick)
    target = route[np.arange(nv0), np.maximum(length - 1, 0)]
    keep = (length >= 1) & (target != origin)
    origin, route, length, target, tick_of = origin[keep], route[keep], length[keep], target[keep], tick_of[keep]
    nv = len(origin)
    ev_off = np.zeros(nv + 1, np.int64)
    ev_off[1:] = np.cumsum(length)
    P = 2147483647                                        # prime > any vehicle count: v -> (a v + b) mod P is injective
    a = rng.integers(1, P, size=(n_ticks, 1), dtype=np.int64)
    b = rng.integers(0, P, size=(n_ticks, 1), dtype=np.int64)
    rank = np.empty((n_ticks, nv), np.int32)              # row by row: no [T, V] int64 temporaries
    idx = np.arange(nv, dtype=np.int64)
    for t in range(n_ticks):
        rank[t] = (a[t, 0] * idx + b[t, 0]) % P
    base.update(spawn_tick=tick_of, origin=origin.astype(np.int32), target=target.astype(np.int32),
                speed=rng.integers(1, 6, size=(n_ticks, nv), dtype=np.uint8), malfunction=np.zeros((n_ticks, nv), np.uint8), rank=rank,
                ev_tick=tick_of.copy(), ev_vehicle=np.arange(nv, dtype=np.int32), ev_off=ev_off, ev_cells=route[route >= 0].astype(np.int32),
                rain_map=np.zeros((H, W), np.uint8))
    return base


def synth_planned_trips(seed: int, W: int, H: int, cell_type: np.ndarray, trips_per_tick: int, n_ticks: int, malfunction_p: float = 0.0):
    """Tick tapes WITHOUT routes for ``replan.PlannedTraffic`` (SURVEY.md §8d config 4 as named: trips injected every tick, the
    vehicles plan their own routes).  A trip starts on a cell of ``CityModel.get_start_blocks()`` (BlockEntrance / HighwayEntrance,
    city_model.py:2102-2109) and ends on one of ``get_exit_blocks()`` (BlockEntrance / HighwayExit, :2111-2118) other than its
    origin -- the harness's trip rule (oracle/refharness/ticks.py).  Activation order: an affine permutation per tick, as in
    ``synth_trips``."""
    rng = np.random.default_rng(seed)
    T = np.asarray(cell_type).reshape(-1)
    starts = np.flatnonzero((T == 19) | (T == 13))
    exits = np.flatnonzero((T == 19) | (T == 14))
    if len(starts) == 0 or len(exits) < 2:
        raise ValueError("the city has no entrance / exit cells to run trips between")
    nv = int(trips_per_tick) * int(n_ticks)
    origin = starts[rng.integers(len(starts), size=nv)]
    target = exits[rng.integers(len(exits), size=nv)]
    same = target == origin
    while same.any():                     # a trip to its own origin is never generated
        target[same] = exits[rng.integers(len(exits), size=int(same.sum()))]
        same = target == origin
    P = 2147483647
    a = rng.integers(1, P, size=n_ticks, dtype=np.int64)
    b = rng.integers(0, P, size=n_ticks, dtype=np.int64)
    rank = np.empty((n_ticks, nv), np.int32)
    idx = np.arange(nv, dtype=np.int64)
    for t in range(n_ticks):
        rank[t] = (a[t] * idx + b[t]) % P
    return dict(spawn_tick=np.repeat(np.arange(n_ticks, dtype=np.int32), trips_per_tick), origin=origin.astype(np.int32),
                target=target.astype(np.int32), speed=rng.integers(1, 6, size=(n_ticks, nv), dtype=np.uint8),
                malfunction=(rng.random((n_ticks, nv)) < malfunction_p).astype(np.uint8) if malfunction_p > 0 else np.zeros((n_ticks, nv), np.uint8),
                rank=rank, ev_tick=np.zeros(0, np.int32), ev_vehicle=np.zeros(0, np.int32), ev_off=np.zeros(1, np.int64),
                ev_cells=np.zeros(0, np.int32), rain_map=np.zeros((H, W), np.uint8))


def spawn_tape_from_trips(depart_secs, origin_cells, target_cells, dt, n_ticks: int, elapsed: float = 0.0):
    """The reference's trip schedule as a spawn tape.

    ``DynamicTrafficAgent._generate_day`` (agents/dynamic_traffic_generator.py:307-396) fills ``pending`` with trips -- departure
    time, origin cell, destination cell -- from draws of Python's ``random``; those draws are decisions, i.e. an input here (DESIGN.md
    §2).  What is restated is WHEN ``step`` (:151-189) hands a pending trip to ``_spawn``: the clock advances by ``dt`` per tick BY
    REPEATED FLOAT ADDITION (``self.elapsed += self.dt``) and tick k spawns, in ``pending`` order, the trips with
    ``elapsed_k < depart_secs <= elapsed_k + dt``.  Trips that depart at or before ``elapsed`` are never spawned by the reference
    either (the comparison is strict); trips after the last tick stay pending.

    depart_secs / origin_cells / target_cells: one entry per pending trip, in ``pending`` order (cell = y * W + x).
    Returns dict(spawn_tick, origin, target, trip) sorted by tick (stable), ``trip`` = index into the inputs.
    """
    d = np.asarray(depart_secs, np.float64)
    edges = np.empty(n_ticks + 1, np.float64)
    e = float(elapsed)
    edges[0] = e
    for k in range(n_ticks):   # the reference's own accumulation, not k * dt
        e += dt
        edges[k + 1] = e
    tick = np.searchsorted(edges, d, side="left") - 1          # edges[tick] < d <= edges[tick + 1]
    keep = np.flatnonzero((tick >= 0) & (tick < n_ticks))
    order = keep[np.argsort(tick[keep], kind="stable")]
    return dict(spawn_tick=tick[order].astype(np.int32), origin=np.asarray(origin_cells, np.int64)[order].astype(np.int32),
                target=np.asarray(target_cells, np.int64)[order].astype(np.int32), trip=order.astype(np.int32))


def spawn_tape_from_generator(gen, width: int, n_ticks: int, kinds=("internal", "through")):
    """``spawn_tape_from_trips`` for a live ``DynamicTrafficAgent`` (``model.dynamic_traffic_generator``): reads its ``pending`` list,
    clock and ``dt`` without touching them.  Service trips (``service_food`` / ``service_waste``: ``ServiceVehicleAgent``, out of
    scope) are left out.  Returns the tape and the trips it was made from."""
    trips = [t for t in gen.pending if t.kind in kinds]
    cell = lambda a: a.position[1] * width + a.position[0]
    tape = spawn_tape_from_trips([t.depart_secs for t in trips], [cell(t.origin) for t in trips], [cell(t.destination) for t in trips],
                                 gen.dt, n_ticks, gen.elapsed)
    return tape, trips
