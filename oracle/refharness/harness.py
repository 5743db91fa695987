"""Drive the UNMODIFIED reference `CityModel` under a recorded random tape and extract planes.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Runs only where `/root/reference` exists
(the build container); its outputs travel to the GPU box as fixtures in `tests/golden/`.

What it does:
  * installs the import stubs, imports `Simulation.city_model`;
  * wraps the global `random` module's draw functions so every *decision* the layout passes
    make is recorded (band lists, zone index per block, R4 carve parameters per blob, chosen
    longest run per block) -- the "layout tape";
  * wraps each of the 13 pass methods called at `city_model.py:125-148` so planes can be
    snapshotted after any pass, and guards the F8 crash in light-group construction;
  * converts the per-cell `CellAgent` objects into the packed planes the GPU path uses.

Encoding (must equal `trafficsimulation_b200/encoding.py`; a test checks it):
  cell_type u8  = index into `Defaults.ZONES` (config.py:74-95)
  dirs      u16 = bits0-3 mask (N=1,E=2,S=4,W=8 as city_model.py:2191-2196),
                  bits 4+2i..5+2i = i-th list entry (N=0,E=1,S=2,W=3), bits12-14 = len
  aux       u8  = bits0-4 original type code (ControlledRoad only), bit5 ring-corner cell
                  (`_ring_road_cells`), bit6 ever-intersection (`_intersection_cells`),
                  bit7 cell.light is not None
  block_id  i32 = 1-based block id on region cells and on the block's BlockEntrance, else 0
"""
from __future__ import annotations

import contextlib
import os
import random as _random
import tempfile

import numpy as np

from . import stubs

DIR_INDEX = {"N": 0, "E": 1, "S": 2, "W": 3}
DIR_NAMES = ["N", "E", "S", "W"]
ROAD_CODE = {"R1": 1, "R2": 2, "R3": 3}

LAYOUT_PASSES = [
    "_place_thick_wall", "_place_sidewalk_inner_ring", "_clear_interior",
    "_build_roads_and_sidewalks", "_carve_subblock_roads",
    "_flood_fill_blocks_storing_data", "_eliminate_dead_ends",
    "_upgrade_r2_to_intersections", "_final_place_block_entrances",
    "_remove_invalid_intersection_directions", "_add_entrance_directions",
    "_add_traffic_lights",
]

_ref = None


def load_reference():
    """Import the reference once; returns a namespace with its main symbols."""
    global _ref
    if _ref is not None:
        return _ref
    stubs.install()
    import types
    from Simulation.config import Defaults
    # constructing a CityModel writes ./Results unless these are off
    Defaults.SAVE_TOTAL_RESULTS = False
    Defaults.SAVE_INDIVIDUAL_RESULTS = False
    import Simulation.city_model as cm
    from Simulation.agents.vehicles.vehicle_base import VehicleAgent
    _ref = types.SimpleNamespace(cm=cm, CityModel=cm.CityModel, Defaults=Defaults,
                                 VehicleAgent=VehicleAgent)
    return _ref


def encode_dirs(dirs) -> int:
    code = 0
    for i, d in enumerate(dirs):
        k = DIR_INDEX[d]
        code |= (1 << k)
        code |= k << (4 + 2 * i)
    code |= len(dirs) << 12
    return code


def decode_dirs(code: int):
    n = (code >> 12) & 7
    return [DIR_NAMES[(code >> (4 + 2 * i)) & 3] for i in range(n)]


def extract_planes(model, blocks=True):
    """Planes of the *current* grid state of a (possibly half-built) reference model."""
    ref = load_reference()
    zones = ref.Defaults.ZONES
    tcode = {z: i for i, z in enumerate(zones)}
    W, H = model.width, model.height
    ctype = np.zeros((H, W), np.uint8)
    dirs = np.zeros((H, W), np.uint16)
    aux = np.zeros((H, W), np.uint8)
    bid = np.zeros((H, W), np.int32)
    ring = getattr(model, "_ring_road_cells", set())
    ever = getattr(model, "_intersection_cells", set())
    for x in range(W):
        col = model.grid._cells[x]
        for y in range(H):
            cell = col[y][0]
            t = tcode[cell.cell_type]
            ctype[y, x] = t
            if cell.directions:
                dirs[y, x] = encode_dirs(cell.directions)
            a = 0
            if cell.cell_type == "ControlledRoad":
                a |= tcode[cell.road_type]
            if cell.light is not None:
                a |= 0x80
            aux[y, x] = a
            if cell.cell_type == "BlockEntrance" and cell.block_id is not None:
                bid[y, x] = cell.block_id
    for (x, y) in ring:
        aux[y, x] |= 0x20
    for (x, y) in ever:
        aux[y, x] |= 0x40
    if blocks:
        for info in model._blocks_data:
            b = info["block_id"]
            for (x, y) in info["region"]:
                bid[y, x] = b
    return {"cell_type": ctype, "dirs": dirs, "aux": aux, "block_id": bid}


def extract_simple_maps(model):
    """The four derived maps of `_build_simple_maps` (city_model.py:2151-2199)."""
    return {
        "is_road_map": model.is_road_map.astype(np.uint8).copy(),
        "road_type_map": model.road_type_map.astype(np.uint8).copy(),
        "intersection_map": model.intersection_map.astype(np.uint8).copy(),
        "allowed_dirs_map": model.allowed_dirs_map.astype(np.uint8).copy(),
    }


def extract_light_links(model):
    """Traffic-light link tables as sorted multisets.

    Returns int32 arrays of cell indices (y*W+x):
      lights[n]            TrafficLight cells, sorted
      ctrl[m,2]            (light cell, controlled road cell) pairs, sorted, with multiplicity
      incoming[k,2]        (light cell, assigned incoming lane cell) pairs, sorted, with multiplicity
      outgoing[j,2]        same for assigned_outgoing_road_blocks (forward_traffic_light_range only)
    """
    W = model.width
    lights, ctrl, inc, outg = [], [], [], []
    for tl in model.traffic_lights:
        lx, ly = tl.position
        li = ly * W + lx
        lights.append(li)
        for cb in tl.controlled_blocks:
            ctrl.append((li, cb.position[1] * W + cb.position[0]))
        for rb in tl.assigned_incoming_road_blocks:
            inc.append((li, rb.position[1] * W + rb.position[0]))
        for rb in tl.assigned_outgoing_road_blocks:
            outg.append((li, rb.position[1] * W + rb.position[0]))
    lights = np.array(sorted(lights), np.int32)
    ctrl = np.array(sorted(ctrl), np.int32).reshape(-1, 2)
    inc = np.array(sorted(inc), np.int32).reshape(-1, 2)
    outg = np.array(sorted(outg), np.int32).reshape(-1, 2)
    return {"lights": lights, "ctrl": ctrl, "incoming": inc, "outgoing": outg}


class _Recorder:
    """Logs every draw made through the global `random` module while active."""

    NAMES = ("random", "randint", "choice", "choices", "gauss", "uniform")

    def __init__(self):
        self.log = []
        self.tag = None
        self._orig = {}

    def __enter__(self):
        for n in self.NAMES:
            self._orig[n] = getattr(_random, n)
            setattr(_random, n, self._wrap(n, self._orig[n]))
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(_random, n, f)

    def _wrap(self, name, fn):
        def wrapped(*a, **k):
            r = fn(*a, **k)
            self.log.append((self.tag, name, a, r))
            return r
        return wrapped


def _bands_array(bands):
    out = np.zeros((len(bands), 4), np.int32)
    for i, (st, en, rt, bd) in enumerate(bands):
        out[i] = (st, en, ROAD_CODE[rt], DIR_INDEX.get(bd, -1))
    return out


def _parse_carve(log, chance):
    """Turn the draw log of `_carve_subblock_roads` (city_model.py:649-682) into per-blob records.

    One row per Nothing blob in discovery order:
      (drawn<=chance, carved, px, py, hor_dir, ver_dir, inbound_is_horizontal, n_tries)
    `carved` is 1 only when a pivot was accepted (the leg choice was drawn).
    """
    rows = []
    i, n = 0, len(log)
    while i < n:
        _, name, a, r = log[i]
        assert name == "random", (name, a)
        i += 1
        row = [int(not (r > chance)), 0, 0, 0, 0, 0, 0, 0]
        tries = 0
        last = None
        while i + 3 < n + 1 and i < n and log[i][1] == "randint":
            px = log[i][3]; py = log[i + 1][3]
            hd = log[i + 2][3]; vd = log[i + 3][3]
            assert log[i + 2][1] == "choice" and log[i + 3][1] == "choice"
            last = (px, py, DIR_INDEX[hd], DIR_INDEX[vd])
            tries += 1
            i += 4
        row[7] = tries
        if i < n and log[i][1] == "choice":
            leg = log[i][3]
            i += 1
            row[1] = 1
            row[2], row[3], row[4], row[5] = last
            row[6] = 1 if leg[0] == "horizontal" else 0
        rows.append(row)
    return np.array(rows, np.int32).reshape(-1, 8)


def _parse_entrances(log):
    """Canonical index of the chosen longest run per block that reached `random.choice`
    (city_model.py:944).  Runs are ordered by their minimal (y, x) cell."""
    out = []
    for _, name, a, r in log:
        assert name == "choice"
        runs = a[0]
        keyed = sorted(range(len(runs)), key=lambda j: min((p[1], p[0]) for p in runs[j]))
        pos = [j for j in range(len(runs)) if runs[j] is r][0]
        out.append(keyed.index(pos))
    return np.array(out, np.int32)


@contextlib.contextmanager
def _in_tmpdir():
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield
        finally:
            os.chdir(cwd)


def run_layout(seed, snapshots=(), enable_traffic=False, enable_rain=False, keep_model=False,
               **kwargs):
    """Build the reference city with `random.seed(seed)`; return tapes + planes.

    `snapshots`: names from LAYOUT_PASSES after which planes are captured too.
    Layout state is captured right after `_add_traffic_lights`, so the F8 crash in
    `_create_intersection_light_groups` cannot lose it.
    """
    ref = load_reference()
    cm, Defaults = ref.cm, ref.Defaults
    rec = _Recorder()
    snaps = {}
    state = {}
    originals = {}

    def make_wrapper(name, fn):
        def wrapper(self, *a, **k):
            rec.tag = name
            out = fn(self, *a, **k)
            rec.tag = None
            if name in snapshots:
                snaps[name] = extract_planes(self)
            if name == "_add_traffic_lights":
                state["final"] = extract_planes(self)
                state["links"] = extract_light_links(self)
                state["hbands"] = _bands_array(self.horizontal_bands)
                state["vbands"] = _bands_array(self.vertical_bands)
            return out
        return wrapper

    old_flags = (Defaults.ENABLE_TRAFFIC, Defaults.RAIN_ENABLED)
    Defaults.ENABLE_TRAFFIC = enable_traffic
    Defaults.RAIN_ENABLED = enable_rain
    for name in LAYOUT_PASSES:
        originals[name] = getattr(cm.CityModel, name)
        setattr(cm.CityModel, name, make_wrapper(name, originals[name]))
    model = None
    crashed = None
    try:
        with rec, _in_tmpdir():
            _random.seed(seed)
            try:
                model = cm.CityModel(seed=seed, **kwargs)
            except ZeroDivisionError as e:  # F8
                crashed = repr(e)
    finally:
        for name, fn in originals.items():
            setattr(cm.CityModel, name, fn)
        Defaults.ENABLE_TRAFFIC, Defaults.RAIN_ENABLED = old_flags

    chance = kwargs.get("subblock_chance", Defaults.SUBBLOCK_CHANGE)
    by_tag = {}
    for e in rec.log:
        by_tag.setdefault(e[0], []).append(e)
    zone_names = Defaults.AVAILABLE_CITY_BLOCKS
    out = {
        "seed": seed,
        "kwargs": dict(kwargs),
        "hbands": state["hbands"], "vbands": state["vbands"],
        "tape_zone": np.array([zone_names.index(e[3][0])
                               for e in by_tag.get("_flood_fill_blocks_storing_data", [])], np.uint8),
        "tape_carve": _parse_carve(by_tag.get("_carve_subblock_roads", []), chance),
        "tape_entrance": _parse_entrances(by_tag.get("_final_place_block_entrances", [])),
        "final": state["final"], "links": state["links"],
        "snaps": snaps, "crashed": crashed,
    }
    if model is not None:
        out["maps"] = extract_simple_maps(model)
        out["n_blocks"] = len(model._blocks_data)
    if keep_model:
        out["model"] = model
    return out
