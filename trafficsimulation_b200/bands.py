"""Host-side road-band generator: O(W+H) work, stays in Python.

Mirrors the draw ORDER of the reference so that ``random.seed(k)`` followed by ``make_city_bands``
yields the very band lists ``CityModel._build_roads_and_sidewalks`` builds at city_model.py:380-394
(`_make_road_bands_for_interior` :1076-1177, `_choose_road_type` :1179-1204, `_force_one_highway`
:1206-1231, `_ensure_minimum_highways` :1233-1267).  All draws go through the ``rnd`` object, by
default the global ``random`` module, exactly as the reference does.

A band is ``(start, end, type, dir)`` with type in {"R1","R2","R3"} and dir a direction letter (or
"" for an inserted highway).  ``bands_to_array`` gives the int32 [n,4] form of the C ABI.
"""
from __future__ import annotations

import random as _random
from dataclasses import dataclass

import numpy as np

from .encoding import DIR_INDEX, ROAD_CODE

ROAD_THICKNESS = {"R1": 4, "R2": 2, "R3": 1}          # config.py:45-49
_OPPOSITE = {"N": "S", "S": "N", "E": "W", "W": "E"}  # config.py:65


@dataclass
class BandParams:
    """The constructor kwargs the band generator reads (city_model.py:27-46)."""
    width: int = 200
    height: int = 200
    wall_thickness: int = 15
    sidewalk_ring_width: int = 2
    ring_road_type: str | None = "R2"
    r1_chance_mean: float = 0.15
    r1_chance_std: float = 0.03
    r2_chance_mean: float = 0.70
    r2_chance_std: float = 0.05
    min_r1_bands: int = 2
    min_block_spacing: int = 6
    max_block_spacing: int = 18
    highway_offset_from_edges: int = 7

    @property
    def interior(self):
        m = self.wall_thickness + self.sidewalk_ring_width
        return m, self.width - m - 1, m, self.height - m - 1   # x_min, x_max, y_min, y_max (:91-94)


def _clip01(v):
    return max(0.0, min(1.0, v))


def _draw_road_type(p: BandParams, rnd) -> str:
    p1 = _clip01(rnd.gauss(p.r1_chance_mean, p.r1_chance_std))
    p2 = _clip01(min(1.0 - p1, rnd.gauss(p.r2_chance_mean, p.r2_chance_std)))
    r = rnd.random()
    if r < p1:
        return "R1"
    return "R2" if r < p1 + p2 else "R3"


def _axis_bands(lo: int, hi: int, horizontal: bool, p: BandParams, rnd):
    letters = ["E", "W"] if horizontal else ["N", "S"]
    out, pos, prev_r3 = [], lo, None
    while pos <= hi:
        kind = _draw_road_type(p, rnd)
        last = min(pos + ROAD_THICKNESS[kind] - 1, hi)
        if kind == "R3" and prev_r3 is not None:
            heading = _OPPOSITE[prev_r3]          # neighbouring one-way streets alternate
        else:
            heading = rnd.choice(letters)
        out.append((pos, last, kind, heading))
        prev_r3 = heading if kind == "R3" else None
        if last + 1 > hi:
            break
        gap = rnd.randint(p.min_block_spacing, p.max_block_spacing)
        if last + gap > hi:
            break
        pos = last + gap + 1
    ring = p.ring_road_type
    if ring is not None:
        t = ROAD_THICKNESS[ring]
        if ring == "R3":
            d_first, d_last = ("E", "W") if horizontal else ("S", "N")
        else:
            d_first = rnd.choice(letters)
            d_last = rnd.choice(letters)
        first, final = (lo, lo + t - 1, ring, d_first), (hi - t + 1, hi, ring, d_last)
        if not out:
            out = [first, final]
        elif len(out) == 1:
            out = [first] if first == final else [first, final]
        else:
            out[0], out[-1] = first, final
    return out


def _insert_highway(bands, total: int, p: BandParams, rnd):
    t = ROAD_THICKNESS["R1"]
    inset = p.wall_thickness + p.sidewalk_ring_width + p.highway_offset_from_edges   # :309-310
    lo, hi = inset, total - t - inset
    if lo > hi:
        lo, hi = 0, total - t
        if hi < 0:
            return
    s = rnd.randint(lo, hi)
    e = s + t - 1
    bands.append((s, e, "R1", ""))
    bands.sort(key=lambda b: b[0])
    keep_lo, keep_hi = s - p.min_block_spacing, e + p.min_block_spacing
    bands[:] = [b for b in bands
                if (b[2] == "R1" and (b[0], b[1]) == (s, e)) or b[1] < keep_lo or b[0] > keep_hi]


def _ensure_highways(bands, total: int, p: BandParams, rnd):
    def count():
        idx = range(1, len(bands) - 1) if (p.ring_road_type == "R1" and len(bands) >= 2) else range(len(bands))
        return sum(1 for i in idx if bands[i][2] == "R1")
    tries = 0
    while count() < p.min_r1_bands and tries < 20:
        _insert_highway(bands, total, p, rnd)
        tries += 1


def make_city_bands(p: BandParams, rnd=_random):
    """(horizontal_bands, vertical_bands) in the reference's draw order (city_model.py:380-394)."""
    x_min, x_max, y_min, y_max = p.interior
    hb = _axis_bands(y_min, y_max, True, p, rnd)
    vb = _axis_bands(x_min, x_max, False, p, rnd)
    _ensure_highways(hb, p.height, p, rnd)
    _ensure_highways(vb, p.width, p, rnd)
    return hb, vb


def bands_to_array(bands) -> np.ndarray:
    out = np.zeros((len(bands), 4), np.int32)
    for i, (s, e, kind, heading) in enumerate(bands):
        out[i] = (s, e, ROAD_CODE[kind], DIR_INDEX.get(heading, -1))
    return out
