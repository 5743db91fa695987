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
