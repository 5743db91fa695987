"""`fill_model_from_planes` (the bulk materialiser of the seam) on a duck-typed model, without the reference
and without a GPU: planes + link tables of the reference fixtures go in, the model's cells / trackers are read back with the
harness's extraction rules and must give the same planes and tables (the same check against the LIVE reference model is
tests/test_adaptor_reference.py)."""
import os

import numpy as np
import pytest

from fake_model import FakeModel, extract_links, extract_planes, fake_defaults
from golden_util import layout_fixtures, load, PLANES

FIX = [p for p in layout_fixtures() if any(k in p for k in ("default12345", "s14_150x110_carve", "s16_64_carve", "s26_hw4", "s22_fwd_inrange",
                                                            "s23_fwd_extra_carve"))]


@pytest.mark.parametrize("path", FIX, ids=lambda p: os.path.basename(p)[7:-4])
def test_fill_round_trips_fixture(path):
    from trafficsimulation_b200.adaptor import fill_model_from_planes
    g = load(path)
    m = FakeModel(**g["meta"]["cfg"])
    planes = {k: g[k] for k in PLANES}
    links = {"lights": g["links_lights"], "ctrl": g["links_ctrl"], "incoming": g["links_incoming"], "outgoing": g["links_outgoing"]}
    fill_model_from_planes(m, planes, links, g["hbands"], g["vbands"], defaults=fake_defaults())
    H, W = planes["cell_type"].shape
    assert m.place_calls == W * H
    got = extract_planes(m)
    for f in PLANES:
        assert np.array_equal(got[f], planes[f]), f
    gl = extract_links(m)
    for k in ("lights", "ctrl", "incoming", "outgoing"):
        assert np.array_equal(gl[k], links[k]), k
    # trackers (city_model.py:96-107)
    T = planes["cell_type"]
    count = lambda name: int((T == g_code(name)).sum())
    from trafficsimulation_b200.encoding import TYPE_CODE
    g_code = TYPE_CODE.__getitem__
    assert len(m.block_entrances) == count("BlockEntrance") and len(m.controlled_roads) == count("ControlledRoad")
    assert len(m.highway_entrances) == count("HighwayEntrance") and len(m.highway_exits) == count("HighwayExit")
    assert len(m._blocks_data) == g["meta"]["n_blocks"] == int(planes["block_id"].max())
    assert [b["block_id"] for b in m._blocks_data] == list(range(1, len(m._blocks_data) + 1))
    zone = np.isin(T, [g_code(z) for z in ("Residential", "Office", "Market", "Leisure", "Other", "Empty")])
    for info in m._blocks_data[:: max(1, len(m._blocks_data) // 25)]:
        b = info["block_id"]
        ys, xs = np.nonzero(zone & (planes["block_id"] == b))
        assert sorted(info["region"]) == sorted(zip(xs.tolist(), ys.tolist()))
        reg = set(info["region"])
        ring = {(x + dx, y + dy) for x, y in reg for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))
                if 0 <= x + dx < W and 0 <= y + dy < H and (x + dx, y + dy) not in reg}
        assert info["ring"] == sorted(ring)
    assert (m.stop_map.reshape(-1)[links["lights"]] == 0).all()


def test_fill_window_materialises_only_the_window():
    from trafficsimulation_b200.adaptor import fill_model_from_planes
    g = load([p for p in layout_fixtures() if "default12345" in p][0])
    planes = {k: g[k] for k in PLANES}
    links = {"lights": g["links_lights"], "ctrl": g["links_ctrl"], "incoming": g["links_incoming"]}
    m = FakeModel(**g["meta"]["cfg"])
    fill_model_from_planes(m, planes, links, window=(40, 50, 90, 80), defaults=fake_defaults())
    assert m.place_calls == 50 * 30 and set(m.cells) == {(x, y) for x in range(40, 90) for y in range(50, 80)}
    full = FakeModel(**g["meta"]["cfg"])
    fill_model_from_planes(full, planes, links, defaults=fake_defaults())
    for xy, c in m.cells.items():   # (cell.light may point at a light outside the window: not linked, everything else is)
        f = full.cells[xy]
        assert (c.cell_type, c.directions, c.road_type, c.block_id) == (f.cell_type, f.directions, f.road_type, f.block_id), xy
    assert len(m._blocks_data) == len(full._blocks_data)
