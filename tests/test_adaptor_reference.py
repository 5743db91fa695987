"""The drop-in seam of INTEGRATION.md §2, checked against the LIVE reference (runs where /root/reference exists).

A second reference CityModel is constructed with its twelve layout pass methods replaced by the adaptor's fill
(`fill_model_from_planes`, fed with the planes / link tables of a normal reference run -- the GPU path is proven
equal to those by tests/test_gpu_layout.py).  Everything AFTER the passes in `CityModel.__init__` -- light-group
construction, city blocks, cell cache, `_build_simple_maps` -- then runs unmodified on the filled grid and must
produce the model the reference builds itself.
"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.reference

CASES = [(12345, {}), (7, {"carve_subblock_roads": True}), (14, {"width": 150, "height": 110, "carve_subblock_roads": True})]


@pytest.mark.parametrize("seed,kw", CASES, ids=[f"s{s}" for s, _ in CASES])
def test_fill_model_from_planes_rebuilds_the_reference_model(seed, kw):
    from oracle.refharness import harness as Hn
    from trafficsimulation_b200.adaptor import fill_model_from_planes
    ref = Hn.load_reference()
    want = Hn.run_layout(seed, keep_model=True, **kw)
    assert want["crashed"] is None
    m_ref = want["model"]

    cm, Defaults = ref.cm, ref.Defaults
    originals = {n: getattr(cm.CityModel, n) for n in Hn.LAYOUT_PASSES}

    def filled_first_pass(self):
        fill_model_from_planes(self, want["final"], want["links"], want["hbands"], want["vbands"])

    old = (Defaults.ENABLE_TRAFFIC, Defaults.RAIN_ENABLED)
    Defaults.ENABLE_TRAFFIC = Defaults.RAIN_ENABLED = False
    try:
        for n in Hn.LAYOUT_PASSES:
            setattr(cm.CityModel, n, (filled_first_pass if n == "_place_thick_wall" else (lambda self, *a, **k: None)))
        with Hn._in_tmpdir():
            random.seed(seed)
            m_new = cm.CityModel(seed=seed, **kw)
    finally:
        for n, f in originals.items():
            setattr(cm.CityModel, n, f)
        Defaults.ENABLE_TRAFFIC, Defaults.RAIN_ENABLED = old

    # 1. the grid: same planes when extracted with the harness's own extractor
    got = Hn.extract_planes(m_new)
    for f in ("cell_type", "dirs", "aux", "block_id"):
        assert np.array_equal(got[f], want["final"][f]), f
    # 2. the derived maps the reference computed itself on the filled grid
    maps = Hn.extract_simple_maps(m_new)
    for k, v in want["maps"].items():
        assert np.array_equal(maps[k], v), k
    # 3. link tables and trackers
    links = Hn.extract_light_links(m_new)
    for k in ("lights", "ctrl", "incoming"):
        assert np.array_equal(links[k], want["links"][k]), k
    pos = lambda cells: sorted(c.position for c in cells)
    for name in ("block_entrances", "highway_entrances", "highway_exits", "controlled_roads", "traffic_lights"):
        assert pos(getattr(m_new, name)) == pos(getattr(m_ref, name)), name
    assert m_new._intersection_cells == m_ref._intersection_cells
    assert m_new._ring_road_cells == m_ref._ring_road_cells
    assert len(m_new._blocks_data) == len(m_ref._blocks_data)
    for a, b in zip(m_new._blocks_data, m_ref._blocks_data):
        assert a["block_id"] == b["block_id"] and a["block_type"] == b["block_type"]
        assert sorted(a["region"]) == sorted(b["region"]) and sorted(a["ring"]) == sorted(b["ring"])
    # 4. what the reference built on top of the grid: light groups and city blocks
    groups = lambda m: sorted(sorted(tl.position for tl in g.traffic_lights) for g in m.intersection_light_groups)
    assert groups(m_new) == groups(m_ref)
    per_cell = lambda m: sorted((c.position, c.cell_type, tuple(c.directions), c.road_type, c.block_id, c.block_type, c.light is None,
                                 c.highway_orientation) for col in m.grid._cells for cell in col for c in cell[:1])
    assert per_cell(m_new) == per_cell(m_ref)
