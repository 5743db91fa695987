"""Host-side light-group tables against the LIVE reference (CPU): lane lists incl. the out lanes (intersection_light_group.py:141-154)
and the neighbour links of populate_links (:175-242).  scipy's labelling stands in for tsim_label_mask (same numbering: raster order
of a cluster's first cell)."""
import numpy as np
import pytest

pytestmark = pytest.mark.reference

CASES = [
    dict(seed=12345),
    dict(seed=7),
    dict(seed=9),
    dict(seed=14, layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
    dict(seed=13, layout_kwargs=dict(optimized_intersections=False)),
    dict(seed=11, layout_kwargs=dict(ring_road_type="R1")),
    dict(seed=22, layout_kwargs=dict(forward_traffic_light_range=True)),
    dict(seed=3, layout_kwargs=dict(width=120, height=140, carve_subblock_roads=True, subblock_roads_have_intersections=False)),
]


def tables_from_reference_planes(lay, W, H):
    from scipy import ndimage
    from trafficsimulation_b200.light_groups import build_light_tables
    fin, links = lay["final"], lay["links"]
    ever = (fin["aux"].reshape(H, W) & 0x40) != 0
    lab, nc = ndimage.label(ever, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    table = np.zeros((nc, 6), np.int32)
    for k, sl in enumerate(ndimage.find_objects(lab)):
        table[k, :4] = (sl[1].start, sl[0].start, sl[1].stop - 1, sl[0].stop - 1)
    lights = links["lights"]

    def csr(pairs):
        idx = np.searchsorted(lights, pairs[:, 0]) if len(pairs) else np.zeros(0, np.int64)
        off = np.zeros(len(lights) + 1, np.int64)
        np.add.at(off, idx + 1, 1)
        return np.cumsum(off), pairs[np.argsort(idx, kind="stable"), 1] if len(pairs) else np.zeros(0, np.int32)
    c_off, c_cell = csr(links["ctrl"]); i_off, i_cell = csr(links["incoming"]); o_off, o_cell = csr(links["outgoing"])
    tabs = build_light_tables(W, H, fin["cell_type"].reshape(H, W), fin["dirs"].reshape(H, W), lab.astype(np.int32), table, lights,
                              c_off, c_cell, i_off, i_cell, o_off, o_cell)
    return tabs, lab


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c['seed']}")
def test_group_tables_and_neighbour_links_match_reference(case):
    from oracle.refharness import harness as H, ticks
    from trafficsimulation_b200.light_groups import groups_as_cell_lists, neighbor_links
    lay = H.run_layout(case["seed"], enable_traffic=True, enable_rain=False, keep_model=True, **case.get("layout_kwargs", {}))
    model = lay["model"]
    assert model is not None, lay["crashed"]
    W, Hh = model.width, model.height
    ref_groups, _ = ticks.light_group_tables(model)
    tabs, lab = tables_from_reference_planes(lay, W, Hh)
    mine = groups_as_cell_lists(tabs, lay["links"]["lights"])
    assert len(mine) == len(ref_groups) > 0
    for i, (a, b) in enumerate(zip(mine, ref_groups)):
        for k in a:
            assert np.array_equal(a[k], b[k]), ("group table", i, k)
    order = np.argsort([int(g["creation_rank"][0]) for g in ref_groups])      # the reference's creation order (a Python set's iteration order)
    got = neighbor_links(W, Hh, lay["final"]["cell_type"], lab, tabs, lay["links"]["lights"], lay["hbands"], lay["vbands"], creation_order=order)
    want = np.stack([g["nbr"] for g in ref_groups])
    assert np.array_equal(got["nbr"], want), np.flatnonzero((got["nbr"] != want).any(1))[:10]
    assert (want >= 0).any()
    if not got["order_dependent"]:   # then any order gives the same links: the canonical one in particular
        assert np.array_equal(neighbor_links(W, Hh, lay["final"]["cell_type"], lab, tabs, lay["links"]["lights"], lay["hbands"], lay["vbands"])["nbr"], want)
