"""Host logic of the sharded tick (CPU only): which shard owns / simulates which light group, window-local tables."""
import numpy as np
import pytest

from oracle import oracle as O
from golden_util import tick_fixtures, load_ticks
from trafficsimulation_b200.sharded import ShardPlan
from trafficsimulation_b200.sharded_traffic import group_row_extents, shard_light_tables, _csr_take


def _tables():
    r = load_ticks(tick_fixtures()[0])
    return r, O.light_tables_from_reference(r["links_lights"], r["links_ctrl"], r["groups"])


def test_csr_take_keeps_the_selected_rows():
    off = np.array([0, 2, 2, 5, 6]); val = np.array([10, 11, 20, 21, 22, 30])
    o, v = _csr_take(off, val, np.array([True, False, True, False]))
    assert o.tolist() == [0, 2, 5] and v.tolist() == [10, 11, 20, 21, 22]
    o, v = _csr_take(off, val, np.zeros(4, bool))
    assert o.tolist() == [0] and len(v) == 0


def test_group_extents_cover_every_cell_of_the_group():
    r, tabs = _tables()
    W = r["W"]
    lo, hi, first = group_row_extents(tabs, W)
    for g, grp in enumerate(r["groups"]):
        rows = np.concatenate([np.asarray(grp[k]).reshape(-1) for k in ("cluster", "lights", "ns_in", "ew_in")]) // W
        assert lo[g] <= rows.min() and hi[g] >= rows.max()
        assert first[g] == np.asarray(grp["cluster"])[0] // W
    assert (hi - lo + 1).max() <= 40      # the default city's groups are ~30 rows tall: halo 128 leaves room


@pytest.mark.parametrize("n_shards,halo", [(2, 64), (3, 60)])
def test_every_group_has_one_owner_inside_its_window(n_shards, halo):
    r, tabs = _tables()
    W, H = r["W"], r["H"]
    plan = ShardPlan(H, n_shards, halo)
    lo, hi, first = group_row_extents(tabs, W)
    owner = np.searchsorted(np.asarray(plan.own_hi), first, side="right")
    assert owner.min() >= 0 and owner.max() < n_shards
    total = 0
    for s in range(n_shards):
        sel = (lo >= plan.win_lo[s]) & (hi < plan.win_hi[s])
        assert sel[owner == s].all()
        total += int((owner == s).sum())
        lt = shard_light_tables(tabs, W, plan.win_lo[s], plan.win_hi[s] - plan.win_lo[s], sel)
        assert lt["n_groups"] == int(sel.sum())
        base = plan.win_lo[s] * W
        for j, g in enumerate(np.flatnonzero(sel)):     # local tables translate back to the global ones
            for k in ("g_nsin", "g_ewin", "g_cl"):
                mine = lt[k][lt[k + "_off"][j]:lt[k + "_off"][j + 1]].astype(np.int64) + base
                ref = tabs[k][tabs[k + "_off"][g]:tabs[k + "_off"][g + 1]]
                assert np.array_equal(mine, ref)
            for k in ("g_all", "g_ns", "g_ew"):
                mine = lt[k][lt[k + "_off"][j]:lt[k + "_off"][j + 1]]
                assert np.array_equal(mine, tabs[k][tabs[k + "_off"][g]:tabs[k + "_off"][g + 1]])
                for l in mine:                           # the lights of a selected group lie inside the window
                    cells = lt["tl_cells"][lt["tl_off"][l]:lt["tl_off"][l + 1]]
                    assert (cells >= 0).all()
    assert total == tabs["n_groups"]
