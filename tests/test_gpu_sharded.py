"""Row-band shards on ONE GPU (all shards local to the process): the N-shard city must equal the 1-shard city
byte for byte -- planes, derived maps and light link tables (SURVEY.md §4 "multi-GPU test", §8e)."""
import numpy as np
import pytest

from golden_util import layout_fixtures, load, PLANES, MAPS

pytestmark = pytest.mark.gpu


def _single(cfgd, carve, hb, vb, tz, tc, te):
    from trafficsimulation_b200.layout import GpuCityLayout
    gc = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
    gc.set_bands(hb, vb)
    gc.generate(tz, tc, te)
    return gc


def _compare(ref, sh):
    want, got = ref.planes_host(), sh.planes_host()
    for f in PLANES:
        bad = np.argwhere(want[f] != got[f])
        assert len(bad) == 0, (f, len(bad), [(int(y), int(x), int(want[f][y, x]), int(got[f][y, x])) for y, x in bad[:8]])
    wm, gm = ref.maps_host(), sh.maps_host()
    for k in MAPS:
        assert np.array_equal(wm[k], gm[k]), k
    wl, gl = ref.light_links_host(), sh.light_links_host()
    for k in ("lights", "ctrl", "incoming"):
        assert np.array_equal(np.asarray(wl[k], np.int64), np.asarray(gl[k], np.int64)), ("links", k, len(wl[k]), len(gl[k]))


@pytest.mark.parametrize("n_shards", [2, 3])
def test_sharded_equals_single_on_reference_fixture(n_shards):
    from trafficsimulation_b200.sharded import ShardedCityLayout
    path = [p for p in layout_fixtures() if "400x300_carve" in p][0]
    g = load(path)
    cfgd = dict(g["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads")
    ref = _single(cfgd, carve, g["hbands"], g["vbands"], g["tape_zone"], g["tape_carve"], g["tape_entrance"])
    sh = ShardedCityLayout(n_shards, halo=40, carve_subblock_roads=carve, **cfgd)
    sh.set_bands(g["hbands"], g["vbands"])
    sh.generate(g["tape_zone"], g["tape_carve"], g["tape_entrance"])
    assert sh.n_blocks == ref.n_blocks
    _compare(ref, sh)
    got = sh.planes_host()
    for f in PLANES:   # and the reference's own planes
        assert np.array_equal(got[f], g[f]), ("golden", f)


@pytest.mark.parametrize("size,n_shards,carve,global_reach", [((768, 1024), 4, True, False), ((1024, 1024), 8, True, False),
                                                              ((1000, 600), 2, False, False), ((768, 1024), 4, True, True),
                                                              ((512, 1536), 3, False, True)])
def test_sharded_equals_single_synthetic(size, n_shards, carve, global_reach):
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.sharded import ShardedCityLayout
    W, H = size
    seed = 700 + n_shards
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    tz, te = tapes.synth_zone_tape(seed, cap), np.zeros(cap, np.int32)
    tc = None
    if carve:
        g0 = GpuCityLayout(width=W, height=H, carve_subblock_roads=True)
        g0.set_bands(hb, vb)
        g0._build_roads_and_sidewalks()
        n, table = g0.label_nothing()
        tc = tapes.synth_carve_tape(seed, table.cpu().numpy())
        assert tc[:, 1].sum() > 0
    ref = _single(dict(width=W, height=H), carve, hb, vb, tz, tc, te)
    sh = ShardedCityLayout(n_shards, halo=64, width=W, height=H, carve_subblock_roads=carve, global_reach=global_reach)
    sh.set_bands(hb, vb)
    sh.generate(tz, tc, te)
    assert sh.n_blocks == ref.n_blocks
    assert sh.reach_rounds >= (2 if global_reach else 1) and sh.dead_end_rounds >= 1
    _compare(ref, sh)


def test_halo_too_small_is_refused_loudly():
    from trafficsimulation_b200 import _lib, tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout
    W = H = 512
    hb, vb = tapes.synth_bands(5, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    sh = ShardedCityLayout(8, halo=3, width=W, height=H)   # 7 cuts, 3 halo rows: some block must straddle a cut
    sh.set_bands(hb, vb)
    with pytest.raises(_lib.TsimError):
        sh.generate(tapes.synth_zone_tape(5, cap), None, None)


@pytest.mark.parametrize("size,n_shards,carve", [((768, 1024), 4, True), ((1024, 2048), 8, True), ((1000, 1200), 2, False), ((640, 1800), 3, True)])
def test_lean_shards_equal_single(size, n_shards, carve):
    """No halo exchange at all (sharded.py, lean mode): every shard computes its halo itself, the digests of the rows around
    every cut are compared after every pass; the own rows must be the single-device city byte for byte."""
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.sharded import ShardedCityLayout
    W, H = size
    seed = 900 + n_shards
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    tz, te = tapes.synth_zone_tape(seed, cap), np.zeros(cap, np.int32)
    tc = None
    if carve:
        g0 = GpuCityLayout(width=W, height=H, carve_subblock_roads=True)
        g0.set_bands(hb, vb)
        g0._build_roads_and_sidewalks()
        n, table = g0.label_nothing()
        tc = tapes.synth_carve_tape(seed, table.cpu().numpy())
    ref = _single(dict(width=W, height=H), carve, hb, vb, tz, tc, te)
    sh = ShardedCityLayout(n_shards, halo=192, lean=True, width=W, height=H, carve_subblock_roads=carve)
    sh.set_bands(hb, vb)
    sh.generate(tz, tc, te)
    assert sh.n_blocks == ref.n_blocks
    _compare(ref, sh)
    # the digests are those of real bytes: every pass left a non-zero digest on both sides of a cut
    d = sh._dig[1].cpu().numpy()
    assert (d[[0, 2, 3, 4, 5, 6]][:, :2] != 0).all()


def test_lean_shards_refuse_a_halo_that_is_too_small():
    """With a halo of 10 rows, all of them verified, the zone that a shard cannot compute correctly (blocks cut by the window
    edge, their entrances, the lights next to them) lies in the compared rows: the digests differ (or a block straddles the
    whole halo) and the call raises instead of returning a different city."""
    from trafficsimulation_b200 import _lib, tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout
    W, H = 768, 1024
    hb, vb = tapes.synth_bands(31, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    sh = ShardedCityLayout(4, halo=10, lean=True, verify=10, width=W, height=H)
    sh.set_bands(hb, vb)
    with pytest.raises(_lib.TsimError):
        sh.generate(tapes.synth_zone_tape(31, cap), None, None)
