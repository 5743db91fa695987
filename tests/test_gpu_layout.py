"""GPU parity: every layout pass of libtsim.so vs the C oracle (lock-step) and vs the reference fixtures.

All calls go through the C ABI (trafficsimulation_b200.layout -> ctypes -> libtsim.so).
"""
import os

import numpy as np
import pytest

from golden_util import layout_fixtures, load, PLANES, MAPS

pytestmark = pytest.mark.gpu


def _mk(cfgd, hb, vb):
    from oracle import oracle as O
    from trafficsimulation_b200.layout import GpuCityLayout
    cfgd = dict(cfgd)
    carve = cfgd.pop("carve_subblock_roads", False)
    oc = O.OracleCity(O.make_cfg(fast_reach=1, **cfgd), hb, vb)
    gc = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
    gc.set_bands(hb, vb)
    return oc, gc, carve


def _cmp(stage, oc, gc, fields=PLANES, aux_mask=0xFF):
    got = gc.planes_host()
    want = oc.planes()
    for f in fields:
        w, g = want[f], got[f]
        if f == "aux":
            w, g = w & aux_mask, g & aux_mask
        bad = np.argwhere(w != g)
        assert len(bad) == 0, (stage, f, len(bad), [(int(y), int(x), int(w[y, x]), int(g[y, x])) for y, x in bad[:8]])


def lockstep(cfgd, hb, vb, tape_zone, tape_carve, tape_entrance):
    oc, gc, carve = _mk(cfgd, hb, vb)
    fwd = cfgd.get("forward_traffic_light_range", False)
    oc.frame(); oc.roads()
    gc._place_thick_wall(); gc._place_sidewalk_inner_ring(); gc._clear_interior(); gc._build_roads_and_sidewalks()
    _cmp("roads", oc, gc, ("cell_type", "dirs", "aux"))
    if carve:
        n, table = gc.label_nothing()
        ob = oc.nothing_blobs()
        assert n == len(ob)
        assert np.array_equal(table.cpu().numpy(), ob), "blob table"
        oc.carve(tape_carve)
        gc._carve_subblock_roads(tape_carve)
        _cmp("carve", oc, gc, ("cell_type", "dirs", "aux"))
    nb = oc.zones(tape_zone)
    gc._flood_fill_blocks_storing_data(tape_zone)
    assert gc.n_blocks == nb
    _cmp("zones", oc, gc)
    oc.dead_ends(); gc._eliminate_dead_ends()
    _cmp("dead_ends", oc, gc)
    oc.upgrade_r2(); gc._upgrade_r2_to_intersections()
    _cmp("upgrade_r2", oc, gc)
    ent = oc.entrances_pass(tape_entrance)
    gc._final_place_block_entrances(tape_entrance)
    _cmp("entrances", oc, gc)
    assert np.array_equal(gc.entrances[:nb].cpu().numpy(), ent), "entrance table"
    oc.validate_dirs(); oc.entrance_dirs()
    gc._remove_invalid_intersection_directions(); gc._add_entrance_directions()
    _cmp("fix_dirs", oc, gc)
    links = oc.lights()
    gc._add_traffic_lights()
    _cmp("lights", oc, gc)
    gl = gc.light_links_host()
    for k in ("lights", "ctrl", "incoming") + (("outgoing",) if fwd else ()):
        assert np.array_equal(gl[k], links[k]), ("links", k, len(gl[k]), len(links[k]))
    gc._build_simple_maps()
    om, gm = oc.simple_maps(), gc.maps_host()
    for k in MAPS:
        assert np.array_equal(om[k], gm[k]), k
    return oc, gc


@pytest.mark.parametrize("path", layout_fixtures(), ids=lambda p: os.path.basename(p)[7:-4])
def test_gpu_matches_oracle_and_golden(path):
    g = load(path)
    cfgd = g["meta"]["cfg"]
    oc, gc = lockstep(cfgd, g["hbands"], g["vbands"], g["tape_zone"], g["tape_carve"], g["tape_entrance"])
    got = gc.planes_host()
    for f in PLANES:
        assert np.array_equal(got[f], g[f]), ("golden", f)
    gm = gc.maps_host()
    for k in MAPS:
        assert np.array_equal(gm[k], g[k]), ("golden", k)
    gl = gc.light_links_host()
    for k in ("lights", "ctrl", "incoming") + (("outgoing",) if cfgd["forward_traffic_light_range"] else ()):
        assert np.array_equal(gl[k], g["links_" + k]), ("golden links", k)


def test_unsupported_range_is_refused_loudly():
    """traffic_light_range beyond what the 5-bit scan-length fields hold must raise, not truncate."""
    from trafficsimulation_b200 import _lib
    from trafficsimulation_b200.layout import GpuCityLayout
    g = load(layout_fixtures()[0])
    gc = GpuCityLayout(traffic_light_range=31)
    gc.set_bands(g["hbands"], g["vbands"])
    gc._build_roads_and_sidewalks()
    with pytest.raises(_lib.TsimError, match="UNSUPPORTED"):
        gc._add_traffic_lights()


@pytest.mark.parametrize("mode", ["rows", "rows_legacy", "rows_tma2048", "rows_tma8192", "lut", "none"])
def test_frame_pass_kernels_agree(mode, monkeypatch):
    """The kernels of the frame + roads pass (pattern rows copied through registers, eight rows or one row per thread; pattern rows
    as TMA bulk copies with two tile sizes; class look-up; closed form per cell) write the same planes."""
    if mode.startswith("rows_tma"):
        monkeypatch.setenv("TSIM_FRAME_TMA", mode[8:])
        mode = "rows"
    if mode == "rows_legacy":
        monkeypatch.setenv("TSIM_FRAME_COPY", "legacy")
        mode = "rows"
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    W, H = 1024, 640
    hb, vb = tapes.synth_bands(77, width=W, height=H, ring_road_type="R2")
    gc = GpuCityLayout(width=W, height=H, frame_tables=mode)
    gc.set_bands(hb, vb)
    gc._build_roads_and_sidewalks()
    oc = O.OracleCity(O.make_cfg(width=W, height=H), hb, vb)
    oc.frame(); oc.roads()
    _cmp("roads/" + mode, oc, gc, ("cell_type", "dirs", "aux"))


@pytest.mark.parametrize("n_alt", ["0", "1", "2"])
def test_reach_fallback_kernel_gives_the_same_city(n_alt, monkeypatch):
    """The reachability closure enqueues a fixed number of alternations and leaves the rest to the persistent cooperative
    kernel; with 0 / 1 / 2 enqueued alternations that kernel does (almost) all the work and the result must not change."""
    monkeypatch.setenv("TSIM_REACH_ALTERNATIONS", n_alt)
    monkeypatch.setenv("TSIM_LIGHTS_STAGES", "3")   # no local searches: every undecided `leads_to` is read from the planes
    path = [p for p in layout_fixtures() if "s7_carve" in p][0]
    g = load(path)
    lockstep(g["meta"]["cfg"], g["hbands"], g["vbands"], g["tape_zone"], g["tape_carve"], g["tape_entrance"])


@pytest.mark.parametrize("name", ["default12345", "s14_150x110_carve", "s24_400x300_carve"])
def test_general_labelling_equals_product_short_cut(name, monkeypatch):
    """The Nothing cells of a fresh layout are (rows without a band) x (columns without one); k_ccl.cu then reads the components
    off the row / column runs instead of running the union-find.  With the short cut switched off the general path must give
    the same tables, ids and city (the lock-step run compares them with the oracle's)."""
    monkeypatch.setenv("TSIM_CCL_PRODUCT", "0")
    g = load([p for p in layout_fixtures() if name in p][0])
    lockstep(g["meta"]["cfg"], g["hbands"], g["vbands"], g["tape_zone"], g["tape_carve"], g["tape_entrance"])


def test_product_short_cut_is_taken_and_refused():
    """Which path labelled: a fresh layout is a product (short cut), a carved one is not (union-find)."""
    import torch
    g = load([p for p in layout_fixtures() if "s24_400x300_carve" in p][0])
    assert g["tape_carve"][:, 1].sum() > 0   # something does get carved
    oc, gc, carve = _mk(g["meta"]["cfg"], g["hbands"], g["vbands"])
    gc._build_roads_and_sidewalks()
    gc.label_nothing()
    head = lambda: gc.workspace[: 64 * 4].view(torch.int32).cpu().numpy()
    assert head()[1] == 1 and head()[2] * head()[3] == int(gc.flags[2].item())
    gc._carve_subblock_roads(g["tape_carve"])
    gc.label_nothing()
    assert head()[1] == 0 and head()[0] > 0   # not a product any more: the runs were built


LEADS_TO_FIXTURES = ("default12345", "s11_ringR1", "s15_96x128", "s16_64_carve", "s24_400x300_carve", "s13_noopt")


@pytest.mark.parametrize("stages", ["1", "2", "3"])
@pytest.mark.parametrize("name", LEADS_TO_FIXTURES)
def test_leads_to_search_stages_agree(name, stages, monkeypatch):
    """`leads_to` beyond the lane itself is settled by Z witnesses, then window closures, then the reachability planes of the
    whole window (k_lights.cu).  With the first, the second or both local stages switched off the later ones must build the
    same city: the fixtures hold ring corners (window closures), opposite carriageways (Z witnesses) and unreachable lanes
    (only the planes can say "False")."""
    monkeypatch.setenv("TSIM_LIGHTS_STAGES", stages)
    g = load([p for p in layout_fixtures() if name in p][0])
    oc, gc = lockstep(g["meta"]["cfg"], g["hbands"], g["vbands"], g["tape_zone"], g["tape_carve"], g["tape_entrance"])
    got = gc.planes_host()
    for f in PLANES:
        assert np.array_equal(got[f], g[f]), ("golden", f)


def test_leads_to_stage_counters():
    """What the stages leave over on a synthetic city: a handful of candidates reach the window closures, none the planes."""
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    size, seed = 1024, 7
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    gc = GpuCityLayout(width=size, height=size)
    gc.set_bands(hb, vb)
    gc.generate(tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    left = gc.lights_undecided()
    assert 0 < left["after_z_witness"] < 64 and left["after_window"] == 0, left


SYNTH = [
    (101, dict(width=512, height=512), True),
    (102, dict(width=1024, height=768, ring_road_type="R1"), True),
    (103, dict(width=1000, height=1000), False),            # width not a multiple of 16: scalar paths
    (104, dict(width=2048, height=2048, optimized_intersections=False), True),
    (105, dict(width=640, height=512, block_entrance_road_level=1), True),     # road-level filter of the entrance pass
    (106, dict(width=512, height=640, block_entrance_road_level=2, ring_road_type="R3"), False),
    (107, dict(width=768, height=512, forward_traffic_light_range=True, forward_traffic_light_range_intersections="Include in Range"), True),
    (108, dict(width=512, height=512, forward_traffic_light_range=True, forward_traffic_light_range_intersections="Include as Extra",
               traffic_light_range=4), False),
    (109, dict(width=512, height=512, forward_traffic_light_range=True), True),
]


@pytest.mark.parametrize("seed,kw,carve", SYNTH, ids=[f"synth{s}" for s, _, _ in SYNTH])
def test_gpu_matches_oracle_synthetic(seed, kw, carve):
    """Sizes the Python reference cannot reach: synthetic tapes, CUDA vs the (pinned) C oracle."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    band_kw = {k: v for k, v in kw.items() if k in ("width", "height", "ring_road_type")}
    hb, vb = tapes.synth_bands(seed, **band_kw)
    cfgd = dict(kw, carve_subblock_roads=carve)
    # carve tape needs the blob table: take it from the oracle (the lock-step run re-checks it against the GPU)
    tape_carve = None
    if carve:
        o0 = O.OracleCity(O.make_cfg(**{k: v for k, v in cfgd.items() if k != "carve_subblock_roads"}), hb, vb)
        o0.frame(); o0.roads()
        tape_carve = tapes.synth_carve_tape(seed, o0.nothing_blobs())
        assert tape_carve[:, 1].sum() > 0
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    tape_zone = tapes.synth_zone_tape(seed, cap)
    lockstep(cfgd, hb, vb, tape_zone, tape_carve, np.zeros(cap, np.int32))
