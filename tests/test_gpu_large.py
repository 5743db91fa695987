"""Full-size cities (BASELINE.json configs[1] 4096^2 and configs[2] 16384^2), the sizes bench.py times: the finished city is
compared with the C oracle byte for byte (planes, maps, link tables: a few seconds of CPU at 4096^2, about a minute at
16384^2), and checked for size-independent properties computed on the device with plain torch ops."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T = dict(RES=0, EMPTY=5, NOTHING=6, SIDEWALK=7, WALL=8, R1=9, R2=10, R3=11, INTER=12, HIN=13, HOUT=14, TL=15, CR=17, BE=19)
ROAD_LIKE = [9, 10, 11, 12, 13, 14, 19]


def _city(size, seed=4096):
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    sh = ShardedCityLayout(1, width=size, height=size, carve_subblock_roads=True)
    sh.set_bands(hb, vb)
    dev = sh.shards[0].device
    tz = torch.from_numpy(tapes.synth_zone_tape(seed, sh.global_cap)).to(dev)
    te = torch.zeros(sh.global_cap, dtype=torch.int32, device=dev)
    tc = sh.synth_carve_tapes(seed)
    return sh, tz, tc, te


def _isin(t, codes):
    m = torch.zeros_like(t, dtype=torch.bool)
    for c in codes:
        m |= t == c
    return m


def _shift(m, dy, dx):
    """m shifted so that out[y, x] = m[y + dy, x + dx] (False outside)."""
    out = torch.zeros_like(m)
    H, W = m.shape
    ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
    xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
    out[yd, xd] = m[ys, xs]
    return out


@pytest.mark.parametrize("size", [4096, 16384])
def test_full_size_city_properties(size):
    sh, tz, tc, te = _city(size)
    L = sh.shards[0]
    sh.generate(tz, tc, te)
    torch.cuda.synchronize()
    first = {k: getattr(L, k).clone() for k in ("cell_type", "dirs", "aux", "block_id")}
    first_maps = {k: v.clone() for k, v in L.maps.items()}
    n_blocks, n_lights = sh.n_blocks, int(L.flags[3].item())
    # 1. determinism: a second generation from the same tapes gives the same bytes (no order-dependent atomics)
    sh.generate(tz, tc, te)
    torch.cuda.synchronize()
    for k, v in first.items():
        assert torch.equal(getattr(L, k), v), f"non-deterministic plane {k}"
    for k, v in first_maps.items():
        assert torch.equal(L.maps[k], v), f"non-deterministic map {k}"
    H = W = size
    ct = L.cell_type.view(H, W)
    dirs = L.dirs.view(H, W).to(torch.int32) & 0xFFFF
    aux = L.aux.view(H, W)
    bid = L.block_id.view(H, W)
    # 2. the passes ran to completion: nothing unzoned, no dead end left (city_model.py:811-840)
    assert int((ct == T["NOTHING"]).sum()) == 0
    road = _isin(ct, ROAD_LIKE)
    road_then = road | (ct == T["CR"])   # a ControlledRoad was a road cell when the dead ends were pruned
    nb = _shift(road_then, 1, 0).int() + _shift(road_then, -1, 0).int() + _shift(road_then, 0, 1).int() + _shift(road_then, 0, -1).int()
    assert int((_isin(ct, [T["R2"], T["R3"], T["INTER"]]) & (nb < 2)).sum()) == 0, "dead ends survive"
    # 3. derived maps agree with the planes (city_model.py:2151-2199)
    assert torch.equal(L.maps["is_road_map"].view(H, W) != 0, road)
    assert torch.equal(L.maps["intersection_map"].view(H, W) != 0, ct == T["INTER"])
    assert torch.equal(L.maps["allowed_dirs_map"].view(H, W).to(torch.int32), dirs & 0xF)
    # the direction word is self-consistent: mask bits == entries of the ordered list
    n = (dirs >> 12) & 7
    mask = torch.zeros_like(dirs)
    for i in range(4):
        mask |= torch.where(n > i, 1 << ((dirs >> (4 + 2 * i)) & 3), 0)
    assert torch.equal(mask, dirs & 0xF)
    # 4. block ids: exactly the zone cells and the entrances carry one, ids are 1..n_blocks, all present
    zone = ct <= T["EMPTY"]
    assert torch.equal(bid > 0, zone | ((ct == T["BE"]) & (bid > 0)))
    assert int(bid.max()) == n_blocks
    assert int(torch.unique(bid[zone]).numel()) == n_blocks
    # every entrance sits on the ring of its block: a 4-neighbour is a zone cell with the same id
    be = ct == T["BE"]
    touch = torch.zeros_like(be)
    for dy, dx in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        nbid = torch.zeros_like(bid)
        nz = _shift(zone, dy, dx)
        sh_bid = torch.zeros_like(bid)
        ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
        xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
        sh_bid[yd, xd] = bid[ys, xs]
        touch |= nz & (sh_bid == bid)
    assert int((be & ~touch).sum()) == 0
    # 5. lights: numbered in ascending cell order, every light cell is a TrafficLight, every controlled cell a ControlledRoad
    t = L._link_tensors
    lights = t["light_cell"][:n_lights].to(torch.int64)
    assert n_lights == int((ct == T["TL"]).sum())
    assert bool((lights[1:] > lights[:-1]).all())
    flat = L.cell_type
    assert bool((flat[lights] == T["TL"]).all())
    n_ctrl = int(t["ctrl_off"][n_lights].item())
    ctrl = t["ctrl_cell"][:n_ctrl].to(torch.int64)
    assert bool((flat[ctrl] == T["CR"]).all())
    assert bool(((L.aux[ctrl] & 0x80) != 0).all())                       # controlled_road.light = tl
    orig = (aux[ct == T["CR"]] & 0x1F).to(torch.int64)
    assert bool(_isin(orig, [T["R1"], T["R2"], T["R3"], T["HIN"], T["HOUT"], T["BE"]]).all())   # remembered original type
    n_inc = int(t["inc_off"][n_lights].item())
    inc = t["inc_cell"][:n_inc].to(torch.int64)
    assert bool(_isin(flat[inc], [T["R1"], T["R2"], T["R3"], T["HIN"], T["HOUT"], T["BE"], T["CR"]]).all())
    # every ControlledRoad has an arrow into an Intersection (city_model.py:1446-1452)
    inter = ct == T["INTER"]
    into = ((dirs & 1) != 0) & _shift(inter, 1, 0) | ((dirs & 2) != 0) & _shift(inter, 0, 1) | ((dirs & 4) != 0) & _shift(inter, -1, 0) | \
           ((dirs & 8) != 0) & _shift(inter, 0, -1)
    assert int(((ct == T["CR"]) & ~into).sum()) == 0


def test_sharded_equals_single_at_4096():
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout
    size, seed = 4096, 4096
    sh1, tz, tc, te = _city(size, seed)
    sh1.generate(tz, tc, te)
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    sh4 = ShardedCityLayout(4, halo=64, width=size, height=size, carve_subblock_roads=True)
    sh4.set_bands(hb, vb)
    tc4 = sh4.synth_carve_tapes(seed)
    sh4.generate(tz, tc4, te)
    assert sh4.n_blocks == sh1.n_blocks
    a, b = sh1.planes_host(), sh4.planes_host()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    ma, mb = sh1.maps_host(), sh4.maps_host()
    for k in ma:
        assert np.array_equal(ma[k], mb[k]), k


def _oracle_city(size, seed, tz, tc, te):
    """The C oracle on the inputs the GPU run consumed (band lists from the same generator, the same three tapes)."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    oc = O.OracleCity(O.make_cfg(width=size, height=size, fast_reach=1), hb, vb)
    oc.run_all(tz, tc, te, carve=True)
    return oc


@pytest.mark.parametrize("size", [4096, 16384])
def test_full_size_city_equals_oracle(size):
    """The benchmarked workload itself (seed 4096, carve + lights + maps) against the oracle: every plane, every derived map and
    the three link tables, bit for bit.  The carve tape is the one bench.py uses (drawn on the device from the blob table, one
    row per blob id); the oracle discovers the blobs itself, so a wrong table or id order cannot cancel out."""
    seed = 4096
    sh, tz, tc, te = _city(size, seed)
    L = sh.shards[0]
    sh.generate(tz, tc, te)
    torch.cuda.synchronize()
    oc = _oracle_city(size, seed, tz.cpu().numpy(), tc[0][:-1].cpu().numpy(), te.cpu().numpy())
    assert oc.n_blocks == sh.n_blocks
    got, want = L.planes_host(), oc.planes()
    for f in ("cell_type", "dirs", "aux", "block_id"):
        same = np.array_equal(got[f], want[f])
        if not same:
            bad = np.argwhere(got[f] != want[f])
            raise AssertionError((f, len(bad), [(int(y), int(x), int(want[f][y, x]), int(got[f][y, x])) for y, x in bad[:8]]))
    del got
    gm, om = L.maps_host(), oc.simple_maps()
    for k in om:
        assert np.array_equal(gm[k], om[k]), k
    del gm, om
    gl, ol = L.light_links_host(), oc.links
    for k in ("lights", "ctrl", "incoming"):
        assert np.array_equal(gl[k], ol[k]), ("links", k, len(gl[k]), len(ol[k]))
