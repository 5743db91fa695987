"""GPU parity of the batched route planner (tsim_astar_batch through the C ABI) against the reference's golden vectors and
the C oracle."""
import glob
import os

import numpy as np
import pytest

from golden_util import load_astar

pytestmark = pytest.mark.gpu

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "astar_*.npz")))


def _queries(r):
    q = r["queries"]
    return np.stack([q[:, 0], q[:, 1], q[:, 2], q[:, 3], q[:, 4] | (q[:, 5] << 1) | (q[:, 6] << 2), np.full(len(q), 10), q[:, 7]], 1)


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[6:-4])
def test_gpu_astar_matches_reference_vectors(path):
    from trafficsimulation_b200.pathfinding import GpuAstar
    r = load_astar(path)
    planner = GpuAstar(r["W"], r["H"], r["occupancy"], r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
    got = planner.plan_cells(_queries(r))
    assert len(got) == len(r["paths"])
    for q, g, want in zip(r["queries"], got, r["paths"]):
        assert g.tolist() == list(want), tuple(q)
    # a deliberately short path buffer is grown, not truncated; small chunks give the same answers
    planner.chunk = 7
    again = planner.plan_cells(_queries(r)[:40], max_path=3)
    for g, want in zip(again, r["paths"][:40]):
        assert g.tolist() == list(want)
    sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in r["queries"][5])
    one = planner.astar(sx, sy, gx, gy, bool(ra), 10, bool(so), bool(ig), ms)
    assert [y * r["W"] + x for x, y in one] == list(r["paths"][5])


@pytest.mark.parametrize("mode,smem_cap", [("0", None), ("1", None), ("2", None), ("2", "64")],
                         ids=["per_thread", "per_cta", "per_cta_smem", "per_cta_smem_overflow"])
def test_gpu_astar_launch_forms(mode, smem_cap, monkeypatch):
    """The three launch forms of tsim_astar_batch (a query per thread; per CTA -- the default for small batches; per CTA with the open
    list in shared memory) give the reference's paths; so does the fall-back when an open list outgrows its shared memory."""
    from trafficsimulation_b200.pathfinding import GpuAstar
    monkeypatch.setenv("TSIM_ASTAR_MODE", mode)
    if smem_cap:
        monkeypatch.setenv("TSIM_ASTAR_SMEM_CAP", smem_cap)
    r = load_astar(FIXTURES[0])
    planner = GpuAstar(r["W"], r["H"], r["occupancy"], r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
    got = planner.plan_cells(_queries(r)[:300])
    for q, g, want in zip(r["queries"], got, r["paths"]):
        assert g.tolist() == list(want), tuple(q)


def test_gpu_astar_matches_oracle_on_a_synthetic_city():
    """A city the reference never saw (768 x 512 synthetic layout from the GPU pipeline), live maps from the tick planes."""
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.pathfinding import GpuAstar
    W, H, seed = 768, 512, 5
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = GpuCityLayout(width=W, height=H)
    city.set_bands(hb, vb)
    city.generate(tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    maps = city.maps_host()
    rng = np.random.default_rng(seed)
    road = np.flatnonzero(maps["is_road_map"].reshape(-1) == 1)
    occ = np.zeros(W * H, np.uint8); occ[rng.choice(road, len(road) // 10, replace=False)] = 1
    stop = np.zeros(W * H, np.uint8); stop[rng.choice(road, len(road) // 50, replace=False)] = 1
    dens = rng.random((H, W))
    planner = GpuAstar(W, H, occ.reshape(H, W), stop.reshape(H, W), city.maps["is_road_map"], city.maps["road_type_map"],
                       city.maps["allowed_dirs_map"], dens)
    ora = O.OracleAstar(occ.reshape(H, W), stop.reshape(H, W), maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"], dens)
    q = []
    for i in range(300):
        a = rng.choice(road)
        near = road[(np.abs(road % W - a % W) + np.abs(road // W - a // W)) <= (8 if i % 3 == 2 else 120)]
        b = rng.choice(near)
        flags, ms = [(0, 0x7FFFFFFF), (2, 0x7FFFFFFF), (4 | 2, 20), (1, 0x7FFFFFFF), (4, 6)][i % 5]
        q.append((a % W, a // W, b % W, b // W, flags, 10, ms))
    got = planner.plan(q)
    found = 0
    for (sx, sy, gx, gy, flags, aw, ms), g in zip(q, got):
        want = ora.query(sx, sy, gx, gy, bool(flags & 1), aw, bool(flags & 2), bool(flags & 4), ms)
        assert g == want, (sx, sy, gx, gy, flags, ms)
        found += bool(want)
    assert found > 100


@pytest.mark.parametrize("seed,shape,p_road,p_occ", [(1, (200, 200), 0.3, 0.1), (2, (64, 150), 0.6, 0.5), (4, (5, 90), 0.2, 1.0), (6, (300, 257), 0.32, 0.02)])
def test_gpu_density_map_is_bit_exact(seed, shape, p_road, p_occ):
    """tsim_density_map == the oracle == SciPy's float32 uniform_filter (CityModel._update_density_map), bit for bit."""
    from oracle import oracle as O
    from trafficsimulation_b200.pathfinding import GpuAstar
    rng = np.random.default_rng(seed)
    H, W = shape
    road = (rng.random(shape) < p_road).astype(np.uint8)
    occ = ((rng.random(shape) < p_occ) & (road == 1)).astype(np.uint8)
    planner = GpuAstar(W, H, occ, np.zeros(shape, np.uint8), road, road, np.full(shape, 15, np.uint8))
    got = planner.update_density().cpu().numpy()
    want = O.density_map(occ, road)
    assert got.dtype == np.float64 and np.array_equal(got, want.astype(np.float64))
    assert planner.maps["density_map"].data_ptr() == planner._maps_struct().density_map   # the planner reads this plane


def test_gpu_astar_spawn_rank_limit():
    """tsim_astar_query.spawn_rank_limit / tsim_astar_maps.spawn_rank: the k-th spawn of a tick plans as if the vehicles spawned
    after it were not on the grid yet -- ONE ranked batch on the full map == the plain searches on the maps with those cells
    cleared (device and C oracle)."""
    from oracle import oracle as O
    from trafficsimulation_b200.pathfinding import GpuAstar
    r = load_astar(FIXTURES[0])
    W, H = r["W"], r["H"]
    rng = np.random.default_rng(3)
    road = np.flatnonzero(r["is_road_map"].reshape(-1) == 1)
    born = rng.choice(road, 140, replace=False)            # more than the 7-bit rank field holds: the last ones share 127
    occ = np.array(r["occupancy"], np.uint8).reshape(-1).copy()
    occ[rng.choice(road, 300, replace=False)] = 1
    occ[born] = 1
    rank = np.zeros(W * H, np.uint8)
    rank[born] = np.minimum(np.arange(1, len(born) + 1), 127)
    static = (r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"])
    planner = GpuAstar(W, H, occ.reshape(H, W), r["stop_map"], *static, r["density"])
    planner.update(spawn_rank_map=rank.reshape(H, W))
    ks = [1, 2, 17, 60, 126]
    q = []
    for k in ks:
        for flags in (0, 2, 4, 6):
            for g in rng.choice(road, 6):
                q.append([born[k - 1] % W, born[k - 1] // W, g % W, g // W, flags, 10, 0x7FFFFFFF, k])
    q = np.array(q, np.int32)
    ranked = planner.plan_cells(q)
    differ = 0
    for k in ks:
        plain = occ.copy()
        plain[born[k:]] = 0
        sel = np.flatnonzero(q[:, 7] == k)
        planner.update(occupancy_map=plain.reshape(H, W), spawn_rank_map=None)
        alone = planner.plan_cells(q[sel, :7])
        ora = O.OracleAstar(plain.reshape(H, W), r["stop_map"], *static, r["density"])
        for i, p in zip(sel, alone):
            assert ranked[i].tolist() == p.tolist(), (k, q[i].tolist())
            sx, sy, gx, gy, fl = (int(v) for v in q[i, :5])
            want = ora.query(sx, sy, gx, gy, False, 10, bool(fl & 2), bool(fl & 4))
            assert [y * W + x for x, y in want] == p.tolist(), (k, q[i].tolist())
    planner.update(occupancy_map=occ.reshape(H, W), spawn_rank_map=None)
    full = planner.plan_cells(q[:, :7])
    differ = sum(a.tolist() != b.tolist() for a, b in zip(full, ranked))
    assert differ > 0                                      # the later spawns do change some of these routes
