"""Route planner on RANDOM small grids (not city-shaped): arbitrary arrow masks, road types, obstacles, awareness, step
limits -- the corner cases a generated city never produces.  The product's core (csrc/astar_core.cuh, host build) must
equal the C oracle on every query; where the live reference is present the oracle is pinned against Numba on the same grids."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("astar_rnd") / "astar_core_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "astar_core_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.host_astar.restype = C.c_int
    return lib


def random_case(seed):
    rng = np.random.default_rng(seed)
    W, H = int(rng.integers(3, 24)), int(rng.integers(3, 24))
    road = (rng.random((H, W)) < rng.choice([0.5, 0.8, 1.0])).astype(np.uint8)
    rtype = (rng.integers(0, 4, (H, W)) * road).astype(np.uint8)
    adirs = (rng.integers(0, 16, (H, W)) | rng.choice([0, 15], (H, W), p=[0.6, 0.4])).astype(np.uint8)
    occ = (rng.random((H, W)) < rng.choice([0.0, 0.1, 0.3])).astype(np.uint8)
    stop = (rng.random((H, W)) < rng.choice([0.0, 0.05, 0.2])).astype(np.uint8)
    dens = rng.random((H, W)) * rng.choice([0.0, 1.0])
    q = []
    for _ in range(40):
        sx, gx = rng.integers(0, W, 2)
        sy, gy = rng.integers(0, H, 2)
        flags = int(rng.integers(0, 8))
        q.append((int(sx), int(sy), int(gx), int(gy), flags, int(rng.integers(1, 6)), int(rng.choice([0x7FFFFFFF, 3, 9, 30]))))
    return W, H, occ, stop, road, rtype, adirs, dens, q


@pytest.mark.parametrize("seed", range(40))
def test_product_core_equals_oracle_on_random_grids(core, seed):
    W, H, occ, stop, road, rtype, adirs, dens, queries = random_case(seed)
    ora = O.OracleAstar(occ, stop, road, rtype, adirs, dens)
    out = np.zeros(W * H, np.int32)
    u8 = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_uint8))
    keep = [np.ascontiguousarray(a) for a in (occ, stop, road, rtype, adirs)]
    d64 = np.ascontiguousarray(dens, np.float64)
    for sx, sy, gx, gy, flags, aw, ms in queries:
        want = ora.query(sx, sy, gx, gy, bool(flags & 1), aw, bool(flags & 2), bool(flags & 4), ms)
        n = core.host_astar(W, H, *[u8(a) for a in keep], d64.ctypes.data_as(C.POINTER(C.c_double)), sx, sy, gx, gy, flags, aw, ms,
                            out.ctypes.data_as(C.POINTER(C.c_int32)), len(out), W * H)   # the reference's own capacity
        assert n >= 0 and out[:n].tolist() == [y * W + x for x, y in want], (seed, sx, sy, gx, gy, flags, aw, ms)


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(0, 40, 4))
def test_oracle_equals_reference_on_random_grids(seed):
    from oracle.refharness import stubs
    if hasattr(stubs, "install"):
        stubs.install()
    import Simulation.utilities.pathfinding  # noqa: F401
    ref = sys.modules["Simulation.utilities.pathfinding.astar_numba"].astar_numba
    W, H, occ, stop, road, rtype, adirs, dens, queries = random_case(seed)
    ora = O.OracleAstar(occ, stop, road, rtype, adirs, dens)
    i8 = lambda a: np.ascontiguousarray(a.astype(np.int8))
    maps = (i8(occ), i8(stop), i8(road), i8(rtype), i8(adirs))
    for sx, sy, gx, gy, flags, aw, ms in queries:
        want = [(int(x), int(y)) for x, y in ref(W, H, sx, sy, gx, gy, *maps, bool(flags & 1), aw, np.ascontiguousarray(dens, np.float64),
                                                 bool(flags & 2), bool(flags & 4), ms)]
        assert ora.query(sx, sy, gx, gy, bool(flags & 1), aw, bool(flags & 2), bool(flags & 4), ms) == want, (seed, sx, sy, gx, gy, flags, aw, ms)
