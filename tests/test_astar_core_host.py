"""The product's route-planner core (csrc/astar_core.cuh, the code every CUDA thread runs) compiled for the host and
checked against the reference's golden vectors -- CPU only, so the kernel's logic is pinned before it ever sees a GPU."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from golden_util import load_astar

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "astar_*.npz")))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("astar") / "astar_core_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "astar_core_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.host_astar.restype = C.c_int
    return lib


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[6:-4])
def test_product_core_reproduces_golden(core, path):
    r = load_astar(path)
    W, H = r["W"], r["H"]
    keep = [np.ascontiguousarray(r[k], np.uint8) for k in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")]
    dens = np.ascontiguousarray(r["density"], np.float64)
    out = np.zeros(W * H, np.int32)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    for q, want in zip(r["queries"], r["paths"]):
        sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in q)
        n = core.host_astar(W, H, *[u8(a) for a in keep], dens.ctypes.data_as(C.POINTER(C.c_double)), sx, sy, gx, gy,
                            ra | (so << 1) | (ig << 2), 10, ms, out.ctypes.data_as(C.POINTER(C.c_int32)), len(out), 0)
        assert n >= 0 and out[:n].tolist() == list(want), tuple(q)


def test_product_core_reports_a_short_output_buffer(core):
    r = load_astar(FIXTURES[0])
    W, H = r["W"], r["H"]
    keep = [np.ascontiguousarray(r[k], np.uint8) for k in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")]
    i = int(np.argmax([len(p) for p in r["paths"]]))
    sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in r["queries"][i])
    out = np.zeros(4, np.int32)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    n = core.host_astar(W, H, *[u8(a) for a in keep], None, sx, sy, gx, gy, ra | (so << 1) | (ig << 2), 10, ms,
                        out.ctypes.data_as(C.POINTER(C.c_int32)), 4, 0)
    assert n < 0    # -(cells needed); with density = NULL the path may differ in length, never fit in 4 cells


def test_product_core_reports_open_list_overflow(core):
    """The open list is bounded (half the grid in tsim_astar_batch); outgrowing it is an error code, never a silent wrong path."""
    r = load_astar(FIXTURES[0])
    W, H = r["W"], r["H"]
    keep = [np.ascontiguousarray(r[k], np.uint8) for k in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")]
    i = int(np.argmax([len(p) for p in r["paths"]]))
    sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in r["queries"][i])
    out = np.zeros(W * H, np.int32)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    args = [W, H, *[u8(a) for a in keep], r["density"].ctypes.data_as(C.POINTER(C.c_double)), sx, sy, gx, gy, ra | (so << 1) | (ig << 2), 10, ms,
            out.ctypes.data_as(C.POINTER(C.c_int32)), len(out)]
    assert core.host_astar(*args, 8) == -0x40000000
    assert core.host_astar(*args, 0) == len(r["paths"][i])


def test_product_core_spawn_rank_limit(core):
    """tsim_astar_query.spawn_rank_limit: a query of the k-th spawn of a tick must plan exactly as if the vehicles spawned after it
    were not on the grid yet -- the ranked search on the full map == the plain search on the map with those cells cleared."""
    r = load_astar(FIXTURES[0])
    W, H = r["W"], r["H"]
    rng = np.random.default_rng(3)
    keep = {k: np.ascontiguousarray(r[k], np.uint8) for k in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")}
    road = np.flatnonzero(keep["is_road_map"].reshape(-1) == 1)
    born = rng.choice(road, 140, replace=False)                   # more spawns than the 7-bit rank field holds: the last ones share 127
    occ = keep["occupancy"].copy().reshape(-1)
    occ[rng.choice(road, 300, replace=False)] = 1                 # standing traffic
    occ[born] = 1
    rank = np.zeros(W * H, np.uint8)
    rank[born] = np.minimum(np.arange(1, len(born) + 1), 127)
    dens = np.ascontiguousarray(r["density"], np.float64)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    out_a, out_b = np.zeros(W * H, np.int32), np.zeros(W * H, np.int32)
    core.host_astar_ranked.restype = C.c_int
    differ = 0
    for k in (1, 2, 17, 60, 126):
        plain = occ.copy()
        plain[born[k:]] = 0                                       # what spawn k sees: spawns 1..k on the grid
        for flags in (0, 2, 4, 6):
            for _ in range(6):
                g = int(rng.choice(road))
                sx, sy, gx, gy = int(born[k - 1] % W), int(born[k - 1] // W), g % W, g // W
                rest = [u8(keep[n]) for n in ("stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")]
                tail = [dens.ctypes.data_as(C.POINTER(C.c_double)), sx, sy, gx, gy, flags, 10, 0x7FFFFFFF]
                na = core.host_astar_ranked(W, H, u8(occ), *rest, *tail, out_a.ctypes.data_as(C.POINTER(C.c_int32)), len(out_a), 0, u8(rank), k)
                nb = core.host_astar(W, H, u8(plain), *rest, *tail, out_b.ctypes.data_as(C.POINTER(C.c_int32)), len(out_b), 0)
                assert na == nb and np.array_equal(out_a[:na], out_b[:nb]), (k, flags, g)
                nc = core.host_astar(W, H, u8(occ), *rest, *tail, out_b.ctypes.data_as(C.POINTER(C.c_int32)), len(out_b), 0)
                differ += nc != na or not np.array_equal(out_a[:na], out_b[:nc])
    assert differ > 0                                             # the later spawns do change some of these routes
