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
