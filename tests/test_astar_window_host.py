"""Groundwork for r2: the route search restricted to a window (csrc/astar_core.cuh::astar_search_window, host build).
Whenever it answers at all it must give the reference's path (or the reference's "no route"); otherwise it must say that it
would have left the window.  Also measures how often a margin suffices -- the number r2 sizes its work arrays with."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from golden_util import load_astar

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "astar_*.npz")))
ERR_WINDOW = -0x40000001


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("astar_win") / "astar_core_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "astar_core_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.host_astar_window.restype = C.c_int
    return lib


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[6:-4])
@pytest.mark.parametrize("margin", [4, 16, 48])
def test_windowed_search_is_exact_or_says_so(core, path, margin):
    r = load_astar(path)
    W, H = r["W"], r["H"]
    keep = [np.ascontiguousarray(r[k], np.uint8) for k in ("occupancy", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map")]
    dens = np.ascontiguousarray(r["density"], np.float64)
    out = np.zeros(W * H, np.int32)
    cells = C.c_int(0)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    answered = with_route = area = 0
    for q, want in zip(r["queries"], r["paths"]):
        sx, sy, gx, gy, ra, so, ig, ms = (int(v) for v in q)
        n = core.host_astar_window(W, H, *[u8(a) for a in keep], dens.ctypes.data_as(C.POINTER(C.c_double)), sx, sy, gx, gy,
                                   ra | (so << 1) | (ig << 2), 10, ms, margin, out.ctypes.data_as(C.POINTER(C.c_int32)), len(out), C.byref(cells))
        if n == ERR_WINDOW:
            continue
        assert n >= 0 and out[:n].tolist() == list(want), (margin, tuple(q))
        answered += 1
        with_route += len(want) > 0
        area += cells.value
    frac = answered / len(r["queries"])
    print(f"margin {margin}: {answered}/{len(r['queries'])} answered inside the window ({with_route} with a route), mean window {area / max(answered, 1):.0f} of {W * H} cells")
    if margin >= 48:
        assert frac > 0.5
