"""The product's density arithmetic (csrc/density_core.cuh, the roundings the CUDA kernels apply) compiled for the host in the
kernels' two-pass structure: bit-exact with SciPy's float32 uniform_filter, the reference's own arithmetic.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_density_oracle import scipy_density

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dens") / "density_core_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "density_core_host.cpp")],
                   check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("seed,shape,p_road,p_occ", [(1, (200, 200), 0.3, 0.1), (2, (64, 150), 0.6, 0.5), (3, (21, 21), 1.0, 0.0),
                                                     (4, (5, 90), 0.2, 1.0), (5, (130, 7), 0.05, 0.3)])
def test_product_density_core_is_bit_exact_with_scipy(core, seed, shape, p_road, p_occ):
    rng = np.random.default_rng(seed)
    road = (rng.random(shape) < p_road).astype(np.uint8)
    occ = ((rng.random(shape) < p_occ) & (road == 1)).astype(np.uint8)
    out = np.zeros(shape, np.float32)
    u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    core.host_density(shape[1], shape[0], u8(occ), u8(road), out.ctypes.data_as(C.POINTER(C.c_float)))
    want = scipy_density(occ.astype(np.int8), road.astype(np.int8)).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
