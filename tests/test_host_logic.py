"""CPU-only checks of the host side: encoding, band generator, line tables, library symbols."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from trafficsimulation_b200 import _lib, encoding
from trafficsimulation_b200.bands import BandParams, bands_to_array, make_city_bands
from golden_util import layout_fixtures, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "tsim.h")).read()
    declared = set(re.findall(r"\b(tsim_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"libtsim.so does not export {name}"
    assert declared == set(_lib.SYMBOLS)
    assert lib.tsim_version() == 5 == _lib.ABI_VERSION
    assert f"#define TSIM_ABI_VERSION {_lib.ABI_VERSION}" in header


def test_dirs_roundtrip():
    for dirs in ([], ["N"], ["E", "S", "N"], ["E", "N", "S"], ["N", "S", "E", "W"], ["W", "S"]):
        assert encoding.decode_dirs(encoding.encode_dirs(dirs)) == dirs
    assert encoding.encode_dirs(["N", "S", "E", "W"]) & 0xF == 0xF


@pytest.mark.reference
def test_encoding_matches_reference_tables():
    from oracle.refharness import harness as h
    ref = h.load_reference()
    assert list(ref.Defaults.ZONES) == encoding.ZONES
    assert list(ref.Defaults.AVAILABLE_CITY_BLOCKS) == encoding.AVAILABLE_CITY_BLOCKS
    for dirs in (["E", "S", "N"], ["N", "S", "E", "W"]):
        assert h.encode_dirs(dirs) == encoding.encode_dirs(dirs)


@pytest.mark.parametrize("path", layout_fixtures(), ids=lambda p: os.path.basename(p)[7:-4])
def test_band_generator_reproduces_reference_bands(path):
    """random.seed(k) + our generator == the band lists the reference drew (fixture)."""
    g = load(path)
    kw = {k: v for k, v in {**g["meta"]["cfg"], **g["meta"]["kwargs"]}.items() if k in BandParams.__dataclass_fields__}
    random.seed(g["meta"]["seed"])
    hb, vb = make_city_bands(BandParams(**kw))
    assert np.array_equal(bands_to_array(hb), g["hbands"])
    assert np.array_equal(bands_to_array(vb), g["vbands"])


def test_line_table_first_band_wins_and_flags():
    lib = _lib.load()
    bands = np.array([[17, 18, 2, 1], [30, 33, 1, -1], [32, 32, 3, 0], [50, 51, 2, 3]], np.int32)
    out = np.zeros(60, np.uint32)
    _lib.check(lib.tsim_build_line_table(bands.ctypes.data_as(C.c_void_p), 4, 60, out.ctypes.data_as(C.c_void_p)))
    e = int(out[32])
    assert e & 1 and (e >> 1) & 3 == 1 and (e >> 6) & 7 == 2 and (e >> 9) & 7 == 4 and (e >> 3) & 7 == 7   # R1 band wins over R3
    assert out[29] == 0
    e = int(out[18])
    assert (e >> 12) & 1 and (e >> 13) & 3 == 1 and not (e >> 15) & 1
    e = int(out[50])
    assert (e >> 15) & 1 and (e >> 16) & 3 == 0
    bad = np.array([[5, 20, 2, 1]], np.int32)
    assert lib.tsim_build_line_table(bad.ctypes.data_as(C.c_void_p), 1, 60, out.ctypes.data_as(C.c_void_p)) == 1
    assert b"not representable" in lib.tsim_last_error()


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from trafficsimulation_b200.layout import GpuCityLayout
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GpuCityLayout()
