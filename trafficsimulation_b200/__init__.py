"""trafficsimulation_b200 -- B200-native (sm_100a) data-parallel core of TrafficSimulation's CityModel.

Only the hot path lives here: the layout passes and the vehicle tick, as hand-written CUDA
behind the C ABI of include/tsim.h, plus the Python host side that mirrors the reference interface.
"""
from .encoding import ZONES, TYPE_CODE, encode_dirs, decode_dirs  # noqa: F401
from .bands import BandParams, make_city_bands, bands_to_array   # noqa: F401

__all__ = ["ZONES", "TYPE_CODE", "encode_dirs", "decode_dirs", "BandParams", "make_city_bands", "bands_to_array",
           "GpuCityLayout"]


def __getattr__(name):   # torch / CUDA are imported only when the GPU classes are used
    if name == "GpuCityLayout":
        from .layout import GpuCityLayout
        return GpuCityLayout
    raise AttributeError(name)
