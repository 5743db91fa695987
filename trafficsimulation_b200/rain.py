"""The rain map on the device: ``RainManager.step`` (agents/rain.py:154-185) with the rasterisation of the
clouds (``RainAgent.step`` :61-72) as a CUDA kernel writing straight into the plane the tick kernel reads.

The clouds themselves -- spawn draws (:100-148), radius (:42), float motion ``x += dx`` (:59-60), exit test (:74-82) -- are a
handful of scalars per tick and stay with the caller; this class takes their positions and radii as they are after the
reference's own update and keeps ``rain_map`` in step:

    rain = GpuRain(width, height, device)                     # rain.rain_map: uint8 [H, W] on the device (GpuTraffic's tape plane)
    rain.step([(cloud.x, cloud.y, cloud.radius) for cloud in model.rains])
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class GpuRain:
    def __init__(self, width, height, device="cuda:0", rain_map=None):
        if not torch.cuda.is_available():
            raise RuntimeError("trafficsimulation_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.W, self.H = int(width), int(height)
        self.device = torch.device(device)
        self.cfg = _lib.Cfg(self.W, self.H, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, self.H, 0)
        self.rain_map = torch.zeros(self.H, self.W, dtype=torch.uint8, device=self.device) if rain_map is None else rain_map
        self._prev = None

    @staticmethod
    def discs(clouds):
        """(x, y, radius) per cloud -> int32 [n, 3] = (int(x), int(y), radius): the centre cell RainAgent.step uses (:65)."""
        return np.array([(int(x), int(y), int(r)) for x, y, r in clouds], np.int32).reshape(-1, 3)

    def _write(self, discs, value):
        if discs is None or len(discs) == 0:
            return
        d = torch.from_numpy(np.ascontiguousarray(discs, np.int32)).to(self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.tsim_rain_discs(C.byref(self.cfg), C.c_void_p(d.data_ptr()), len(discs), value, C.c_void_p(self.rain_map.data_ptr()), stream))

    def step(self, clouds):
        """RainManager.step: clear what rained last tick (:156-158), set this tick's covered cells (:172-180)."""
        cur = self.discs(clouds)
        self._write(self._prev, 0)
        self._write(cur, 1)
        self._prev = cur
        return self.rain_map
