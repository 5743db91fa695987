"""Batched route planning on the device: the reference's ``astar_numba`` for many queries at once.

Mirrors ``Simulation/utilities/pathfinding/astar_numba.py::astar_numba(width, height, start_x, start_y, goal_x, goal_y,
occupancy_map, stop_map, is_road_map, road_type_map, allowed_dirs_map, respect_awareness, awareness_range, density_map,
soft_obstacles, ignore_flow, maximum_steps)`` (:240-281): same argument meaning, same result -- the reference's path cell
for cell, not merely one of equal cost -- as ``[(x, y), ...]`` from the first step to the goal, ``[]`` when
there is no route.  The maps live on the device; ``plan`` takes any number of queries and runs one CUDA thread per query
(``tsim_astar_batch``).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

RESPECT_AWARENESS, SOFT_OBSTACLES, IGNORE_FLOW = 1, 2, 4
UNBOUNDED = 0x7FFFFFFF


class GpuAstar:
    def __init__(self, width, height, occupancy_map, stop_map, is_road_map, road_type_map, allowed_dirs_map, density_map=None,
                 device="cuda:0", scratch_bytes=4 << 30):
        if not torch.cuda.is_available():
            raise RuntimeError("trafficsimulation_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.W, self.H = int(width), int(height)
        self.cfg = _lib.Cfg(self.W, self.H, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, self.H, 0)
        self.maps = {}
        self.update(occupancy_map=occupancy_map, stop_map=stop_map, is_road_map=is_road_map, road_type_map=road_type_map,
                    allowed_dirs_map=allowed_dirs_map, density_map=density_map)
        per_query = C.c_size_t(0)
        _lib.check(self.lib.tsim_astar_scratch_bytes(C.byref(self.cfg), 1, C.byref(per_query)))
        self.chunk = max(1, int(scratch_bytes) // per_query.value)
        self._scratch = None
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    def update(self, **maps):
        """Replace maps (host arrays or device tensors, [H][W]); the tick's occupancy / stop planes can be passed as they are.
        spawn_rank_map (u8, optional; None removes it): which spawn of the running tick stands on a cell, see ``plan_cells``."""
        for k, v in maps.items():
            if v is None:
                self.maps[k] = None
                continue
            if k not in ("occupancy_map", "stop_map", "is_road_map", "road_type_map", "allowed_dirs_map", "density_map", "spawn_rank_map"):
                raise ValueError(f"unknown map {k!r}")
            dt = torch.float64 if k == "density_map" else torch.uint8
            t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            t = t.to(device=self.device, dtype=dt).contiguous().view(-1)
            if t.numel() != self.W * self.H:
                raise ValueError(f"{k}: expected {self.H} x {self.W} cells")
            self.maps[k] = t

    def update_density(self):
        """``CityModel._update_density_map`` (city_model.py:1764-1778) on the device, from the current occupancy and road maps;
        the result (float64 plane, the reference's float32 values widened) becomes the planner's ``density_map``."""
        n = self.W * self.H
        if self.maps.get("density_map") is None:
            self.maps["density_map"] = torch.empty(n, dtype=torch.float64, device=self.device)
        if self._scratch is None or self._scratch.numel() < 2 * n:
            self._scratch = torch.empty(max(2 * n, 1 << 20), dtype=torch.uint8, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.tsim_density_map(C.byref(self.cfg), C.c_void_p(self.maps["occupancy_map"].data_ptr()),
                                             C.c_void_p(self.maps["is_road_map"].data_ptr()), None, C.c_void_p(self.maps["density_map"].data_ptr()),
                                             C.c_void_p(self._scratch.data_ptr()), C.c_size_t(self._scratch.numel()), stream))
        return self.maps["density_map"].view(self.H, self.W)

    def _maps_struct(self):
        m = self.maps
        ptr = lambda k: m[k].data_ptr() if m.get(k) is not None else None
        return _lib.AstarMaps(ptr("occupancy_map"), ptr("stop_map"), ptr("is_road_map"), ptr("road_type_map"), ptr("allowed_dirs_map"),
                              ptr("density_map"), ptr("spawn_rank_map"))

    def plan_cells(self, queries, max_path=None):
        """queries: int array [n, 7] = (sx, sy, gx, gy, flags, awareness_range, maximum_steps), or [n, 8] with a spawn-rank limit
        as the last column: the query of the k-th vehicle the spawner placed this tick carries k and sees an occupied cell whose
        ``spawn_rank_map`` entry is above k as free -- the vehicles spawned after it are not on the grid yet when it plans
        (vehicle_base.py:72-76), so all spawns of a tick plan in one batch.  Returns a list of int32 arrays of cell indices
        ``y * W + x`` (first step first, goal last; empty = no route)."""
        q = np.ascontiguousarray(queries, np.int32)
        q = q.reshape(-1, 8 if q.ndim == 2 and q.shape[1] == 8 else 7)
        if q.shape[1] == 7:
            q = np.concatenate([q, np.zeros((len(q), 1), np.int32)], 1)
        out = []
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        maps = self._maps_struct()
        cap = int(max_path or min(self.W * self.H, 4 * (self.W + self.H)))
        for a in range(0, len(q), self.chunk):
            part = q[a:a + self.chunk]
            n = len(part)
            dq = torch.from_numpy(np.ascontiguousarray(part)).to(self.device)
            need = C.c_size_t(0)
            _lib.check(self.lib.tsim_astar_scratch_bytes(C.byref(self.cfg), n, C.byref(need)))
            if self._scratch is None or self._scratch.numel() < need.value:
                self._scratch = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            while True:
                lens = torch.empty(n, dtype=torch.int32, device=self.device)
                cells = torch.empty(n * cap, dtype=torch.int32, device=self.device)
                self.flag.zero_()
                _lib.check(self.lib.tsim_astar_batch(C.byref(self.cfg), C.byref(maps), C.c_void_p(dq.data_ptr()), n, C.c_void_p(lens.data_ptr()),
                                                     C.c_void_p(cells.data_ptr()), cap, C.c_void_p(self.flag.data_ptr()),
                                                     C.c_void_p(self._scratch.data_ptr()), C.c_size_t(self._scratch.numel()), stream))
                err = int(self.flag.item())
                h_len = lens.cpu().numpy()
                if err == 50:                      # some path is longer than the buffer: once more with room for the longest
                    cap = int(-h_len.min())
                    continue
                if err:
                    raise _lib.TsimError(6 if err == 51 else 1, f"tsim_astar_batch error flag {err}")
                break
            h_cells = cells.view(n, cap).cpu().numpy()
            out += [h_cells[i, :h_len[i]].copy() for i in range(n)]
        return out

    def plan(self, queries, max_path=None):
        """Same, as the reference's ``[(x, y), ...]`` lists."""
        return [[(int(c % self.W), int(c // self.W)) for c in p] for p in self.plan_cells(queries, max_path)]

    def astar(self, start_x, start_y, goal_x, goal_y, respect_awareness=False, awareness_range=10, soft_obstacles=False, ignore_flow=False,
              maximum_steps=UNBOUNDED):
        """One query with the reference's argument names (drop-in for a single ``astar_numba`` call; batch with ``plan`` for speed)."""
        flags = (RESPECT_AWARENESS if respect_awareness else 0) | (SOFT_OBSTACLES if soft_obstacles else 0) | (IGNORE_FLOW if ignore_flow else 0)
        return self.plan([[start_x, start_y, goal_x, goal_y, flags, awareness_range, maximum_steps]])[0]
