"""CPU stand-ins for the two device collaborators of ``trafficsimulation_b200.replan.PlannedTraffic`` -- the tick oracle and the
route-planner oracle behind the interfaces of ``GpuTraffic(route_capacity=...)`` and ``GpuAstar``.  TEST INFRASTRUCTURE: they let the
state machine of replan.py be checked against the reference's route events on a machine without a GPU; the GPU tests run the same
checks with the CUDA collaborators."""
import numpy as np

from oracle import oracle as O

ALGO = {"QUEUE_ACTUATED": 0, "FIXED_TIME": 1, "PRESSURE_CONTROL": 2, "NEIGHBOR_GREEN_WAVE": 3}


def without_routes(r):
    """The tapes of a tick fixture / harness run minus every route event."""
    t = {k: r[k] for k in ("spawn_tick", "origin", "target", "speed", "malfunction", "rank", "rain_map")}
    t.update(ev_tick=np.zeros(0, np.int32), ev_vehicle=np.zeros(0, np.int32), ev_off=np.zeros(1, np.int64), ev_cells=np.zeros(0, np.int32))
    return t


class OracleTrafficBackend:
    def __init__(self, W, H, tables, tapes, n_ticks, algo="QUEUE_ACTUATED", rain_enabled=False, route_capacity=1 << 22):
        nv = len(tapes["spawn_tick"])
        ne = max(nv, 1)
        # OracleTicks derives ev_first from ev_tick: hand it buffers of the right capacity, then take the index over
        tp = dict(tapes, ev_tick=np.zeros(ne, np.int32), ev_vehicle=np.zeros(ne, np.int32), ev_off=np.zeros(ne + 1, np.int64),
                  ev_cells=np.zeros(route_capacity, np.int32))
        self.sim = O.OracleTicks(W, H, tables, tp, n_ticks, algo=ALGO[algo], rain_enabled=rain_enabled)
        a = self.sim.a
        first = np.zeros(n_ticks + 2, np.int32)
        a["ev_first"] = first
        self.sim.sim.ev_first = first.ctypes.data
        self.nv, self.n_ticks, self.used, self.cap, self.t = nv, n_ticks, 0, route_capacity, 0

    def route_room(self):
        return self.cap - self.used

    def push_route_events(self, vehicles, paths, compact=False):
        a = self.sim.a
        if compact:
            self.used = 0
        n = len(vehicles)
        off = np.zeros(n + 1, np.int64)
        off[1:] = np.cumsum([len(p) for p in paths])
        total = int(off[-1])
        assert self.used + total <= self.cap, "route buffer full"
        a["ev_vehicle"][:n] = np.asarray(vehicles, np.int32)
        a["ev_off"][: n + 1] = off + self.used
        if total:
            a["ev_cells"][self.used: self.used + total] = np.concatenate([np.asarray(p, np.int32) for p in paths])
        self.used += total
        a["ev_first"][self.t] = 0
        a["ev_first"][self.t + 1:] = n

    def step(self, n=1):
        assert n == 1
        self.sim.run(1)
        self.t += 1

    def plan_snapshot(self):
        a = self.sim.a
        return dict(occupancy=a["occ"].copy(), stop_map=a["stop"].copy(), alive=a["alive"] == 1, pos=a["pos"].copy(), path_len=a["path_len"].copy(),
                    stuck_ticks=a["stuck_ticks"].copy(), stranded=a["stranded"].copy(), malfunction_flag=a["malfunction"] != 0,
                    collision_flag=a["collision"] != 0, base_speed=a["base_speed"].copy(), cur_speed=a["cur_speed"].copy(),
                    is_stuck=a["is_stuck"].copy(), direction=a["direction"].copy())

    def state_host(self):
        return self.sim.state()


_KEEP = object()


class OraclePlannerBackend:
    speculate = False   # a CPU search is not free: the soft search only runs when the strict one found nothing, as in the reference

    def __init__(self, W, H, is_road_map, road_type_map, allowed_dirs_map):
        self.W, self.H = W, H
        self.rank = None
        self.static = (np.asarray(is_road_map), np.asarray(road_type_map), np.asarray(allowed_dirs_map))
        self.occ = np.zeros((H, W), np.uint8)
        self.stop = np.zeros((H, W), np.uint8)
        self.density = np.zeros((H, W), np.float64)
        self._astar = None

    def update(self, occupancy_map=None, stop_map=None, spawn_rank_map=_KEEP):
        if occupancy_map is not None:
            self.occ = np.array(occupancy_map, np.uint8).reshape(self.H, self.W)
        if stop_map is not None:
            self.stop = np.array(stop_map, np.uint8).reshape(self.H, self.W)
        if spawn_rank_map is not _KEEP:
            self.rank = None if spawn_rank_map is None else np.array(spawn_rank_map, np.uint8).reshape(self.H, self.W)
        self._astar = None

    def update_density(self):
        self.density = O.density_map(self.occ, self.static[0]).astype(np.float64)
        self._astar = None

    def plan_cells(self, queries):
        if self._astar is None:
            self._astar = {}
        out = []
        for sx, sy, gx, gy, fl, aw, ms, lim in np.asarray(queries).reshape(-1, 8).tolist():
            key = lim if self.rank is not None else 0
            if key not in self._astar:   # the k-th spawn of a tick plans on the map without the spawns after it (tsim_astar_query.spawn_rank_limit)
                occ = self.occ if self.rank is None else np.where(self.rank > lim, 0, self.occ).astype(np.uint8)
                self._astar[key] = O.OracleAstar(occ, self.stop, *self.static, self.density)
            p = self._astar[key].query(sx, sy, gx, gy, bool(fl & 1), aw, bool(fl & 2), bool(fl & 4), ms)
            out.append(np.array([y * self.W + x for x, y in p], np.int32))
        return out
