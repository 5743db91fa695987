"""Route planning joined to the tick: the vehicles plan their own routes, no route tape.

Mirrors the planner side of ``VehicleAgent`` (``Simulation/agents/vehicles/vehicle_base.py``):

* ``_compute_path`` :143-167 (cooldown, the path cache shared by all vehicles, used at spawn time only),
* ``_compute_path_internal`` :199-420 (phase 0: merge back onto the saved route while overtaking / detouring; phase 1: strict
  avoidance; phase 2: soft obstacles; phase 3: contraflow bypass of a stranded vehicle on the next cell; phase 4: contraflow detour
  when stuck for long),
* ``_recompute_path_on_stuck`` :506-517 and ``_recompute_path_on_obstacle`` :454-504 (the two re-plan triggers of ``step_decide``
  :616-663, with the cooldown and its stranded-blocker exception).

How it is joined.  With ``PATHFINDING_BATCHING`` (config.py:411, the shipped setting) the reference runs every vehicle's
``step_decide`` before anything moves (``run_parallel_decide``, city_model.py:1811-1829, then ``schedule.step()`` :1858): every
re-plan of a tick reads the tick-START occupancy / stop / density maps and the vehicle's own state.  The re-plans of one tick are
therefore independent of each other -- the one cross-vehicle read is ``blocker.is_stranded()``, which sees the blocker's state after
ITS ``step_decide`` when the blocker is earlier in ``active_vehicle_agents`` (spawn order) and its tick-start state otherwise; both
are functions of the tick-start state and the draw tapes (a sideswipe draw that fires strands two vehicles in the middle of the phase:
``PlannedTraffic._early_exits`` replays the head of every ``step_decide`` in list order to know who is stranded at whose turn).  So a tick is: snapshot -> every vehicle's trigger logic as a
coroutine that yields its A* queries -> the queries of all vehicles answered round by round as ONE ``tsim_astar_batch`` launch per
round -> the new routes handed to the tick kernel as this tick's route events -> ``tsim_tick_run`` for one tick.  A vehicle that
spawns plans on the maps as they are at that moment of the tick (after the moves, with the earlier spawns of the same tick on the
grid, city_model.py:1897-1908 / vehicle_base.py:72-76) against the density map of the tick's start; its route reaches the device as
an event of the next tick, before phase A.

This module holds the state machine and the batching; the searches run on the device (``pathfinding.GpuAstar``), the tick on the
device (``traffic.GpuTraffic(route_capacity=...)``).  There is no CPU fallback: both collaborators are handed in by the caller, and
``PlannedTraffic.on_gpu`` builds the CUDA ones.
"""
from __future__ import annotations

import numpy as np

from .pathfinding import IGNORE_FLOW, SOFT_OBSTACLES, UNBOUNDED

# Defaults of Simulation/config.py the planner side reads (line numbers there)
AWARENESS_RANGE = 10                 # VEHICLE_AWARENESS_RANGE :279
OVERTAKE_STEPS = 6                   # VEHICLE_MAX_CONTRAFLOW_OVERTAKE_STEPS :302
OVERTAKE_DURATION = 30               # VEHICLE_CONTRAFLOW_OVERTAKE_DURATION :304
STUCK_RECOMPUTE = 30                 # VEHICLE_STUCK_RECOMPUTE_THRESHOLD :306
STUCK_RECOMPUTE_INTERSECTION = 1     # VEHICLE_STUCK_RECOMPUTE_THRESHOLD_INTERSECTION :307
STUCK_CONTRAFLOW = 60                # VEHICLE_STUCK_CONTRAFLOW_THRESHOLD :310
STUCK_CONTRAFLOW_INTERSECTION = 10   # VEHICLE_STUCK_CONTRAFLOW_THRESHOLD_INTERSECTION :311
DETOUR_STEPS = 20                    # VEHICLE_MAX_CONTRAFLOW_STUCK_DETOUR_STEPS :312
DETOUR_DURATION = 10                 # VEHICLE_CONTRAFLOW_STUCK_DETOUR_DURATION :313
COOLDOWN = 5                         # PATHFINDING_COOLDOWN :409
MALFUNCTION_TICKS = 400              # VEHICLE_MALFUNCTION_DURATION :321
COLLISION_TICKS = 600                # VEHICLE_SIDESWIPE_COLLISION_DURATION :326
RAIN_SPEED_REDUCTION = 2             # RAIN_SPEED_REDUCTION :263
RANK_CHUNK = 126                     # spawns of one tick that plan in one batch (tsim_astar_maps.spawn_rank holds 7 bits)


class PlanState:
    """The planner-side fields of one ``VehicleAgent`` (vehicle_base.py:43-56); cells are ``y * W + x``."""
    __slots__ = ("v", "target", "path", "cooldown", "is_overtaking", "overtake_path", "pre_overtake_path", "overtaking_duration",
                 "is_in_stuck_detour", "stuck_detour_path", "pre_stuck_detour_path", "stuck_detour_duration", "pos", "stuck_ticks",
                 "planned")

    def __init__(self, v, target):
        self.v, self.target = v, target
        self.path = []
        self.cooldown = 0
        self.is_overtaking, self.overtake_path, self.pre_overtake_path, self.overtaking_duration = False, [], [], -1
        self.is_in_stuck_detour, self.stuck_detour_path, self.pre_stuck_detour_path, self.stuck_detour_duration = False, [], [], -1
        self.pos, self.stuck_ticks, self.planned = -1, 0, False


class _View:
    """What a vehicle's trigger logic may read of the city at the moment it runs."""
    __slots__ = ("occ", "stop", "inter", "veh_at", "stranded_of")

    def __init__(self, occ, stop, inter, veh_at, stranded_of):
        self.occ, self.stop, self.inter, self.veh_at, self.stranded_of = occ, stop, inter, veh_at, stranded_of


def _first_free(cells, occ):
    for i, c in enumerate(cells):
        if occ[c] == 0:
            return i
    return None


def scan_ahead(path, occ, stop):
    """``_scan_ahead_for_obstacles`` :422-452: indices of the first stop cell / occupied cell among the next cells of the route."""
    idx_stop = idx_vehicle = None
    for i in range(min(AWARENESS_RANGE, len(path))):
        c = path[i]
        if idx_stop is None and stop[c] == 1:
            idx_stop = i
        if idx_vehicle is None and occ[c] == 1:
            idx_vehicle = i
        if idx_stop == 0 or idx_vehicle == 0:
            break
    return idx_stop, idx_vehicle


def compute_path_internal(s, view):
    """``_compute_path_internal`` :199-420 as a coroutine: yields ``(start, goal, flags, maximum_steps)`` (or a list of such
    alternatives), is sent the path."""
    occ, stop = view.occ, view.stop
    pos, goal = s.pos, s.target
    # phase 0: back onto the saved route while overtaking / detouring (:218-274)
    for overtaking in (True, False):
        active, saved = (s.is_overtaking, s.pre_overtake_path) if overtaking else (s.is_in_stuck_detour, s.pre_stuck_detour_path)
        if active and saved:
            merge = _first_free(saved, occ)
            if merge is not None:
                b = saved[merge]
                bypass = yield (pos, b, IGNORE_FLOW, OVERTAKE_STEPS)
                if bypass and bypass[-1] == b:
                    if overtaking:
                        s.overtake_path = bypass
                    else:
                        s.stuck_detour_path = bypass
                    return bypass + saved[merge + 1:]
    # phase 1: strict avoidance, phase 2: soft obstacles (:276-303)
    # (a LIST of queries = "the first of these that finds a route": the strict search fails whenever a vehicle or a red light
    # stands anywhere on every way to the goal, which is nearly always, so both searches travel in the same batch)
    path = yield [(pos, goal, 0, UNBOUNDED), (pos, goal, SOFT_OBSTACLES, UNBOUNDED)]
    # phase 3: contraflow bypass of a stranded / parked vehicle on the next cell (:305-364)
    if path:
        idx_stop = idx_vehicle = None
        for i in range(min(AWARENESS_RANGE, len(path))):
            c = path[i]
            if idx_stop is None and stop[c] == 1:
                idx_stop = i
            if idx_vehicle is None and occ[c] == 1:
                idx_vehicle = i
            if idx_stop is not None and idx_vehicle is not None:
                break
        if idx_vehicle == 0:
            blocker = view.veh_at.get(path[0])
            if blocker is not None and view.stranded_of(blocker, s.v):
                k = _first_free(path, occ)
                if k is not None:
                    b = path[k]
                    bypass = yield (pos, b, IGNORE_FLOW, OVERTAKE_STEPS)
                    if bypass and bypass[-1] == b and len(bypass) > 1:
                        idx_bp = path.index(b)
                        s.pre_overtake_path, s.overtake_path = path, bypass
                        s.is_overtaking, s.overtaking_duration = True, 0
                        return bypass + path[idx_bp + 1:]
    # phase 4: contraflow detour when stuck for long (:366-418)
    if path:
        threshold = STUCK_CONTRAFLOW_INTERSECTION if view.inter[pos] else STUCK_CONTRAFLOW
        if s.stuck_ticks >= threshold:
            k = _first_free(path, occ)
            if k is not None:
                b = path[k]
                bypass = yield (pos, b, SOFT_OBSTACLES | IGNORE_FLOW, DETOUR_STEPS)
                if bypass and bypass[-1] == b and len(bypass) > 1:
                    merge = path.index(b)
                    s.pre_stuck_detour_path, s.stuck_detour_path = list(path), bypass
                    s.is_in_stuck_detour, s.stuck_detour_duration = True, 0
                    return bypass + path[merge + 1:]
            return path
    return path


def compute_path(s, view, cache=None):
    """``_compute_path`` :143-167.  ``cache``: the model-wide ``_path_cache`` dict (consulted when given: at spawn time)."""
    s.cooldown = COOLDOWN
    s.planned = True
    key = (s.pos, s.target)
    if cache is not None and key in cache:
        return list(cache[key])
    path = yield from compute_path_internal(s, view)
    if cache is not None and path and not s.is_overtaking and not s.is_in_stuck_detour:
        cache[key] = list(path)
    return path


def decide_replans(s, view):
    """The planner side of ``step_decide`` :645-649 for a vehicle that got past its early exits."""
    # _recompute_path_on_stuck :506-517
    thresh = STUCK_RECOMPUTE_INTERSECTION if view.inter[s.pos] else STUCK_RECOMPUTE
    if s.stuck_ticks >= thresh:
        s.path = yield from compute_path(s, view)
    # _recompute_path_on_obstacle :454-504
    if s.is_overtaking and (not s.overtake_path or s.pos not in s.overtake_path):
        s.overtake_path, s.is_overtaking = None, False
    if s.is_in_stuck_detour and (not s.stuck_detour_path or s.pos not in s.stuck_detour_path):
        s.stuck_detour_path, s.is_in_stuck_detour = None, False
    idx_stop, idx_vehicle = scan_ahead(s.path, view.occ, view.stop)
    if s.is_overtaking:
        s.overtaking_duration += 1
        if s.overtaking_duration <= OVERTAKE_DURATION:
            return
    if s.is_in_stuck_detour:
        s.stuck_detour_duration += 1
        if s.stuck_detour_duration <= DETOUR_DURATION:
            return
    if s.cooldown > 0:
        hard_block = False
        if idx_vehicle == 0:
            blocker = view.veh_at.get(s.path[0])
            hard_block = blocker is not None and view.stranded_of(blocker, s.v)
        if not hard_block:
            s.cooldown -= 1
            return
    if idx_stop is not None or idx_vehicle is not None:
        path = yield from compute_path(s, view)
        if path:
            s.path = path


def run_coroutines(jobs, answer, limits=None, speculate=True):
    """Drive coroutines that yield A* queries: every round, the pending queries of all of them go to ``answer(list of queries)``
    (one batch launch) and each coroutine is resumed with its path.  A coroutine may yield a LIST of alternatives ("the first one
    that finds a route"): with ``speculate`` they all travel in the same round, otherwise one per round as long as they fail.
    ``jobs``: list of generators; ``limits[i]``: spawn-rank limit of every query of job i (default 0); returns the return values."""
    results = [None] * len(jobs)
    pending = []

    def advance(i, g, send):
        try:
            q = g.send(send) if send is not None else next(g)
            pending.append([i, g, q if isinstance(q, list) else [q], 0])
        except StopIteration as e:
            results[i] = e.value

    for i, g in enumerate(jobs):
        advance(i, g, None)
    while pending:
        todo, pending = pending, []
        asked, span = [], []
        for i, g, alts, k in todo:
            take = alts[k:] if speculate else alts[k:k + 1]
            span.append(len(take))
            asked += [q + ((limits[i] if limits else 0),) for q in take]
        paths = answer(asked)
        at = 0
        for (i, g, alts, k), n in zip(todo, span):
            got = next((p for p in paths[at:at + n] if p), [])
            at += n
            if not got and k + n < len(alts):
                pending.append([i, g, alts, k + n])     # (without speculation) the next alternative, next round
            else:
                advance(i, g, got)
    return results


class _WithoutLater:
    """An occupancy plane as the k-th spawn of a tick sees it: the cells of the vehicles spawned after it read as free."""
    __slots__ = ("occ", "later")

    def __init__(self, occ, later):
        self.occ, self.later = occ, later

    def __getitem__(self, c):
        return 0 if c in self.later else self.occ[c]


class PlannedTraffic:
    """The tick with the vehicles' own route planning: ``CityModel.step()`` (city_model.py:1831-1860) without a route tape.

    traffic: a ``GpuTraffic`` built with ``route_capacity=`` (``plan_snapshot()``, ``push_route_events()``, ``step()``);
    planner: a ``GpuAstar`` on the same city (``update``, ``update_density``, ``plan_cells``); ``on_gpu`` builds both.
    tapes: the draw tapes ``malfunction[T, V]`` (bit 0 malfunction, bit 1 sideswipe) and ``speed[T, V]``, the spawn targets and, with
    ``rain_enabled``, ``rain_map`` (the same tapes the tick kernel consumes -- minus the route events).
    """

    def __init__(self, traffic, planner, width, height, intersection_map, tapes, record_events=True, rain_enabled=False):
        self.traffic, self.planner = traffic, planner
        self.W, self.H = int(width), int(height)
        self.inter = np.ascontiguousarray(intersection_map, np.uint8).reshape(-1)
        self.target = np.asarray(tapes["target"], np.int64)
        self.malfunction = np.asarray(tapes["malfunction"])
        self.speed = np.asarray(tapes["speed"])
        rain = tapes.get("rain_map") if rain_enabled else None
        self.rain = None if rain is None else np.ascontiguousarray(rain, np.uint8).reshape(-1)
        self.veh = {}                 # live vehicle -> PlanState
        self.cache = {}               # CityModel._path_cache
        self.pending = {}             # vehicle -> route to hand to the device with the next tick's events
        self.tick = 0
        self.events = []              # (tick, vehicle, cells) in the reference's order: what a route tape of this run would hold
        self.record_events = bool(record_events)   # (a long run plans millions of routes: switch the log off and count instead)
        self.routes_planned = 0
        self.searches = 0             # A* queries answered
        self.batches = 0              # planner launches
        self.compactions = 0          # times the route buffer was started again
        self._snap = None
        self.speculate = bool(getattr(planner, "speculate", True))   # alternatives of a search in one batch (cheap on the device)

    @classmethod
    def on_gpu(cls, width, height, light_tables, tapes, n_ticks, maps, algo="QUEUE_ACTUATED", rain_enabled=False, device="cuda:0",
               route_cells=None, record_events=True):
        """Both collaborators on the device.  maps: the city's simple maps (``GpuCityLayout.maps_host()`` or device planes):
        is_road_map, road_type_map, intersection_map, allowed_dirs_map."""
        from .pathfinding import GpuAstar
        from .traffic import GpuTraffic
        nv = len(tapes["spawn_tick"])
        cells = int(route_cells or max(1 << 22, 64 * nv * 8))
        traffic = GpuTraffic(width, height, light_tables, tapes, n_ticks, algo=algo, rain_enabled=rain_enabled, device=device,
                             route_capacity=cells)
        zero = np.zeros((height, width), np.uint8)
        planner = GpuAstar(width, height, zero, zero, maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"], device=device)
        inter = maps["intersection_map"]
        inter = inter.cpu().numpy() if hasattr(inter, "cpu") else inter
        return cls(traffic, planner, width, height, inter, tapes, record_events=record_events, rain_enabled=rain_enabled)

    def _log(self, t, v, path):
        self.routes_planned += 1
        if self.record_events:
            self.events.append((t, v, list(path)))

    # ---- A* batches
    def _answer(self, queries):
        q = np.array([[a % self.W, a // self.W, b % self.W, b // self.W, fl, AWARENESS_RANGE, ms, lim] for a, b, fl, ms, lim in queries], np.int32)
        self.searches += len(q)
        self.batches += 1
        return [p.tolist() for p in self.planner.plan_cells(q)]

    def _early_exits(self, t, snap, live, veh_at):
        """The head of every vehicle's ``step_decide`` (:616-643) in ``active_vehicle_agents`` order, on the tick-start state: the
        stranded countdown, the malfunction draw, the sideswipe check -- which strands BOTH vehicles when its draw fires (:567-605) -- and
        the stop cell.  Returns the vehicles that get as far as the re-plan triggers, and ``stranded_of(u, me)``: ``u.is_stranded()`` as
        vehicle ``me`` reads it when its own turn comes (u's state after its own head if it is earlier in the list, plus whatever
        collision an earlier vehicle has inflicted on it by then)."""
        W, H = self.W, self.H
        lv = np.array(live)
        malf = dict(zip(live, (snap["malfunction_flag"][lv] != 0).tolist()))
        coll = dict(zip(live, (snap["collision_flag"][lv] != 0).tolist()))
        left = dict(zip(live, snap["stranded"][lv].tolist()))
        draws = dict(zip(live, self.malfunction[t, lv].tolist()))
        swipe = bool((self.malfunction[t, lv] & 2).any())
        if swipe:   # what the sideswipe check reads of the neighbours: speed granted by their last step_decide, is_stuck, direction
            cur = dict(zip(live, snap["cur_speed"][lv].tolist()))
            base = dict(zip(live, snap["base_speed"][lv].tolist()))
            stuck = dict(zip(live, (snap["is_stuck"][lv] != 0).tolist()))
            heading = dict(zip(live, snap["direction"][lv].tolist()))
            speed = dict(zip(live, self.speed[t, lv].tolist()))
        stop = snap["stop_map"]
        turn = {v: i for i, v in enumerate(live)}
        history = {}            # vehicle -> [(turn, stranded)] where its state changed during this phase A
        start = {v: malf[v] or coll[v] for v in live}
        who = []
        LEFT, RIGHT, DX, DY = (3, 0, 1, 2), (1, 2, 3, 0), (0, 1, 0, -1), (1, 0, -1, 0)
        for i, v in enumerate(live):
            if malf[v] or coll[v]:                              # _tick_stranded :552-565
                left[v] -= 1
                if left[v] <= 0:
                    malf[v] = coll[v] = False
                    left[v] = 0
                    history.setdefault(v, []).append((i, False))
                if malf[v] or coll[v]:
                    if swipe:
                        base[v] = cur[v] = 0
                    continue
            if draws[v] & 1:                                    # _check_malfunction :608-610
                malf[v], coll[v], left[v] = True, False, MALFUNCTION_TICKS
                history.setdefault(v, []).append((i, True))
                if swipe:
                    base[v] = cur[v] = 0
                continue
            pos = self.veh[v].pos
            if swipe and heading[v] >= 0:                       # _check_sideswipe_collision :567-605
                x, y = pos % W, pos // W
                for side in (LEFT[heading[v]], RIGHT[heading[v]]):
                    nx, ny = x + DX[side], y + DY[side]
                    if not (0 <= nx < W and 0 <= ny < H):
                        continue
                    u = veh_at.get(ny * W + nx)
                    if u is None or cur[u] <= 0 or stuck[u] or coll[u] or malf[u] or heading[u] != (heading[v] + 2) % 4:
                        continue
                    if draws[v] & 2:                            # the draw fires: _set_collision for both (:534-541)
                        for k in (v, u):
                            coll[k], malf[k], left[k], base[k], cur[k] = True, False, COLLISION_TICKS, 0, 0
                            history.setdefault(k, []).append((i, True))
                    break                                       # one draw per step_decide, whatever it gave
                if coll[v]:
                    continue
            if stop[pos] == 1:                                  # _is_at_stopped_cell :639-643
                if swipe:
                    base[v] = cur[v] = 0
                continue
            if swipe:                                           # _compute_speed :94-107 (the neighbours' checks read current_speed)
                if base[v] == 0:
                    base[v] = speed[v]
                sp = base[v]
                if self.rain is not None and self.rain[pos] == 1:
                    sp = max(1, sp - RAIN_SPEED_REDUCTION)
                cur[v] = sp
            who.append(v)

        def stranded_of(u, me):
            state = start[u]
            for when, value in history.get(u, ()):
                if when <= turn[me]:
                    state = value
            return state

        return who, stranded_of

    # ---- one tick
    def step(self, n=1, check=True):
        """``CityModel.step()`` n times.  (check: accepted for ``GpuTraffic.step``'s signature; every tick is checked -- the planner
        reads the device state between two ticks anyway.)"""
        for _ in range(int(n)):
            self._step_one()

    def state_host(self):
        """The city as ``GpuTraffic.state_host()`` gives it (positions, flags, maps, light groups)."""
        return self.traffic.state_host()

    def planner_fields(self, v):
        """The planner-side attributes of live vehicle v under the reference's names (vehicle_base.py:43-56), paths as ``(x, y)`` lists:
        what ``adaptor.GpuTickMirror.sync_to_model`` writes into the ``VehicleAgent``."""
        s, W = self.veh[v], self.W
        xy = lambda cells: None if cells is None else [(c % W, c // W) for c in cells]
        return dict(path=xy(s.path), path_retry_cooldown=s.cooldown, is_overtaking=s.is_overtaking, overtake_path=xy(s.overtake_path),
                    pre_overtake_path=xy(s.pre_overtake_path), overtaking_duration=s.overtaking_duration,
                    is_in_stuck_detour=s.is_in_stuck_detour, stuck_detour_path=xy(s.stuck_detour_path),
                    pre_stuck_detour_path=xy(s.pre_stuck_detour_path), stuck_detour_duration=s.stuck_detour_duration)

    def _step_one(self):
        t = self.tick
        if self._snap is None:
            if t != 0 or self.veh:
                raise RuntimeError("PlannedTraffic drives its GpuTraffic from tick 0 on")
            n = self.W * self.H   # nothing has run yet: an empty city (no vehicle reads its stop_map before the first tick has run)
            self._snap = dict(occupancy=np.zeros(n, np.uint8), stop_map=np.zeros(n, np.uint8))
        snap = self._snap
        occ, stop = snap["occupancy"], snap["stop_map"]
        # CityModel._update_density_map :1764-1778, from the tick-start occupancy; it also serves the spawns of this tick
        self.planner.update(occupancy_map=occ.reshape(self.H, self.W), stop_map=stop.reshape(self.H, self.W), spawn_rank_map=None)
        self.planner.update_density()
        live = sorted(self.veh)
        if live:
            veh_at = {self.veh[v].pos: v for v in live}
            who, stranded_of = self._early_exits(t, snap, live, veh_at)
            view = _View(occ, stop, self.inter, veh_at, stranded_of)
            jobs = []
            for v in who:
                self.veh[v].planned = False
                jobs.append(decide_replans(self.veh[v], view))
            run_coroutines(jobs, self._answer, speculate=self.speculate)
            for v in who:
                if self.veh[v].planned:
                    self._log(t, v, self.veh[v].path)
                    self.pending[v] = self.veh[v].path
        # ---- the tick itself, with this tick's routes
        vs = sorted(self.pending)
        if sum(len(self.pending[v]) for v in vs) > self.traffic.route_room():
            # the append-only route buffer is full: start it again with the remaining route of every live vehicle (a route event
            # that repeats the route a vehicle already follows changes nothing)
            vs = sorted(self.veh)
            self.traffic.push_route_events(vs, [self.veh[v].path for v in vs], compact=True)
            self.compactions += 1
        else:
            self.traffic.push_route_events(vs, [self.pending[v] for v in vs])
        self.pending = {}
        self.traffic.step(1)
        self.tick = t + 1
        snap = self._snap = self.traffic.plan_snapshot()
        alive = snap["alive"]
        # ---- routes advance by what the vehicles moved; the arrived are gone
        for v in live:
            s = self.veh[v]
            if not alive[v]:
                del self.veh[v]
                continue
            moved = len(s.path) - int(snap["path_len"][v])
            if moved < 0 or moved > 5 or (moved and s.path[moved - 1] != snap["pos"][v]):
                raise RuntimeError(f"tick {t}: vehicle {v} left its route (host route of {len(s.path)} cells, device {int(snap['path_len'][v])})")
            if moved:
                del s.path[:moved]
            s.pos, s.stuck_ticks = int(snap["pos"][v]), int(snap["stuck_ticks"][v])
        # ---- the spawns of this tick plan their first route (VehicleAgent.__init__ :72-76).  The reference plans them one after the
        # other, a later spawn of the same tick not being on the grid yet when an earlier one plans: every query carries its spawn's
        # rank and the planner reads the cells of higher ranks as free, so they all plan in the same batches
        born = [int(v) for v in np.flatnonzero(alive) if int(v) not in self.veh]
        occ2, stop2 = snap["occupancy"], snap["stop_map"]
        stranded_now = snap["malfunction_flag"] | snap["collision_flag"]
        veh_at = {int(snap["pos"][v]): v for v in list(self.veh) + born}
        for j0 in range(0, len(born), RANK_CHUNK):
            chunk = born[j0:j0 + RANK_CHUNK]
            rank = np.zeros(self.W * self.H, np.uint8)
            rank[[int(snap["pos"][u]) for u in born[j0 + RANK_CHUNK:]]] = RANK_CHUNK + 1
            rank[[int(snap["pos"][u]) for u in chunk]] = np.arange(1, len(chunk) + 1)
            self.planner.update(occupancy_map=occ2.reshape(self.H, self.W), stop_map=stop2.reshape(self.H, self.W),
                                spawn_rank_map=rank.reshape(self.H, self.W))
            jobs, states = [], []
            for i, v in enumerate(chunk):
                s = PlanState(v, int(self.target[v]))
                s.pos = int(snap["pos"][v])
                later = {int(snap["pos"][u]) for u in born[j0 + i + 1:]}
                o = _WithoutLater(occ2, later) if later else occ2
                at = veh_at if not later else {c: u for c, u in veh_at.items() if c not in later}
                view = _View(o, stop2, self.inter, at, lambda u, me: bool(stranded_now[u]))
                jobs.append(compute_path(s, view, self.cache))
                states.append(s)
            paths = run_coroutines(jobs, self._answer, limits=list(range(1, len(chunk) + 1)), speculate=self.speculate)
            for s, path in zip(states, paths):
                s.path = path or []
                self.veh[s.v] = s
                self._log(t, s.v, s.path)
                self.pending[s.v] = s.path
