"""Drive the UNMODIFIED reference tick loop (`CityModel.step`) under recorded tapes.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference; its outputs travel as tests/golden/ticks_*.npz.

Tape conventions:
  * activation order inside `schedule.step()` (city_model.py:1858) is injected through `model.random`:
    light groups first (canonical order: by the minimum (y,x) cell of their intersection cluster),
    then city blocks, then vehicles by ascending `rank[tick, vehicle]`, the traffic generator last;
  * `run_parallel_decide` runs with ONE worker (F4) so phase A visits vehicles in spawn order;
  * `random.randint` inside `step_decide` (vehicle_base.py:112) returns `speed[tick, vehicle]`; the FIRST
    `random.random` of a `step_decide` (the malfunction draw, vehicle_base.py:609) returns 0.0 where bit 0 of
    `malfunction[tick, vehicle]` is set, else 0.5; the SECOND one (the sideswipe draw, :600 -- it only happens when a
    moving vehicle heading the opposite way stands next to the vehicle) returns 0.0 where bit 1 is set, else 0.5;
  * the traffic generator's step is replaced by a tape-driven spawner: attempt `k` creates vehicle `k` at
    `origin[k]` heading for `target[k]` at tick `spawn_tick[k]` unless the origin cell is occupied, in
    which case the attempt is DROPPED (the reference's own generator would stack vehicles; the generator
    is outside the hot path and both arms use this spawner);
  * every path the reference plans (`_compute_path`, vehicle_base.py:143) is recorded as a route event
    (tick, vehicle, cells) -- at spawn, and whenever `step_decide` re-planned.
"""
from __future__ import annotations

import random as _random

import numpy as np

from . import harness as H


class TapeRandom:
    """Stands in for `model.random`; only `shuffle` is used by the scheduler stub."""

    def __init__(self, model, sort_key):
        self.model = model
        self.sort_key = sort_key

    def shuffle(self, keys):
        agents = self.model.schedule._agents
        keys.sort(key=lambda k: self.sort_key(agents[k]))

    def __getattr__(self, name):   # anything else falls through to a private generator
        return getattr(_random.Random(0), name)


def light_group_tables(model):
    """Canonical light-group tables of the reference model (for parity of the construction).

    Returns (groups, order) where groups[i] is a dict of sorted int arrays (cell indices y*W+x):
      cluster, lights, ns_lights, ew_lights, ns_in, ew_in, ns_out, ew_out (multisets)
    in canonical order, and `order` lists the reference group objects in that order.
    """
    W = model.width
    idx = lambda pos: pos[1] * W + pos[0]
    groups = []
    for g in model.intersection_light_groups:
        cluster = sorted(idx(c.position) for c in g.intersection_cells)
        pairs = g.get_opposite_traffic_lights()
        groups.append((cluster[0], g, dict(
            cluster=np.array(cluster, np.int32),
            lights=np.array(sorted(idx(t.position) for t in g.traffic_lights), np.int32),
            ns_lights=np.array(sorted(idx(t.position) for t in pairs["N-S"]), np.int32),
            ew_lights=np.array(sorted(idx(t.position) for t in pairs["W-E"]), np.int32),
            ns_in=np.array(sorted(int(y) * W + int(x) for x, y in np.asarray(g.ns_in_coords).reshape(-1, 2)), np.int32),
            ew_in=np.array(sorted(int(y) * W + int(x) for x, y in np.asarray(g.ew_in_coords).reshape(-1, 2)), np.int32),
            ns_out=np.array(sorted(int(y) * W + int(x) for x, y in np.asarray(g.ns_out_coords).reshape(-1, 2)), np.int32),
            ew_out=np.array(sorted(int(y) * W + int(x) for x, y in np.asarray(g.ew_out_coords).reshape(-1, 2)), np.int32),
        )))
    creation = {id(g): i for i, g in enumerate(model.intersection_light_groups)}
    groups.sort(key=lambda t: t[0])
    canon = {id(t[1]): i for i, t in enumerate(groups)}
    for i, (_, g, d) in enumerate(groups):
        # neighbour links as populate_links left them (intersection_light_group.py:175-242; the second pass has just run: the
        # get_opposite_traffic_lights() call above), canonical indices, columns N, S, E, W; and the group's creation rank
        d["nbr"] = np.array([canon[id(g.neighbor_groups[k])] if k in (g.neighbor_groups or {}) else -1 for k in ("N", "S", "E", "W")], np.int32)
        d["creation_rank"] = np.array([creation[id(g)]], np.int32)
    return [t[2] for t in groups], [t[1] for t in groups]


def run_ticks(seed, n_ticks, spawns_per_tick=6, malfunction_p=0.0, rain_rect=None, layout_kwargs=None, tape_seed=None, sideswipe_p=0.0,
              algo=None):
    """Build the reference city with `random.seed(seed)` and run `n_ticks` ticks under tapes.
    `algo`: Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM for the run (config.py:341; None = as shipped, QUEUE_ACTUATED).

    Returns a dict with the layout (as harness.run_layout), the tapes and the per-tick states.
    """
    ref = H.load_reference()
    cm, Defaults, VehicleAgent = ref.cm, ref.Defaults, ref.VehicleAgent
    import Simulation.agents.vehicles.vehicle_base as vb
    from Simulation.agents.city_structure_entities.intersection_light_group import IntersectionLightGroup
    from Simulation.agents.city_structure_entities.city_block import CityBlock

    saved = (Defaults.TOTAL_SERVICE_VEHICLES_FOOD, Defaults.TOTAL_SERVICE_VEHICLES_WASTE, Defaults.RAIN_ENABLED)
    saved_algo = Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM
    if algo is not None:
        Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM = algo
    Defaults.TOTAL_SERVICE_VEHICLES_FOOD = 0
    Defaults.TOTAL_SERVICE_VEHICLES_WASTE = 0
    lay = H.run_layout(seed, enable_traffic=True, enable_rain=False, keep_model=True, **(layout_kwargs or {}))
    model = lay["model"]
    assert model is not None, lay["crashed"]
    Defaults.ENABLE_TRAFFIC = True
    Defaults.RAIN_ENABLED = rain_rect is not None   # only the `rain_map` byte plane enters the hot path (vehicle_base.py:103-105)
    W, H_ = model.width, model.height
    if rain_rect is not None:
        x0, y0, x1, y1 = rain_rect
        model.rain_map[y0:y1, x0:x1] = 1
    rng = np.random.default_rng(seed if tape_seed is None else tape_seed)
    n_attempts = n_ticks * spawns_per_tick
    starts = model.get_start_blocks()
    plan = []          # per attempt: (tick, origin cell agent, target cell agent)
    for k in range(n_attempts):
        o = starts[rng.integers(len(starts))]
        # get_valid_exits() is empty for highway entrances (its adjacency test returns a distance,
        # city_model.py:2093-2099,2138-2146), so destinations are drawn from get_exit_blocks()
        exits = [e for e in model.get_exit_blocks() if e.position != o.position]
        t = exits[rng.integers(len(exits))]
        plan.append((k // spawns_per_tick, o, t))
    speed = rng.integers(1, 6, size=(n_ticks, n_attempts), dtype=np.uint8)
    malf = (rng.random((n_ticks, n_attempts)) < malfunction_p)
    rank = np.stack([rng.permutation(n_attempts) for _ in range(n_ticks)]).astype(np.int32)
    swipe = (rng.random((n_ticks, n_attempts)) < sideswipe_p) if sideswipe_p > 0 else np.zeros((n_ticks, n_attempts), bool)   # drawn last: older fixtures keep their tapes

    groups, group_order = light_group_tables(model)
    gindex = {id(g): i for i, g in enumerate(group_order)}
    state = {"tick": 0, "veh": None, "planned": False, "swipes": 0}
    spawned = np.zeros(n_attempts, np.uint8)
    route_events = []   # (tick, vidx, [cells])
    vehicles = {}

    def sort_key(agent):
        if isinstance(agent, IntersectionLightGroup):
            return (0, gindex[id(agent)])
        if isinstance(agent, CityBlock):
            return (1, agent.unique_id)
        if isinstance(agent, VehicleAgent):
            return (2, int(rank[state["tick"], agent._tsim_idx]))
        return (3, 0)

    model.random = TapeRandom(model, sort_key)
    dta = model.dynamic_traffic_generator

    def spawner():
        dta.elapsed += dta.dt
        t = state["tick"]
        for k in range(t * spawns_per_tick, (t + 1) * spawns_per_tick):
            _, o, tg = plan[k]
            ox, oy = o.position
            if model.occupancy_map[oy, ox] == 1:
                continue
            v = VehicleAgent(f"V{k}", model, o, tg, population_type="internal")
            v._tsim_idx = k
            vehicles[k] = v
            spawned[k] = 1
            route_events.append((t, k, [py * W + px for px, py in v.path]))

    dta.step = spawner

    orig_decide = VehicleAgent.step_decide
    orig_compute = VehicleAgent._compute_path
    real_randint, real_random = _random.randint, _random.random

    def tape_randint(a, b):
        return int(speed[state["tick"], state["veh"]])

    def tape_random():
        calls = state["rcalls"]
        state["rcalls"] = calls + 1
        if calls == 0 and malf[state["tick"], state["veh"]]:
            return 0.0
        if calls == 1 and swipe[state["tick"], state["veh"]]:
            state["swipes"] += 1
            return 0.0
        return 0.5

    def decide(self):
        state["veh"] = self._tsim_idx
        state["planned"] = False
        state["rcalls"] = 0
        _random.randint, _random.random = tape_randint, tape_random
        try:
            orig_decide(self)
        finally:
            _random.randint, _random.random = real_randint, real_random
        if state["planned"]:
            route_events.append((state["tick"], self._tsim_idx, [py * W + px for px, py in self.path]))

    def compute(self, use_cache=True):
        state["planned"] = True
        return orig_compute(self, use_cache)

    VehicleAgent.step_decide = decide
    VehicleAgent._compute_path = compute
    old_cpu = cm.multiprocessing.cpu_count
    cm.multiprocessing.cpu_count = lambda: 1

    DIRS = {None: -1, "N": 0, "E": 1, "S": 2, "W": 3}
    pos = np.full((n_ticks, n_attempts), -1, np.int32)
    base = np.zeros((n_ticks, n_attempts), np.int8)
    stuck = np.zeros((n_ticks, n_attempts), np.int16)
    flags = np.zeros((n_ticks, n_attempts), np.uint8)   # bit0 is_stuck, bit1 malfunction, bits 2-4 direction+1, bit5 collision
    occ_cells, stop_cells, stuckmap_cells, group_phase, group_ext = [], [], [], [], []
    try:
        with H._in_tmpdir():
            for t in range(n_ticks):
                state["tick"] = t
                model.step()
                for k, v in vehicles.items():
                    if v.pos is None:
                        continue
                    pos[t, k] = v.pos[1] * W + v.pos[0]
                    base[t, k] = v.base_speed
                    stuck[t, k] = v.stuck_ticks
                    flags[t, k] = (1 if v.is_stuck else 0) | (2 if v.is_in_malfunction else 0) | ((DIRS[v.direction] + 1) << 2) | \
                                  (32 if v.is_in_collision else 0)
                occ_cells.append(np.flatnonzero(model.occupancy_map.reshape(-1)).astype(np.int32))
                stop_cells.append(np.flatnonzero(model.stop_map.reshape(-1)).astype(np.int32))
                stuckmap_cells.append(np.flatnonzero(model.stuck_map.reshape(-1)).astype(np.int32))
                group_phase.append(np.array([[-1 if g.current_phase is None else g.current_phase,
                                              -1 if g.pending_phase is None else g.pending_phase,
                                              g.queue_timer, g.gap_timer, g.last_arrival] for g in group_order], np.int32))
                clamp = lambda v: max(-2 ** 31, min(2 ** 31 - 1, int(v)))   # NEIGHBOR_PRESSURE_CONTROL's pressures outgrow any fixed width
                group_ext.append(np.array([[g.fixed_time_timer, g._ft_phase, clamp(g.ns_pressure), clamp(g.ew_pressure)] for g in group_order], np.int32))
    finally:
        VehicleAgent.step_decide = orig_decide
        VehicleAgent._compute_path = orig_compute
        cm.multiprocessing.cpu_count = old_cpu
        Defaults.TOTAL_SERVICE_VEHICLES_FOOD, Defaults.TOTAL_SERVICE_VEHICLES_WASTE, Defaults.RAIN_ENABLED = saved
        Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM = saved_algo
        Defaults.ENABLE_TRAFFIC = False

    def ragged(lst):
        off = np.zeros(len(lst) + 1, np.int64)
        off[1:] = np.cumsum([len(a) for a in lst])
        return off, (np.concatenate(lst) if lst and off[-1] else np.zeros(0, np.int32)).astype(np.int32)

    ev_off, ev_cells = ragged([np.array(c, np.int32) for _, _, c in route_events])
    out = dict(
        layout=lay, W=W, H=H_, n_ticks=n_ticks, n_attempts=n_attempts, spawns_per_tick=spawns_per_tick,
        spawn_tick=np.array([p[0] for p in plan], np.int32),
        origin=np.array([p[1].position[1] * W + p[1].position[0] for p in plan], np.int32),
        target=np.array([p[2].position[1] * W + p[2].position[0] for p in plan], np.int32),
        spawned=spawned, speed=speed, malfunction=(malf.astype(np.uint8) | (swipe.astype(np.uint8) << 1)), rank=rank, sideswipes_fired=state["swipes"],
        rain_map=model.rain_map.astype(np.uint8).copy(),
        ev_tick=np.array([e[0] for e in route_events], np.int32), ev_vehicle=np.array([e[1] for e in route_events], np.int32),
        ev_off=ev_off, ev_cells=ev_cells,
        pos=pos, base_speed=base, stuck_ticks=stuck, vflags=flags,
        groups=groups, group_state=np.stack(group_phase) if group_phase and len(group_order) else np.zeros((n_ticks, 0, 5), np.int32),
        group_ext=np.stack(group_ext) if group_ext and len(group_order) else np.zeros((n_ticks, 0, 4), np.int32),   # fixed-time timer / phase, pressures
    )
    for name, lst in (("occ", occ_cells), ("stop", stop_cells), ("stuckmap", stuckmap_cells)):
        out[name + "_off"], out[name + "_cells"] = ragged(lst)
    return out
