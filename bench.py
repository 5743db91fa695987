#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native CityModel hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

Workload: a S x S synthetic city, ALL layout generation passes.  Default S = 16384 (BASELINE.json configs[2],
the size the north star's roofline target is quoted on); configs[1] (4096 x 4096) is timed beside it and
reported under "config1_4096".  With --gpus N the city has S x N*S cells in N row-band shards (weak scaling).
Passes -- frame, roads, sub-block carving, flood-fill zoning, dead ends, R2 upgrade, block
entrances, direction fixes, traffic lights with controlled-road linking, derived maps.  One "step" is
one full generation of the city from its band lists and tapes.

Prints ONE JSON line (see the task contract): `value` = cells/s with inputs resident in HBM,
`e2e` = the same through the public API with host buffers (tapes H2D, planes + maps D2H inside the
timed region), `roofline` for the dominant pass, `cpu_baseline` = the C oracle port on the host.
`--impl reference` times the CPU restatement (oracle port; the Python reference cannot travel to
the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per cell per pass.  SURVEY.md §8d counts dense reads + writes (frame 4, carve 6, zones 6, dead ends 2 per sweep,
# R2 4, entrances 6, direction fixes 5 + 5, lights 5, maps 8 = 52 with one sweep); where a pass only CHANGES a sparse set of cells
# the compulsory traffic is the read alone, and that smaller figure is the one the fractions below use (so that no fraction can
# exceed 1): R2 upgrade reads T (1 B), the direction fixes read T + D (3 B).
SURVEY_BYTES = {"frame_roads": 4, "carve": 6, "zones": 6, "dead_ends": 2, "upgrade_r2": 4, "entrances": 6, "fix_dirs": 10, "lights": 5, "maps": 8}
PASS_BYTES = dict(SURVEY_BYTES, upgrade_r2=1, fix_dirs=3)
CPU_SAMPLE = 2048   # the CPU arm runs a CPU_SAMPLE x CPU_SAMPLE city per step
PASS_TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r2_pass_traffic_16384.json")   # written by profiles/pass_traffic.py from an ncu launch list


def pass_traffic(size):
    """DRAM bytes and kernel shares per pass of one 16384 x 16384 city, from the committed ncu capture (None for other sizes)."""
    if size != 16384 or not os.path.exists(PASS_TRAFFIC_JSON):
        return None
    return json.load(open(PASS_TRAFFIC_JSON))["passes"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_inputs(size, seed, carve=True):
    from trafficsimulation_b200 import tapes
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    return hb, vb, cap, tapes.synth_zone_tape(seed, cap), np.zeros(cap, np.int32)


# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(size, seed):
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    hb, vb, cap, tz, te = synth_inputs(size, seed)
    cfg = O.make_cfg(width=size, height=size, fast_reach=1)
    o0 = O.OracleCity(cfg, hb, vb)
    o0.frame(); o0.roads()
    _CPU.update(O=O, cfg=cfg, hb=hb, vb=vb, tz=tz, te=te, tc=tapes.synth_carve_tape(seed, o0.nothing_blobs()))


def _cpu_one(_):
    c = _CPU
    oc = c["O"].OracleCity(c["cfg"], c["hb"], c["vb"])
    oc.run_all(c["tz"], c["tc"], c["te"], carve=True)
    oc.simple_maps()
    return 1


def cpu_arm(size, steps, warmup, seed=4096, procs=1):
    """The C oracle port, full pipeline on a size x size city; `procs` cities at once in worker processes (the
    reference itself is single-threaded Python, so more cores = independent replicas)."""
    import multiprocessing as mp
    times = []
    if procs == 1:
        _cpu_init(size, seed)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _cpu_one(0)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        with mp.get_context("fork").Pool(procs, initializer=_cpu_init, initargs=(size, seed)) as pool:
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                pool.map(_cpu_one, range(procs), chunksize=1)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
    return size * size * procs * len(times) / sum(times), sum(times) / len(times)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms by a reader thread (the recipe's clocks
    line, B200_PROFILING.md).  Rows are time-stamped; `summary(t0, t1)` uses the rows inside the window."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def _reader(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7 and parts[0].isdigit():
                self.rows.append((time.monotonic(), parts))

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._reader, daemon=True)
            self.thread.start()
            t_end = time.monotonic() + 5.0
            while not self.rows and time.monotonic() < t_end and self.proc.poll() is None:
                time.sleep(0.02)   # first sample in hand before the timed region starts
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self, t0=None, t1=None):
        rows = [r for t, r in self.rows if (t0 is None or t >= t0 - 0.06) and (t1 is None or t <= t1 + 0.06)]
        if not rows:
            rows = [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()) if any(r[2].replace(".", "", 1).isdigit() for r in rows) else None}


_T0 = time.time()


def log(msg):
    """progress on stderr (the JSON line on stdout stays alone)"""
    print(f"[bench +{time.time() - _T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


def vehicle_bench(dev, n_ticks=200, size=2048, n_vehicles=100000, cpu_ticks=20, route_len=400, e2e_ticks=50, parity_check=True):
    """Second headline metric: agent-updates/s of the vehicle CA tick (BASELINE.json configs[3], in the
    simultaneous-occupancy form SURVEY.md §8d defines: 100k vehicles live at once on a 2048^2 city)."""
    import torch
    from oracle import oracle as O
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    seed = 2048
    hb, vb, cap, tz, te = synth_inputs(size, seed)
    city = GpuCityLayout(width=size, height=size, device=dev)
    city.set_bands(hb, vb)
    city.generate(tz, None, te)
    log(f"  tick leg {size}: city generated")
    tabs = light_tables_from_layout(city)
    planes = city.planes_host()
    log(f"  tick leg {size}: light tables built")
    tp = tapes.synth_traffic(seed, size, size, planes["cell_type"], planes["dirs"], n_vehicles, n_ticks, route_len=route_len, spawn_ticks=1)
    nv = len(tp["origin"])
    log(f"  tick leg {size}: tapes for {nv} vehicles")
    sim = GpuTraffic(size, size, tabs, tp, n_ticks, device=dev)
    sim.step(5)           # warm-up ticks (also spawns everybody)
    torch.cuda.synchronize()
    log(f"  tick leg {size}: device state ready, 5 warm-up ticks done")
    c0 = sim.counters()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sim.step(n_ticks - 5, check=False)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c1 = sim.counters()
    updates = c1["vehicle_updates"] - c0["vehicle_updates"]
    ticks = c1["tick"] - c0["tick"]
    peak, peak_src = measured_peaks()
    ups = updates / (ms * 1e-3)
    log(f"  tick leg {size}: timed ticks done ({ms / max(ticks, 1):.4f} ms/tick)")
    # end to end through GpuTraffic with HOST buffers: every tick its tape rows (speed, malfunction, rank of that tick) go up from pinned
    # memory, one launch advances the tick, and the positions of all vehicles + the tick counters come back to the host
    sim2 = GpuTraffic(size, size, tabs, tp, n_ticks, device=dev)
    sim2.step(5)
    torch.cuda.synchronize()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_rows = {k: pin(tp[k][5:5 + e2e_ticks]) for k in ("speed", "malfunction", "rank")}
    d_rows = {k: sim2.tt[k].reshape(-1, nv) for k in ("speed", "malfunction", "rank")}
    h_pos = torch.empty(nv, dtype=torch.int32).pin_memory()
    e0 = sim2.counters()["vehicle_updates"]
    t0 = time.perf_counter()
    for k in range(e2e_ticks):
        for name in ("speed", "malfunction", "rank"):
            d_rows[name][5 + k].copy_(h_rows[name][k], non_blocking=True)
        sim2.step(1, check=False)
        sim2.export()
        h_pos.copy_(sim2.s["pos"][:nv], non_blocking=True)
        sim2.counters()   # host read of the tick counters: synchronises
    t1 = time.perf_counter()
    e2e = (sim2.counters()["vehicle_updates"] - e0) / (t1 - t0)
    e2e_h2d, e2e_d2h = nv * (1 + 1 + 4), nv * 4 + 16 * 4
    del sim2
    log(f"  tick leg {size}: end-to-end ticks done")
    # CPU port on the same tapes: timed on `cpu_ticks` ticks, then run on to the tick the device is at, where the WHOLE state --
    # every vehicle's position / speed / stuck counter / flags, the three maps, the light groups -- must be identical
    cpu, parity = None, None
    if cpu_ticks > 0 or parity_check:
        ora = O.OracleTicks(size, size, tabs, tp, n_ticks)
        ora.run(5)
        live0 = int(ora.a["alive"].sum())
        if cpu_ticks > 0:
            t0 = time.perf_counter()
            upd_cpu = 0
            for _ in range(cpu_ticks):
                upd_cpu += int(ora.a["alive"].sum())
                ora.run(1)
            cpu_s = time.perf_counter() - t0
            cpu = {"value": upd_cpu / cpu_s, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                   "sample": f"oracle/vehicle_oracle.c, same tapes, {cpu_ticks} ticks, {live0} live vehicles"}
        log(f"  tick leg {size}: oracle at tick {ora.sim.tick}")
        if parity_check:
            ora.run(n_ticks - ora.sim.tick)
            log(f"  tick leg {size}: oracle at tick {ora.sim.tick}, comparing")
            got, want = sim.state_host(), ora.state()
            parity = {"parity_checked": bool(all(np.array_equal(got[k], want[k]) for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"))),
                      "against": f"oracle/vehicle_oracle.c after {n_ticks} ticks: positions, speeds, stuck counters, flags of all {nv} vehicles, occupancy / stop / stuck maps, light-group state"}
    return {"metric": "agent-updates/sec (vehicle CA tick with traffic-light gating)", "value": ups, "unit": "agent-updates/s",
            "config": {"workload": f"{size}x{size} city, {nv} vehicles spawned at tick 0, {ticks} timed ticks in one persistent launch",
                       "groups": sim.n_groups, "lights": sim.n_lights},
            "ms_per_tick": ms / ticks, "vehicle_updates": updates,
            "fixed_point_iterations_per_tick": (c1["fixed_point_iterations"] - c0["fixed_point_iterations"]) / ticks,
            "e2e": {"value": e2e, "unit": "agent-updates/s", "h2d_bytes_per_tick": e2e_h2d, "d2h_bytes_per_tick": e2e_d2h,
                    "note": "per tick: that tick's speed / malfunction / rank tape rows up from pinned host memory, one launch, all vehicle positions + the tick counters down"},
            "roofline": {"bound": "hbm", "alg_bytes_per_update": 84, "alg_bytes_per_group_update": 104,
                         "achieved": round((updates * 84 + ticks * sim.n_groups * 104) / (ms * 1e-3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                         "frac": round((updates * 84 + ticks * sim.n_groups * 104) / (ms * 1e-3) / 1e9 / peak, 5), "peak_source": peak_src,
                         "note": "SURVEY.md 8(d) compulsory bytes: 84 per agent-update + 104 per light-group update.  What bounds the tick is not bytes: "
                                 "every access is a 32-byte-sector gather or an L2 atomic (~5 atomics and ~25 sectors per vehicle and tick), and each of "
                                 "its phases is a chain of 3-5 dependent round trips between grid-wide barriers (DESIGN.md §5)"},
            "cpu_baseline": cpu, "parity": parity}


def sharded_vehicle_bench(dev, world, rank, size=4096, per_shard=150000, n_ticks=45, warm=5, route_len=100, halo=128):
    """The tick over `world` row-band shards, one per rank (SURVEY.md 8e "Vehicle step"): ONE city of `size` columns x
    `size * world` rows with `per_shard * world` vehicles live at once (weak scaling).  Every rank builds the same city,
    light tables and tapes (seeded), simulates its own band + halo, and refreshes the halos over NCCL after every tick.
    Rank 0 also runs the same tapes on one GPU (`GpuTraffic`) and the merged sharded state must be identical."""
    import torch
    import torch.distributed as dist
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
    from trafficsimulation_b200.sharded_traffic import ShardedTraffic
    W, H, seed = size, size * world, 4096
    # every rank holds the whole city's planes, cluster labels and routes on the host while it builds the tables (~6 GB per
    # rank at 8 shards): leave the leg out rather than push a small host into swap
    try:
        import psutil
        free_gb = torch.tensor([psutil.virtual_memory().available / 2**30], dtype=torch.float64, device=dev)
        dist.all_reduce(free_gb, op=dist.ReduceOp.MIN)
        if float(free_gb.item()) < 8.0 * world:
            return {"skipped": f"host memory: {float(free_gb.item()):.0f} GiB available, {8 * world} GiB wanted for {world} ranks"}
    except ImportError:
        pass
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
    city = GpuCityLayout(width=W, height=H, device=dev)
    city.set_bands(hb, vb)
    city.generate(tapes.synth_zone_tape(seed, cap), None, np.zeros(cap, np.int32))
    tabs = light_tables_from_layout(city)
    planes = city.planes_host()
    del city
    torch.cuda.empty_cache()
    tp = tapes.synth_traffic(seed, W, H, planes["cell_type"], planes["dirs"], per_shard * world, n_ticks, route_len=route_len, spawn_ticks=1)
    del planes
    nv = len(tp["origin"])
    sim = ShardedTraffic(W, H, tabs, tp, n_ticks, world, halo=halo, devices=[dev], distributed=True)
    sim.step(warm)
    u0 = sim.counters()["vehicle_updates"]
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sim.step(n_ticks - warm, check=False)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    upd = torch.tensor([sim.counters()["vehicle_updates"] - u0], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(upd)
    sim.check()
    got = sim.state_host()          # collective: every rank contributes what it owns
    same = None
    if rank == 0:
        one = GpuTraffic(W, H, tabs, tp, n_ticks, device=dev)
        one.step(n_ticks)
        want = one.state_host()
        same = all(np.array_equal(got[k], want[k]) for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"))
    ms = float(t.item())
    ticks = n_ticks - warm
    return {"metric": "agent-updates/sec (vehicle CA tick, row-band shards)", "value": float(upd.item()) / (ms * 1e-3), "unit": "agent-updates/s",
            "n_gpus": world, "ms_per_tick": ms / ticks, "vehicle_updates": int(upd.item()), "scaling": "weak",
            "config": {"workload": f"{W}x{H} city, {nv} vehicles spawned at tick 0, {ticks} timed ticks, one launch + one halo refresh per tick",
                       "halo_rows": halo, "groups": int(tabs["n_groups"])},
            "matches_single_gpu": same}


def sharded_layout_parity(dev, world, rank, sh, W, H, seed, d_tz, d_te):
    """Is the city the N ranks just generated the right one?  Two independent recomputations of the SAME city from the same
    tapes, compared through position-weighted digests (tsim_rows_digest, planes + maps) of the 2N half-bands of rows:
      * different cuts: the N ranks generate the city again with every cut moved by half a band, so that every row next to a
        cut of the timed run lies deep inside a shard (and the other way round);
      * one device: where the whole city fits the int32 cell index of one window (N = 2 at 65536 columns), rank 0 generates it
        alone, uncut.
    Returns {"parity_checked": ..., ...}; all ranks take part (collectives)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from trafficsimulation_b200 import _lib
    from trafficsimulation_b200.sharded import ShardedCityLayout
    n_half, hh = 2 * world, H // (2 * world)

    def digests(layout):   # [n_half, 8] int64: cell_type, dirs, aux, block_id, 4 maps of the half-bands this process owns
        out = torch.zeros(n_half, 8, dtype=torch.int64, device=dev)
        for s, L in layout.shards.items():
            lo, hi, w0 = layout.plan.own_lo[s], layout.plan.own_hi[s], layout.plan.win_lo[s]
            for j in range(n_half):
                a, b = j * hh, (j + 1) * hh
                if a < lo or b > hi:
                    continue
                for k, what in enumerate((1, 2, 4, 8)):
                    _lib.check(L.lib.tsim_rows_digest(C.byref(L.cfg), C.byref(L._planes), a - w0, b - w0, what,
                                                      C.c_void_p(out.data_ptr() + 8 * (8 * j + k)), L._stream))
                for k, name in enumerate(("is_road_map", "road_type_map", "intersection_map", "allowed_dirs_map")):
                    as_plane = _lib.Planes(L.maps[name].data_ptr(), 0, 0, 0)
                    _lib.check(L.lib.tsim_rows_digest(C.byref(L.cfg), C.byref(as_plane), a - w0, b - w0, 1,
                                                      C.c_void_p(out.data_ptr() + 8 * (8 * j + 4 + k)), L._stream))
        return out

    mine = digests(sh)
    dist.all_reduce(mine)                      # every half-band has one owner: the sum is its digest
    n_blocks = int(sh._total.item())
    result = {"half_bands": n_half, "digest_fields": 8}
    # ---- the same city, cut elsewhere
    h = H // world
    cuts = [k * h + h // 2 for k in range(1, world)]
    shb = ShardedCityLayout(world, halo=192, lean=True, distributed=True, cuts=cuts, width=W, height=H, carve_subblock_roads=True, device=dev)
    shb.set_bands(sh.shards[rank].hbands, sh.shards[rank].vbands)
    tcb = shb.synth_carve_tapes(seed)[rank]
    shb.generate(d_tz, {rank: tcb}, d_te, check=True)
    other = digests(shb)
    dist.all_reduce(other)
    result["vs_shifted_cuts"] = bool(torch.equal(mine, other)) and int(shb._total.item()) == n_blocks
    result["shifted_cuts_rows"] = cuts
    del shb, tcb
    torch.cuda.empty_cache()
    # ---- the same city on one device
    result["vs_single_gpu"] = None
    if W * (H + 1) < 2 ** 31:
        flag = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            one = ShardedCityLayout(1, width=W, height=H, carve_subblock_roads=True, device=dev)
            one.set_bands(sh.shards[rank].hbands, sh.shards[rank].vbands)
            tc1 = one.synth_carve_tapes(seed)
            one.generate(d_tz, tc1, d_te, check=True)
            flag[0] = int(torch.equal(digests(one), mine) and one.n_blocks == n_blocks)
            del one, tc1
            torch.cuda.empty_cache()
        dist.broadcast(flag, 0)
        result["vs_single_gpu"] = bool(flag.item())
    result["parity_checked"] = result["vs_shifted_cuts"] and result["vs_single_gpu"] is not False
    return result



def guarded(fn, rank, limit_s, on_timeout):
    """Run a collective leg that must never take the headline line down with it: an exception becomes an {"error": ...}
    entry; if the leg hangs (a rank died inside a collective) every rank leaves after `limit_s`, rank 0 printing first."""
    def bail():
        if rank == 0:
            on_timeout()
        sys.stdout.flush()
        os._exit(0)
    timer = threading.Timer(limit_s, bail)
    timer.daemon = True
    timer.start()
    try:
        return fn()
    except Exception as e:   # noqa: BLE001 -- reported, not swallowed
        return {"error": f"{type(e).__name__}: {e}"[:400]}
    finally:
        timer.cancel()


LEGS = {
    "tick100k": lambda dev: vehicle_bench(dev),
    "tick1m": lambda dev: vehicle_bench(dev, n_ticks=60, size=8192, n_vehicles=1000000, cpu_ticks=0, route_len=100, e2e_ticks=20),
    "routes": lambda dev: route_planning_leg(dev),
    "planned": lambda dev: planned_trips_leg(dev),
}


def run_leg(name, limit_s):
    """One of LEGS in a child process (`bench.py --leg NAME` prints its JSON object on stdout, progress on stderr)."""
    import subprocess
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", name], stdout=subprocess.PIPE, timeout=limit_s)
    except subprocess.TimeoutExpired:
        return {"error": f"leg {name}: no result within {limit_s} s (killed)"}
    lines = p.stdout.decode(errors="replace").strip().splitlines()
    if p.returncode != 0 or not lines:
        return {"error": f"leg {name}: exit code {p.returncode}"}
    try:
        out = json.loads(lines[-1])
    except ValueError:
        return {"error": f"leg {name}: unreadable output"}
    out["leg_wall_s"] = round(time.time() - t0, 1)
    return out


def leg_main(name):
    import torch
    try:
        out = LEGS[name](torch.device("cuda", 0))
    except Exception as e:   # noqa: BLE001 -- reported in the line
        out = {"error": f"{type(e).__name__}: {e}"[:400]}
    print(json.dumps(out))


def route_planning_leg(dev, n_queries=16384):
    """reference-exact routes/s of the batched planner on the reference's default city (the maps of the committed
    fixture tests/golden/astar_default12345.npz), the C oracle on a sample of the same queries beside it (same code as
    profiles/astar_bench.py)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import load_astar
    from oracle import oracle as O
    from trafficsimulation_b200.pathfinding import GpuAstar
    r = load_astar(os.path.join(ROOT, "tests", "golden", "astar_default12345.npz"))
    W, H = r["W"], r["H"]
    rng = np.random.default_rng(1)
    road = np.flatnonzero(r["is_road_map"].reshape(-1) == 1)
    a, b = rng.choice(road, n_queries), rng.choice(road, n_queries)
    q = np.stack([a % W, a // W, b % W, b // W, np.zeros(n_queries, np.int64), np.full(n_queries, 10), np.full(n_queries, 0x7FFFFFFF)], 1)
    free = np.zeros((H, W), np.uint8)
    planner = GpuAstar(W, H, free, r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"], device=dev)
    planner.plan_cells(q[:256])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    paths = planner.plan_cells(q)
    wall = time.perf_counter() - t0
    ora = O.OracleAstar(free, r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
    ns = 200
    t0 = time.perf_counter()
    ref = [ora.query(*[int(v) for v in q[i, :4]]) for i in range(ns)]
    cpu = time.perf_counter() - t0
    same = all([y * W + x for x, y in ref[i]] == paths[i].tolist() for i in range(ns))
    return {"metric": "routes/s (batched A*, the reference's path cell for cell)", "value": n_queries / wall, "unit": "routes/s",
            "config": {"workload": f"{n_queries} plain routes between random road cells of the default {W}x{H} city, one CUDA thread per query",
                       "with_route": int(sum(len(p) > 0 for p in paths)), "mean_path_cells": float(np.mean([len(p) for p in paths]))},
            "note": "host wall time: queries up, paths back on the host",
            "cpu_baseline": {"value": ns / cpu, "unit": "routes/s", "cores": 1, "kind": "port", "sample": f"{ns} of the same queries, oracle/astar_oracle.c"},
            "sample_matches_oracle": bool(same)}


def planned_trips_leg(dev, trips_per_tick=20, n_ticks=100, seed=1):
    """BASELINE.json configs[3] in the form it is named -- trips injected every tick on the reference's default city, NO route tape:
    the vehicles plan and re-plan their own routes (trafficsimulation_b200/replan.py: the re-plan triggers and the four-phase planner
    of vehicle_base.py:143-517, the searches as tsim_astar_batch launches, the tick as tsim_tick_run) -- bounded to `n_ticks` ticks.
    The city is the one of the committed reference fixture tests/golden/ticks_default12345.npz, built on the device from its tapes.
    Parity: the same loop with the C oracles behind the device interfaces must leave the same routes and the same final state."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import load_ticks
    from planning_backends import OracleTrafficBackend, OraclePlannerBackend
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.layout import GpuCityLayout
    from trafficsimulation_b200.replan import PlannedTraffic
    from trafficsimulation_b200.traffic import light_tables_from_layout
    r = load_ticks(os.path.join(ROOT, "tests", "golden", "ticks_default12345.npz"))
    W, H = r["W"], r["H"]
    cfgd = dict(r["meta"]["cfg"])
    carve = cfgd.pop("carve_subblock_roads", False)
    city = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
    city.set_bands(r["hbands"], r["vbands"])
    city.generate(r["tape_zone"], r["tape_carve"], r["tape_entrance"])
    tabs = light_tables_from_layout(city)
    maps, planes = city.maps_host(), city.planes_host()
    tp = tapes.synth_planned_trips(seed, W, H, planes["cell_type"], trips_per_tick, n_ticks, malfunction_p=0.002)
    sim = PlannedTraffic.on_gpu(W, H, tabs, tp, n_ticks, maps, device=str(dev))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.step(n_ticks)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    got = sim.traffic.state_host()
    updates = sim.traffic.counters()["vehicle_updates"]
    log(f"planned trips: {n_ticks} ticks in {wall:.1f} s, {sim.searches} searches in {sim.batches} batches")
    # the same loop on the CPU oracles: checker and baseline
    ora = PlannedTraffic(OracleTrafficBackend(W, H, tabs, tp, n_ticks, route_capacity=1 << 24),
                         OraclePlannerBackend(W, H, maps["is_road_map"], maps["road_type_map"], maps["allowed_dirs_map"]), W, H, maps["intersection_map"], tp)
    t0 = time.perf_counter()
    ora.step(n_ticks)
    cpu = time.perf_counter() - t0
    want = ora.traffic.state_host()
    same = sim.events == ora.events and all(np.array_equal(got[k], want[k]) for k in ("pos", "base_speed", "stuck_ticks", "vflags", "occ", "stop", "stuckmap", "groups"))
    return {"metric": "agent-updates/s with the vehicles planning their own routes (no route tape)", "value": updates / wall, "unit": "agent-updates/s",
            "ms_per_tick": wall / n_ticks * 1e3, "routes_per_s": sim.searches / wall,
            "config": {"workload": f"default {W}x{H} city, {trips_per_tick} trips injected per tick for {n_ticks} ticks, every route planned on the device "
                                   "(re-plan triggers + four-phase planner of vehicle_base.py:143-517), malfunction chance 0.002",
                       "trips": int(len(tp["spawn_tick"])), "spawned": int(len({v for _, v, _ in sim.events})), "live_at_end": int((got["pos"] >= 0).sum()),
                       "searches": int(sim.searches), "planner_launch_rounds": int(sim.batches), "routes_planned": int(len(sim.events)),
                       "route_buffer_compactions": int(sim.compactions)},
            "note": "host-orchestrated: per tick one state snapshot to the host, the trigger logic in Python, one tsim_astar_batch launch per round "
                    "of pending searches, one tsim_tick_run launch; wall time",
            "cpu_baseline": {"value": updates / cpu, "unit": "agent-updates/s", "cores": 1, "kind": "port",
                             "sample": f"the same {n_ticks} ticks: tick oracle + A* oracle behind the same loop, {cpu:.1f} s"},
            "parity_checked": bool(same)}


def capture_or_none(city, fn):
    """CUDA graph of one step (GpuCityLayout.capture), or (None, None) with the reason on stderr: the numbers then are eager ones."""
    import torch
    try:
        return city.capture(fn)
    except Exception as e:   # noqa: BLE001 -- reported; the eager path is the same kernels
        print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
        torch.cuda.synchronize()
        return None, None


def small_city_leg(dev, size=4096, steps=10, warmup=3, seed=4096):
    """BASELINE.json configs[1] (4096 x 4096, all passes, 1 GPU) next to the headline size: device-resident cells/s."""
    import torch
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout
    hb, vb = tapes.synth_bands(seed, width=size, height=size)
    sh = ShardedCityLayout(1, width=size, height=size, carve_subblock_roads=True, device=dev)
    sh.set_bands(hb, vb)
    tz = torch.from_numpy(tapes.synth_zone_tape(seed, sh.global_cap)).to(dev)
    te = torch.zeros(sh.global_cap, dtype=torch.int32, device=dev)
    tc = sh.synth_carve_tapes(seed)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        sh.generate(tz, tc, te, check=False)

    def timed(step_fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            step_fn()
            b.record()
        torch.cuda.synchronize()
        sh.shards[0]._check_flag("small_city_leg")
        return sum(a.elapsed_time(b) for a, b in ev) / steps
    ms_eager = timed(lambda: sh.generate(tz, tc, te, check=False))
    graph, launches = capture_or_none(sh.shards[0], lambda: sh.generate(tz, tc, te, check=False))
    ms = timed(graph.replay) if graph is not None else ms_eager
    return {"workload": f"{size}x{size} synthetic city layout, all generation passes, 1 GPU (BASELINE.json configs[1])",
            "value": size * size / (ms * 1e-3), "unit": "cells/s", "ms_per_step": ms, "steps": steps,
            "cuda_graph": graph is not None, "launches_per_step": launches, "ms_per_step_without_graph": ms_eager,
            "frac_of_measured_peak": round(size * size * sum(PASS_BYTES.values()) / (ms * 1e-3) / 1e9 / measured_peaks()[0], 4)}


def ours(args):
    import torch
    import torch.distributed as dist
    from trafficsimulation_b200 import tapes
    from trafficsimulation_b200.sharded import ShardedCityLayout

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    # N = 1: the size x size city (BASELINE.json configs[2]).  N > 1: BASELINE.json configs[4], ONE city of 65536 columns x 8192 * N
    # rows (65536 x 65536 at N = 8) in N row-band shards, one per GPU: weak scaling with 8192 rows per GPU.  The shards run LEAN
    # (sharded.py): 192 halo rows computed redundantly, no halo exchange; the rows around every cut are digested after every pass
    # and compared at the end of the step (one all-gather), the labellings all-gather their root counts.
    size, seed = args.size, 4096
    W, H = (size, size) if world == 1 else (args.shard_width, args.shard_rows * world)
    halo = 192 if world > 1 else 0
    hb, vb = tapes.synth_bands(seed, width=W, height=H)
    sh = ShardedCityLayout(world, halo=halo, lean=world > 1, distributed=world > 1, width=W, height=H, carve_subblock_roads=True, device=dev)
    sh.set_bands(hb, vb)
    city = sh.shards[rank]
    cap = sh.global_cap
    tz, te = tapes.synth_zone_tape(seed, cap), np.zeros(cap, np.int32)
    d_tz, d_te = torch.from_numpy(tz).to(dev), torch.from_numpy(te).to(dev)
    d_tc = sh.synth_carve_tapes(seed)[rank]
    own_rows = sh.plan.own_hi[rank] - sh.plan.own_lo[rank]
    cells = W * own_rows                                             # cells this GPU owns
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        sh.generate(d_tz, {rank: d_tc}, d_te, check=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    log(f"inputs ready ({W}x{H}, {world} GPU)")
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    city._check_flag("warmup")
    log("warm-up done")
    # one city = one CUDA graph launch on a single device (every launch is static); shards interleave NCCL collectives and stay eager
    eager_step, graph, graph_launches = step, None, None
    if world == 1 and not args.no_graph:
        graph, graph_launches = capture_or_none(city, eager_step)
        if graph is not None:
            step = graph.replay

    # ---- device-resident timing: K steps, L2 flushed between steps, CUDA events on the launch stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = city.lib.tsim_launch_count()
    clocks = ClockSampler(local).__enter__()   # keeps sampling through the per-pass breakdown below (same workload)
    t_clk0 = time.monotonic()
    for a, b in ev:
        flush.fill_(1)
        if world > 1:
            dist.barrier()          # all ranks enter the step together; the step itself is timed on the device
        a.record()
        step()
        b.record()
    barrier()
    launches = city.lib.tsim_launch_count() - launches0   # counted inside libtsim at every kernel launch
    if graph is not None:
        launches = graph_launches * args.steps              # replayed: the launches the graph holds (counted while it was captured)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    city._check_flag("timed steps")
    t_dev = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms = float(t_dev.item())
    value = W * H * args.steps / (ms * 1e-3)

    # ---- per-pass breakdown (single GPU only: on shards the passes interleave with the exchanges)
    passes = {}
    if world == 1:
        names = ["frame_roads", "carve", "zones", "dead_ends", "upgrade_r2", "entrances", "fix_dirs", "lights", "maps"]
        calls = [city._build_roads_and_sidewalks, lambda: city._carve_subblock_roads(d_tc, check=False),
                 lambda: city._flood_fill_blocks_storing_data(d_tz, check=False), city._eliminate_dead_ends,
                 lambda: city._upgrade_r2_to_intersections(check=False), lambda: city._final_place_block_entrances(d_te, check=False),
                 lambda: (city._remove_invalid_intersection_directions(), city._add_entrance_directions()),
                 lambda: city._add_traffic_lights(check=False), city._build_simple_maps]
        acc = {n: 0.0 for n in names}
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            flush.fill_(1)
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            marks[0].record()
            for i, f in enumerate(calls):
                f()
                marks[i + 1].record()
            torch.cuda.synchronize()
            for i, n in enumerate(names):
                acc[n] += marks[i].elapsed_time(marks[i + 1])
        peak, peak_src = measured_peaks()
        sweeps = city.sweeps()
        traffic = pass_traffic(size)
        for n in names:
            t = acc[n] / reps
            bpc = PASS_BYTES[n] * (sweeps if n == "dead_ends" else 1)
            gbs = cells * bpc / (t * 1e-3) / 1e9
            passes[n] = {"ms": round(t, 4), "cells_per_s": cells / (t * 1e-3), "alg_bytes_per_cell": bpc, "survey_bytes_per_cell": SURVEY_BYTES[n],
                         "achieved_gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4), "frac_of_nominal_8tbs": round(gbs / 8000.0, 4)}
            if traffic and n in traffic:   # what the pass really moved (ncu, committed capture) over the time measured here
                tr = traffic[n]
                passes[n].update(dram_bytes=int(tr["dram_bytes"]), dram_gbs=round(tr["dram_bytes"] / (t * 1e-3) / 1e9, 1), launches=tr["launches"],
                                 top_kernel=tr["top_kernel"], top_kernel_share=tr["top_kernel_share"])
        # the lights pass stage by stage (tsim_lights_prepare / _eval / _links): its top kernel, lights_eval_kernel<0>, and the merge of
        # its march marks ARE the eval stage but for four launches that return at once, so that stage's time is their duration, measured live
        stage = {"prepare": 0.0, "eval": 0.0, "links": 0.0}
        for _ in range(reps):
            for f in calls[:7]:
                f()
            flush.fill_(1)
            ev4 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev4[0].record(); city._lights_prepare(); ev4[1].record(); city._lights_eval(); ev4[2].record(); city._lights_links(check=False); ev4[3].record()
            torch.cuda.synchronize()
            for i, k in enumerate(stage):
                stage[k] += ev4[i].elapsed_time(ev4[i + 1]) / reps
        n_cr = int(city.workspace[: 64 * 4].view(torch.int32)[12].item())
        passes["lights"]["stages_ms"] = {k: round(v, 4) for k, v in stage.items()}
        passes["lights"]["candidates"] = n_cr
        passes["lights"]["undecided"] = city.lights_undecided()
    shard_phases = None
    if world > 1:   # where the step time goes on shards: device time between the marks of one more (untimed) step
        barrier()
        sh.trace = []
        step()
        shard_phases = {k: round(v, 3) for k, v in sh.phase_times()}
        sh.trace = None
    t_clk1 = time.monotonic()
    clocks.__exit__(None, None, None)

    log("timed steps and per-pass timing done")
    # ---- end to end through the public API with host buffers: tapes + band tables H2D, own rows of planes + maps D2H
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_tz, h_te, h_tc = pin(tz), pin(te), d_tc.cpu().pin_memory()
    h_rows, h_cols = city.row_table.cpu().pin_memory(), city.col_table.cpu().pin_memory()
    lo = (sh.plan.own_lo[rank] - sh.plan.win_lo[rank]) * W
    outs = {k: getattr(city, k)[lo: lo + cells] for k in ("cell_type", "dirs", "aux", "block_id")}
    h_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outs.items()}
    h2d = sum(t.numel() * t.element_size() for t in (h_tz, h_tc, h_te, h_rows, h_cols))

    # The 12 B/cell that go back to the host are the whole cost of a step (3.2 GB over PCIe at 16384^2 against 8 ms of kernels), so
    # they are copied out of a device staging copy on a second stream while the next step computes; every step still uploads its
    # inputs, runs all passes, is checked, and has its planes + maps land in host memory inside the timed region.
    h_maps = {k: torch.empty(cells, dtype=v.dtype).pin_memory() for k, v in city.maps.items()}
    stage = {k: torch.empty_like(v) for k, v in outs.items()}
    stage_maps = {k: torch.empty(cells, dtype=v.dtype, device=dev) for k, v in city.maps.items()}
    copy_stream = torch.cuda.Stream(device=dev)
    staged, drained = torch.cuda.Event(), torch.cuda.Event()
    drained.record()

    def e2e_step():
        city.row_table.copy_(h_rows, non_blocking=True); city.col_table.copy_(h_cols, non_blocking=True)
        d_tz.copy_(h_tz, non_blocking=True); d_tc.copy_(h_tc, non_blocking=True); d_te.copy_(h_te, non_blocking=True)
        step()
        torch.cuda.current_stream().wait_event(drained)     # the previous step's results have left the staging copy
        for k, v in outs.items():
            stage[k].copy_(v, non_blocking=True)
        for k, v in city.maps.items():
            stage_maps[k].copy_(v[lo: lo + cells], non_blocking=True)
        staged.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged)
            for k in outs:
                h_out[k].copy_(stage[k], non_blocking=True)
            for k in stage_maps:
                h_maps[k].copy_(stage_maps[k], non_blocking=True)
            drained.record()
        city._check_flag("e2e")                             # host waits for this step's kernels (not for its copy-out)

    e2e_step()
    d2h = sum(t.numel() * t.element_size() for t in list(h_out.values()) + list(h_maps.values()))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()                       # barrier() synchronizes the device first: the last step's results are in host memory
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = W * H * args.steps / float(t_e2e.item())
    log("end-to-end steps done")
    n_blocks = int(sh._total.item())
    n_lights = torch.tensor([int(((city._link_tensors["light_cell"][: int(city.flags[3].item())] >= lo) &
                                  (city._link_tensors["light_cell"][: int(city.flags[3].item())] < lo + cells)).sum().item())],
                            dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_lights)

    if rank == 0:
        peak, peak_src = measured_peaks()
        total_alg_per_cell = sum(PASS_BYTES.values())
        pipeline_gbs = cells * total_alg_per_cell * args.steps / (ms * 1e-3) / 1e9   # per GPU
        if passes:
            top = max(passes, key=lambda n: passes[n]["ms"])
            tp = passes[top]
            if top == "lights":   # kernel-level: the eval stage is lights_eval_kernel<0>
                k_ms = tp["stages_ms"]["eval"]
                # what that kernel must move: T + D of every cell once (3 B/cell: the candidates and the lanes behind them are spread over
                # the whole city), its list entry and its record per candidate (4 + 8 B)
                k_bytes = cells * 3 + tp["candidates"] * 12
                kern = (pass_traffic(size) or {}).get("lights", {}).get("kernels", {})
                k_dram = sum(kern.get(k, {}).get("dram_bytes", 0) for k in ("lights_eval_kernel<0>", "aux_light_merge_kernel")) or None
                roof = {"bound": "hbm", "kernel": "tsim::lights_eval_kernel<0> + its epilogue tsim::aux_light_merge_kernel (the evaluation stage of the top pass, lights)",
                        "kernel_ms": round(k_ms, 4), "kernel_launches_per_step": 2,
                        "achieved": round(k_bytes / (k_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(k_bytes / (k_ms * 1e-3) / 1e9 / peak, 4),
                        "traffic": int(k_dram) if k_dram else None, "dram_gbs": round(k_dram / (k_ms * 1e-3) / 1e9, 1) if k_dram else None,
                        "algorithmic_bytes_per_launch": int(k_bytes), "peak_source": peak_src,
                        "note": "duration = CUDA events around tsim_lights_eval on the launch stream (that stage is these two kernels plus four launches that "
                                "return at once; the merge kernel is ~1/6 of it); traffic = dram__bytes_read + write of the two kernels in the committed ncu "
                                "capture (profiles/r2_pass_traffic_16384.json)",
                        "pass": {"name": top, "ms": tp["ms"], "frac": tp["frac_of_measured_peak"], "dram_gbs": tp.get("dram_gbs"), "dram_bytes": tp.get("dram_bytes")}}
            else:
                roof = {"bound": "hbm", "kernel": f"tsim::{tp.get('top_kernel', top)} (top kernel of the top pass, {top}; share {tp.get('top_kernel_share')})",
                        "achieved": tp["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": tp["frac_of_measured_peak"], "traffic": tp.get("dram_bytes"),
                        "dram_gbs": tp.get("dram_gbs"), "peak_source": peak_src, "algorithmic_bytes_per_launch": cells * tp["alg_bytes_per_cell"],
                        "note": "pass-level: all kernels of the tsim_layout_* call, CUDA events on the launch stream"}
        else:
            roof = {"bound": "hbm", "kernel": "whole pipeline (per GPU)", "achieved": round(pipeline_gbs, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(pipeline_gbs / peak, 4), "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": cells * total_alg_per_cell}
        cpu_val, cpu_s = (cpu_arm(CPU_SAMPLE, 2, 1) if world == 1 else (None, None))
        log("cpu baseline done")
        line = {
            "metric": "grid cells/sec, full layout generation (all passes)", "value": value, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{W}x{H} synthetic city layout, all generation passes (carve + lights + maps)",
                       "cells_per_gpu": cells,
                       "parallelism": (f"{world} row-band shards of {own_rows} rows (+{halo} halo rows computed redundantly, no halo exchange; the rows around "
                                       "every cut are digested after every pass and compared at the end of the step; NCCL all-gather of root counts and digests") if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB write)", "seed": seed, "cuda_graph": graph is not None,
                       "blocks": n_blocks, "lights": int(n_lights.item()), "dead_end_sweeps": city.sweeps(),
                       "shard_rounds": {"dead_ends": getattr(sh, "dead_end_rounds", 1), "reach": getattr(sh, "reach_rounds", 1)}},
            "clocks": clocks.summary(t_clk0, t_clk1),
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "host buffers in and out every step; the copy-out of step i (from a device staging copy, second stream) overlaps the kernels of step i+1"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "pipeline": {"algorithmic_bytes_per_cell": total_alg_per_cell, "survey_bytes_per_cell": sum(SURVEY_BYTES.values()), "achieved_gbs_per_gpu": round(pipeline_gbs, 1),
                         "frac_of_measured_peak": round(pipeline_gbs / peak, 4), "frac_of_nominal_8tbs": round(pipeline_gbs / 8000.0, 4)},
            "passes": passes,
            "shard_phases_ms": shard_phases,
        }
        if world == 1:
            line["config1_4096"] = small_city_leg(dev) if size != 4096 else None
            log("4096 leg done")
            # the legs beside the headline run in processes of their own with a time limit each: a leg that fails, hangs or
            # crawls on a slow host costs its own entry, never the line (the parent keeps its device memory meanwhile)
            for key, leg, limit_s in (() if args.no_legs else (("vehicle_step", "tick100k", 150), ("vehicle_step_1M", "tick1m", 240), ("route_planning", "routes", 90),
                                                                    ("planned_trips", "planned", 150))):
                line[key] = run_leg(leg, limit_s)
                log(f"leg {leg} done" + (f": {line[key]['error']}" if isinstance(line[key], dict) and "error" in line[key] else ""))
            line["cpu_baseline"] = {"value": cpu_val, "unit": "cells/s", "cores": 1, "kind": "port",
                                    "sample": f"C oracle (oracle/city_oracle.c), same pipeline on a {CPU_SAMPLE}x{CPU_SAMPLE} city, {cpu_s:.2f} s/step"}
    if world > 1 and not args.no_parity:   # is it the right city?  (after all timing; collective)
        def parity_timeout():
            line["parity"] = {"error": "timed out"}
            print(json.dumps(line))
        par = guarded(lambda: sharded_layout_parity(dev, world, rank, sh, W, H, seed, d_tz, d_te), rank, 600, parity_timeout)
        if rank == 0:
            line["parity"] = par
    if world > 1:   # the second hot path on the same shards; a failure or hang here never costs the layout line
        tick = None
        if not args.no_sharded_tick:
            def on_timeout():
                line["vehicle_step_sharded"] = {"error": "timed out"}
                print(json.dumps(line))
            tick = guarded(lambda: sharded_vehicle_bench(dev, world, rank), rank, 420, on_timeout)
        if rank == 0:
            line["vehicle_step_sharded"] = tick
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 32))
    val, sec = cpu_arm(CPU_SAMPLE, args.steps, args.warmup, procs=cores)
    line = {
        "impl": "reference", "metric": "grid cells/sec, full layout generation (all passes)", "value": val, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.size}x{args.size} synthetic city layout, all generation passes (carve + lights + maps)",
                   "sample": f"{cores} x {CPU_SAMPLE}x{CPU_SAMPLE} cities per step, one per host core"},
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": f"C restatement of the reference's passes (oracle/city_oracle.c), {cores} independent {CPU_SAMPLE}x{CPU_SAMPLE} cities per step in {cores} worker processes; "
                                   "the Python reference itself cannot travel to the GPU box (BASELINE.md has its numbers)"},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--no-sharded-tick", action="store_true", help="N > 1: skip the sharded vehicle-tick leg")
    ap.add_argument("--shard-width", type=int, default=65536, help="N > 1: columns of the sharded city")
    ap.add_argument("--shard-rows", type=int, default=8192, help="N > 1: rows per GPU of the sharded city")
    ap.add_argument("--no-graph", action="store_true", help="N = 1: time eager launches instead of one CUDA graph per city")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the shifted-cut / single-GPU digest comparison")
    ap.add_argument("--no-legs", action="store_true", help="N = 1: only the layout legs (no vehicle-tick / route-planning child processes)")
    ap.add_argument("--leg", choices=sorted(LEGS), help="run one of the legs beside the headline alone and print its JSON object")
    args = ap.parse_args()
    if args.leg:
        return leg_main(args.leg)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
