#!/usr/bin/env python
"""ms/tick of the live-list tick kernel at the fleet sizes of bench.py (default launch shape): tick_quick.py [100k] [1M] [--parity]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
par = "--parity" in sys.argv
keys = ("ms_per_tick", "value", "fixed_point_iterations_per_tick")
import ctypes as C
from trafficsimulation_b200 import _lib
lib = _lib.load()
def phases(tag, n_updates=None):
    a = (C.c_ulonglong * 16)()
    lib.tsim_debug_tick_phases(a, 1)
    tot = sum(a[:7]) or 1
    names = ("decide", "sideswipe", "sweep0", "sweeps", "move", "tile_scan", "spawn+lights")
    print(tag, "phase share:", {n: round(a[i] / tot, 3) for i, n in enumerate(names)}, "contested vehicle-ticks", a[7],
          "| thread 0 own share of total:", {n: round(a[i] / tot, 3) for i, n in ((8, "decide_veh"), (9, "decide_groups"), (11, "spawns"), (12, "commits"), (13, "events"))}, flush=True)
t0 = time.time()
if "100k" in sys.argv:
    a = bench.vehicle_bench(dev, cpu_ticks=10 if par else 0, parity_check=par, e2e_ticks=2)
    print("100k", {k: a[k] for k in keys if k in a}, a.get("parity"), round(time.time() - t0), "s", flush=True)
    phases("100k")
if "1M" in sys.argv:
    b = bench.vehicle_bench(dev, n_ticks=60, size=8192, n_vehicles=1000000, cpu_ticks=5 if par else 0, route_len=100, e2e_ticks=2, parity_check=par)
    print("1M", {k: b[k] for k in keys if k in b}, b.get("parity"), round(time.time() - t0), "s", flush=True)
    phases("1M")
