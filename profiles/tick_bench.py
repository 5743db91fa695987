#!/usr/bin/env python
"""Vehicle-tick leg of bench.py alone (both fleet sizes), for quick iteration."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
a = bench.vehicle_bench(dev, cpu_ticks=0)
b = bench.vehicle_bench(dev, n_ticks=60, size=8192, n_vehicles=1000000, cpu_ticks=0, route_len=100, e2e_ticks=20)
for v in (a, b):
    print(v["config"]["workload"], "ms/tick", round(v["ms_per_tick"], 4), "updates/s", f"{v['value']:.3e}", "frac", v["roofline"]["frac"],
          "sweeps", round(v["fixed_point_iterations_per_tick"], 2), "e2e", f"{v['e2e']['value']:.3e}")
