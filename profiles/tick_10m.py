#!/usr/bin/env python
"""BASELINE.json configs[4]'s fleet on ONE GPU: 10 M vehicles live at once on a 16384 x 16384 city (live-list tick kernel), the
whole state compared with the C oracle at the end.  Prints one JSON object (bench.vehicle_bench's)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
t0 = time.time()
r = bench.vehicle_bench(torch.device("cuda", 0), n_ticks=25, size=16384, n_vehicles=10_000_000, cpu_ticks=2, route_len=60, e2e_ticks=5,
                        parity_check="--no-parity" not in sys.argv)
r["wall_s"] = round(time.time() - t0, 1)
print(json.dumps(r))
