#!/usr/bin/env python
"""The 1 M-vehicle tick leg of bench.py alone (8192 x 8192 city, live-list kernel, whole state compared with the C oracle), with
bench.py's progress log on stderr.  Prints one JSON object (bench.vehicle_bench's)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
t0 = time.time()
r = bench.vehicle_bench(torch.device("cuda", 0), n_ticks=60, size=8192, n_vehicles=1000000, cpu_ticks=0, route_len=100, e2e_ticks=20,
                        parity_check="--no-parity" not in sys.argv)
r["wall_s"] = round(time.time() - t0, 1)
print(json.dumps(r))
