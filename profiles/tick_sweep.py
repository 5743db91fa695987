#!/usr/bin/env python
"""ms/tick of the live-list tick kernel for TSIM_TICK_CTAS_PER_SM = 1, 2, 4, 8 at both fleet sizes of bench.py."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
for k in ("1", "2", "4", "8"):
    os.environ["TSIM_TICK_CTAS_PER_SM"] = k
    a = bench.vehicle_bench(dev, cpu_ticks=0, parity_check=False, e2e_ticks=2)
    b = bench.vehicle_bench(dev, n_ticks=60, size=8192, n_vehicles=1000000, cpu_ticks=0, route_len=100, e2e_ticks=2, parity_check=False)
    print("ctas/sm", k, "| 100k ms/tick", round(a["ms_per_tick"], 4), f"{a['value']:.3e}", "| 1M ms/tick", round(b["ms_per_tick"], 4), f"{b['value']:.3e}", flush=True)
