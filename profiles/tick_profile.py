#!/usr/bin/env python
"""One persistent tick launch (100 k vehicles, 2048^2 city) for ncu."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from trafficsimulation_b200 import tapes
from trafficsimulation_b200.layout import GpuCityLayout
from trafficsimulation_b200.traffic import GpuTraffic, light_tables_from_layout
dev = torch.device("cuda", 0)
size, seed, nt = (8192, 2048, 30) if len(sys.argv) > 1 and sys.argv[1] == "1M" else (2048, 2048, 100)
nveh, rlen = (1000000, 100) if size == 8192 else (100000, 400)
hb, vb, cap, tz, te = bench.synth_inputs(size, seed)
city = GpuCityLayout(width=size, height=size, device=dev)
city.set_bands(hb, vb)
city.generate(tz, None, te)
tabs = light_tables_from_layout(city)
pl = city.planes_host()
tp = tapes.synth_traffic(seed, size, size, pl["cell_type"], pl["dirs"], nveh, nt, route_len=rlen, spawn_ticks=1)
sim = GpuTraffic(size, size, tabs, tp, nt, device=dev)
sim.step(5)
sim.step(nt - 10, check=False)
torch.cuda.synchronize()
print("ok", sim.counters())
