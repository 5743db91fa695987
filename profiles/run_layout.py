#!/usr/bin/env python
"""Minimal driver for profiling: generates one size x size synthetic city `reps` times (same inputs as bench.py).

    python profiles/run_layout.py --size 16384 --reps 2            # plain run (must exit 0 before any ncu run)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/run_layout.py --size 16384 --reps 2
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trafficsimulation_b200 import tapes                      # noqa: E402
from trafficsimulation_b200.layout import GpuCityLayout       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--seed", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hb, vb = tapes.synth_bands(a.seed, width=a.size, height=a.size)
cap = 3 * (len(hb) + 2) * (len(vb) + 2) + 64
tz, te = tapes.synth_zone_tape(a.seed, cap), np.zeros(cap, np.int32)
city = GpuCityLayout(width=a.size, height=a.size, carve_subblock_roads=True, device=dev)
city.set_bands(hb, vb)
city._build_roads_and_sidewalks()
n_blobs, table = city.label_nothing()
tc = tapes.synth_carve_tape(a.seed, table.cpu().numpy())
d_tz, d_te, d_tc = torch.from_numpy(tz).to(dev), torch.from_numpy(te).to(dev), torch.from_numpy(tc).to(dev)
for _ in range(a.reps):
    city.generate(d_tz, d_tc, d_te, check=False)
torch.cuda.synchronize()
city._check_flag("run_layout")
print("ok", a.size, int(city.flags[2].item()), "blocks", int(city.flags[3].item()), "lights", city.sweeps(), "sweeps")
