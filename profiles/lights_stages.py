#!/usr/bin/env python
"""Time the stages of the lights pass (prepare / seed+reach / finish) and print the number of reach alternations."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trafficsimulation_b200 import tapes
from trafficsimulation_b200.sharded import ShardedCityLayout
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=4096); a = ap.parse_args()
dev = torch.device("cuda", 0)
hb, vb = tapes.synth_bands(4096, width=a.size, height=a.size)
sh = ShardedCityLayout(1, width=a.size, height=a.size, carve_subblock_roads=True, device=dev)
sh.set_bands(hb, vb)
tz = torch.from_numpy(tapes.synth_zone_tape(4096, sh.global_cap)).to(dev)
te = torch.zeros(sh.global_cap, dtype=torch.int32, device=dev)
tc = sh.synth_carve_tapes(4096)
L = sh.shards[0]
for rep in range(3):
    sh.generate(tz, tc, te, check=False, lights=False, maps=False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); L._lights_prepare(); ev[1].record()
    L.flags[10] = L.flags[8]
    L._lights_seed(); L._lights_reach(); ev[2].record()
    L._lights_finish(check=False); ev[3].record()
    torch.cuda.synchronize()
    alt = int(L.workspace[:256].view(torch.int32)[18].item())   # alternations of the per-phase launches
    print(f"size {a.size}: prepare {ev[0].elapsed_time(ev[1]):.3f} ms, reach {ev[1].elapsed_time(ev[2]):.3f} ms ({alt} alternations), finish {ev[2].elapsed_time(ev[3]):.3f} ms")
L._check_flag("lights_stages")
