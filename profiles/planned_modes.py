"""Wall time of bench.py's planned_trips workload (GPU arm only) under the three launch forms of tsim_astar_batch.
    python profiles/planned_modes.py [ticks] > gpurun_out/r2_planned_modes.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
from golden_util import load_ticks  # noqa: E402
from trafficsimulation_b200 import tapes  # noqa: E402
from trafficsimulation_b200.layout import GpuCityLayout  # noqa: E402
from trafficsimulation_b200.replan import PlannedTraffic  # noqa: E402
from trafficsimulation_b200.traffic import light_tables_from_layout  # noqa: E402

n_ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 100
r = load_ticks(os.path.join(ROOT, "tests", "golden", "ticks_default12345.npz"))
W, H = r["W"], r["H"]
cfgd = dict(r["meta"]["cfg"])
carve = cfgd.pop("carve_subblock_roads", False)
city = GpuCityLayout(carve_subblock_roads=carve, **cfgd)
city.set_bands(r["hbands"], r["vbands"])
city.generate(r["tape_zone"], r["tape_carve"], r["tape_entrance"])
tabs = light_tables_from_layout(city)
maps, planes = city.maps_host(), city.planes_host()
tp = tapes.synth_planned_trips(1, W, H, planes["cell_type"], 20, n_ticks, malfunction_p=0.002)
out = {}
ref = None
for mode in ("2", "1", "0"):
    os.environ["TSIM_ASTAR_MODE"] = mode
    sim = PlannedTraffic.on_gpu(W, H, tabs, tp, n_ticks, maps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.step(n_ticks)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ev = sim.events
    ref = ref or ev
    out[mode] = {"wall_s": round(wall, 3), "searches": sim.searches, "rounds": sim.batches, "same_routes_as_first": ev == ref}
    print(mode, out[mode], file=sys.stderr, flush=True)
print(json.dumps({"planned_trips_ticks": n_ticks, "by_TSIM_ASTAR_MODE": out}))
