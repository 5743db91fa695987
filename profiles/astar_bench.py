#!/usr/bin/env python
"""Throughput of the batched route planner on the reference's default city (tests/golden maps, 200 x 200): plain routes
between random road cells, CUDA events around GpuAstar.plan_cells; the C oracle on a sample of the same queries beside it.

    python profiles/astar_bench.py [n_queries]
"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import load_astar                      # noqa: E402
from oracle import oracle as O                          # noqa: E402
from trafficsimulation_b200.pathfinding import GpuAstar  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
r = load_astar(os.path.join(ROOT, "tests", "golden", "astar_default12345.npz"))
W, H = r["W"], r["H"]
rng = np.random.default_rng(1)
road = np.flatnonzero(r["is_road_map"].reshape(-1) == 1)
a, b = rng.choice(road, nq), rng.choice(road, nq)
q = np.stack([a % W, a // W, b % W, b // W, np.zeros(nq, np.int64), np.full(nq, 10), np.full(nq, 0x7FFFFFFF)], 1)
planner = GpuAstar(W, H, np.zeros((H, W), np.uint8), r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
planner.plan_cells(q[:256])                             # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
paths = planner.plan_cells(q)
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
ora = O.OracleAstar(np.zeros((H, W), np.uint8), r["stop_map"], r["is_road_map"], r["road_type_map"], r["allowed_dirs_map"], r["density"])
ns = 200
t0 = time.perf_counter()
ref = [ora.query(*[int(v) for v in q[i, :4]]) for i in range(ns)]
cpu = time.perf_counter() - t0
same = all([y * W + x for x, y in ref[i]] == paths[i].tolist() for i in range(ns))
print(json.dumps({"metric": "routes/s (batched A*, reference-exact)", "queries": nq, "grid": f"{W}x{H}", "with_route": int(sum(len(p) > 0 for p in paths)),
                  "mean_path_cells": float(np.mean([len(p) for p in paths])), "device_ms": e0.elapsed_time(e1), "value": nq / wall, "unit": "routes/s (host wall, results on the host)",
                  "cpu_oracle_routes_per_s": ns / cpu, "cpu_sample": f"{ns} of the same queries, 1 core, oracle/astar_oracle.c", "sample_matches_oracle": same}))
