"""Generates tests/golden/ticks_*.npz from the LIVE reference tick loop (build container only).

    python tests/golden/make_golden_ticks.py [case ...]

Each fixture holds the layout inputs (cfg, bands, layout tapes), the reference's light tables, the tick
tapes (spawn attempts, speed / malfunction / rank tapes, route events) and the per-tick vehicle and map
states of the unmodified reference under those tapes (oracle/refharness/ticks.py).
Route events are stored as a first cell plus 2-bit direction steps to keep the files small.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle.refharness import ticks  # noqa: E402
from make_golden import dense_tapes, model_cfg  # noqa: E402

CASES = {
    "default12345": dict(seed=12345, n_ticks=240, spawns_per_tick=6, malfunction_p=0.002),
    "s7_rain": dict(seed=7, n_ticks=100, spawns_per_tick=10, malfunction_p=0.0, rain_rect=(40, 40, 160, 120)),
    "s14_carve": dict(seed=14, n_ticks=100, spawns_per_tick=4, malfunction_p=0.01,
                      layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
    # sideswipe draws that fire (vehicle_base.py:567-605; stored as a second bit plane beside the malfunction tape)
    "s21_sideswipe": dict(seed=21, n_ticks=140, spawns_per_tick=12, malfunction_p=0.002, sideswipe_p=0.35),
    # many malfunctions: stranded vehicles on the lanes -> contraflow overtakes and stuck detours (vehicle_base.py:305-418), the case the
    # route-planning loop (trafficsimulation_b200/replan.py) is checked on without the route events
    "s5_stranded": dict(seed=5, n_ticks=90, spawns_per_tick=8, malfunction_p=0.03),
    # the other light controllers (Defaults.TRAFFIC_LIGHT_AGENT_ALGORITHM, intersection_light_group.py:396-461)
    "s31_fixed_time": dict(seed=31, n_ticks=100, spawns_per_tick=8, malfunction_p=0.002, algo="FIXED_TIME"),
    "s9_pressure": dict(seed=9, n_ticks=120, spawns_per_tick=8, malfunction_p=0.002, algo="PRESSURE_CONTROL"),
    "s9_green_wave": dict(seed=9, n_ticks=120, spawns_per_tick=8, malfunction_p=0.002, algo="NEIGHBOR_GREEN_WAVE"),
    "s14_green_wave": dict(seed=14, n_ticks=80, spawns_per_tick=4, malfunction_p=0.01, algo="NEIGHBOR_GREEN_WAVE",
                           layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True)),
    "s14_pressure_fwd": dict(seed=14, n_ticks=80, spawns_per_tick=4, malfunction_p=0.01, algo="PRESSURE_CONTROL",
                             layout_kwargs=dict(width=150, height=110, carve_subblock_roads=True, forward_traffic_light_range=True)),
}


def encode_routes(ev_off, ev_cells, W):
    first = np.full(len(ev_off) - 1, -1, np.int32)
    steps = []
    for i in range(len(ev_off) - 1):
        c = ev_cells[ev_off[i]:ev_off[i + 1]]
        if len(c) == 0:
            continue
        first[i] = c[0]
        d = np.diff(c)
        code = np.select([d == W, d == 1, d == -W, d == -1], [0, 1, 2, 3], default=255).astype(np.uint8)
        assert not (code == 255).any(), "non-adjacent path step"
        steps.append(code)
    return first, (np.concatenate(steps) if steps else np.zeros(0, np.uint8))


def csr(lists):
    off = np.zeros(len(lists) + 1, np.int32)
    off[1:] = np.cumsum([len(a) for a in lists])
    return off, (np.concatenate(lists) if len(lists) and off[-1] else np.zeros(0, np.int32)).astype(np.int32)


def main():
    for name, case in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        r = ticks.run_ticks(**case)
        lay = r["layout"]
        model = lay["model"]
        zone, run = dense_tapes(lay)
        first, steps = encode_routes(r["ev_off"], r["ev_cells"], r["W"])
        meta = dict(case={k: v for k, v in case.items()}, cfg=model_cfg(model), n_blocks=len(model._blocks_data),
                    W=r["W"], H=r["H"], n_ticks=r["n_ticks"], n_attempts=r["n_attempts"], rain_enabled=case.get("rain_rect") is not None)
        arrays = dict(
            meta=np.frombuffer(json.dumps(meta).encode(), np.uint8),
            hbands=lay["hbands"], vbands=lay["vbands"], tape_zone=zone, tape_carve=lay["tape_carve"], tape_entrance=run,
            links_lights=lay["links"]["lights"], links_ctrl=lay["links"]["ctrl"],
            spawn_tick=r["spawn_tick"], origin=r["origin"], target=r["target"], spawned=r["spawned"],
            speed=r["speed"], malfunction=np.packbits(r["malfunction"] & 1, axis=1), rank=r["rank"],
            rain_map=np.packbits(r["rain_map"], axis=1),
            ev_tick=r["ev_tick"], ev_vehicle=r["ev_vehicle"], ev_len=np.diff(r["ev_off"]).astype(np.int32),
            ev_first=first, ev_steps=steps,
            pos=r["pos"], base_speed=r["base_speed"], stuck_ticks=r["stuck_ticks"], vflags=r["vflags"],
            group_state=r["group_state"].astype(np.int16),
        )
        if case.get("sideswipe_p"):
            arrays["sideswipe"] = np.packbits((r["malfunction"] >> 1) & 1, axis=1)
        for k in ("occ", "stop", "stuckmap"):
            arrays[k + "_off"], arrays[k + "_cells"] = r[k + "_off"], r[k + "_cells"]
        for f in ("cluster", "lights", "ns_lights", "ew_lights", "ns_in", "ew_in") + (("ns_out", "ew_out") if case.get("algo") else ()):
            arrays["g_" + f + "_off"], arrays["g_" + f] = csr([g[f] for g in r["groups"]])
        if case.get("algo"):   # fixed-time timer / phase, pressures (older fixtures stay byte-identical without them)
            arrays["group_ext"] = r["group_ext"].astype(np.int16)
            # neighbour links of the reference (canonical group indices, N S E W) and the order it created its groups in
            arrays["g_nbr"] = np.stack([g["nbr"] for g in r["groups"]]).astype(np.int32)
            arrays["g_creation_rank"] = np.array([int(g["creation_rank"][0]) for g in r["groups"]], np.int32)
        path = os.path.join(HERE, f"ticks_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, os.path.getsize(path) // 1024, "KiB", "spawned", int(r["spawned"].sum()), "events", len(r["ev_tick"]))


if __name__ == "__main__":
    main()
