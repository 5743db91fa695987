#!/usr/bin/env python
"""DRAM traffic and kernel time per PASS of one city, from an ncu launch list of profiles/run_layout.py:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/run_layout.py --size 16384 --reps 2
    python profiles/pass_traffic.py gpurun_out/launches.csv profiles/r2_pass_traffic_16384.json

The launches of the LAST generated city are cut into passes by the kernels that open them (the pipeline order is fixed:
frame, carve, zones, dead ends, R2 upgrade, entrances, direction fixes, lights, maps).  bench.py reads the JSON for
`roofline.traffic` and for the name / share of the top kernel of every pass -- no constants in bench.py."""
import collections
import json
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from summarize_metrics import load   # noqa: E402

OPENERS = [("frame_roads", "frame_roads"), ("frame_copy", "frame_roads"), ("dead_ends_kernel", "dead_ends"), ("upgrade_r2_kernel", "upgrade_r2"), ("ent_bits_kernel", "entrances"),
           ("validate_dirs_kernel", "fix_dirs"), ("init_pivot_kernel", "lights"), ("lights_bits_kernel", "lights"), ("maps_kernel", "maps")]


def main(src, dst):
    per = collections.OrderedDict()   # launch id -> [name, us, rd, wr], in launch order
    for _id, name, metric, val in load(src):
        e = per.setdefault(int(_id), [name, 0.0, 0.0, 0.0])
        e[1 if metric.startswith("gpu__time") else (2 if "read" in metric else 3)] += val
    launches = [per[k] for k in sorted(per)]
    start = max(i for i, l in enumerate(launches) if "frame_roads" in l[0] or "frame_copy" in l[0])
    passes = collections.OrderedDict()
    cur, labellings = None, 0
    for name, us, rd, wr in launches[start:]:
        short = name.replace("void ", "").replace("tsim::", "")
        if "ccl_bits_kernel" in name:
            labellings += 1
            cur = "carve" if labellings == 1 else "zones"
        for key, p in OPENERS:
            if key in name:
                cur = p
        if cur is None or "at::" in name:
            continue
        d = passes.setdefault(cur, {"us": 0.0, "dram_bytes": 0.0, "launches": 0, "kernels": collections.OrderedDict()})
        d["us"] += us; d["dram_bytes"] += rd + wr; d["launches"] += 1
        k = d["kernels"].setdefault(short, {"us": 0.0, "launches": 0, "dram_bytes": 0.0})
        k["us"] += us; k["launches"] += 1; k["dram_bytes"] += rd + wr
    for d in passes.values():
        top = max(d["kernels"], key=lambda k: d["kernels"][k]["us"])
        d["top_kernel"] = top
        d["top_kernel_share"] = round(d["kernels"][top]["us"] / d["us"], 4)
        d["us"] = round(d["us"], 1)
        for k in d["kernels"].values():
            k["us"] = round(k["us"], 1)
    out = {"source": src, "note": "ncu launch list (per-launch times are serialised and cold-cache: shares, not absolutes); dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum",
           "passes": passes}
    json.dump(out, open(dst, "w"), indent=1)
    for p, d in passes.items():
        print(f"{p:12s} {d['us']:8.1f} us {d['launches']:3d} launches {d['dram_bytes'] / 1e6:9.1f} MB  top {d['top_kernel']} ({d['top_kernel_share']:.0%})")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
