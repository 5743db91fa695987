#!/usr/bin/env python
"""Per-kernel table from an `ncu --set full` report: launches, time, DRAM bytes, instructions, issue slots.

    ncu -i gpurun_out/full16k.ncu-rep --page raw --csv > /tmp/raw.csv ; python profiles/summarize_full.py /tmp/raw.csv [skip_first_n]
"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "inst": 1, "%": 1}


def val(r, name):
    i = ix[name]
    return float(r[i].replace(",", "")) * scale.get(units[i], 1)


agg = collections.OrderedDict()
for r in rows[2 + skip:]:
    k = r[ix["Kernel Name"]].split("(")[0]
    a = agg.setdefault(k, dict(n=0, us=0.0, rd=0.0, wr=0.0, inst=0.0, issue=0.0))
    a["n"] += 1
    a["us"] += val(r, "gpu__time_duration.sum")
    a["rd"] += val(r, "dram__bytes_read.sum")
    a["wr"] += val(r, "dram__bytes_write.sum")
    a["inst"] += val(r, "smsp__inst_executed.sum")
    a["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
tot = sum(a["us"] for a in agg.values())
print(f"{'kernel':40s} {'n':>3s} {'us':>9s} {'share':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'GB/s':>7s} {'Minst':>8s} {'issue%':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k[:40]:40s} {a['n']:3d} {a['us']:9.1f} {a['us'] / tot * 100:5.1f}% {a['rd'] / 1e6:11.1f} {a['wr'] / 1e6:11.1f} "
          f"{(a['rd'] + a['wr']) / 1e3 / a['us']:7.0f} {a['inst'] / 1e6:8.1f} {a['issue'] / a['n']:6.1f}")
print(f"{'TOTAL':40s} {sum(a['n'] for a in agg.values()):3d} {tot:9.1f} {'':6s} {sum(a['rd'] for a in agg.values()) / 1e6:11.1f} {sum(a['wr'] for a in agg.values()) / 1e6:11.1f}")
