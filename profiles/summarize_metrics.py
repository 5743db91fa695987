"""Per-kernel summary of an ncu CSV launch list holding gpu__time_duration.sum and (optionally) dram__bytes_read.sum /
dram__bytes_write.sum per launch:  python profiles/summarize_metrics.py <csv> [reps]   (totals are divided by `reps`)."""
import collections
import csv
import sys


def load(path):
    hdr, rows = None, []
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        val = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        if d["Metric Name"].startswith("gpu__time"):
            val = {"ns": val / 1e3, "us": val, "ms": val * 1e3, "s": val * 1e6}.get(unit, val)            # -> us
        else:
            val = {"byte": val, "Kbyte": val * 1e3, "Mbyte": val * 1e6, "Gbyte": val * 1e9}.get(unit, val)  # -> bytes
        rows.append((d["ID"], d["Kernel Name"].split("(")[0], d["Metric Name"], val))
    return rows


def main(path, reps=1):
    agg = collections.OrderedDict()
    for _id, k, m, v in load(path):
        a = agg.setdefault(k, {"n": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"].add(_id)
        a["us" if m.startswith("gpu__time") else ("rd" if "read" in m else "wr")] += v
    tot = sum(a["us"] for a in agg.values())
    print(f"{'kernel':58s} {'n':>5s} {'us/rep':>9s} {'avg us':>8s} {'share':>6s} {'MB rd/rep':>10s} {'MB wr/rep':>10s} {'GB/s':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = len(a["n"])
        gbs = (a["rd"] + a["wr"]) / a["us"] / 1e3 if a["us"] else 0
        print(f"{k[:58]:58s} {n / reps:5.1f} {a['us'] / reps:9.1f} {a['us'] / n:8.1f} {a['us'] / tot * 100:5.1f}% {a['rd'] / reps / 1e6:10.1f} {a['wr'] / reps / 1e6:10.1f} {gbs:7.0f}")
    print(f"{'TOTAL':58s} {sum(len(a['n']) for a in agg.values()) / reps:5.1f} {tot / reps:9.1f} {'':8s} {'':6s} "
          f"{sum(a['rd'] for a in agg.values()) / reps / 1e6:10.1f} {sum(a['wr'] for a in agg.values()) / reps / 1e6:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1)
