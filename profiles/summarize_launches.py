"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py profiles/r1_launches_4096_a.csv
"""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        val = float(d["Metric Value"].replace(",", ""))
        val = {"ns": val / 1e3, "us": val, "ms": val * 1e3, "s": val * 1e6}.get(d["Metric Unit"], val)
        agg.setdefault(d["Kernel Name"].split("(")[0], []).append(val)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':62s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:62]:62s} {len(v):4d} {sum(v):10.1f} {sum(v) / len(v):9.1f} {sum(v) / tot * 100:5.1f}%")
    print(f"{'TOTAL':62s} {sum(len(v) for v in agg.values()):4d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
