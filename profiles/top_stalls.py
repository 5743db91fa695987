#!/usr/bin/env python
"""Where a kernel's warps wait: top SASS instructions by stall samples from an `ncu --set full --import-source on` report.

    python profiles/top_stalls.py gpurun_out/pass_full.ncu-rep <kernel regex> [launch_skip] [top_n]
"""
import collections, csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:110])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
seen, uniq = set(), []
for r in data:                     # the page lists the kernel twice
    if r[ix["Address"]] in seen:
        continue
    seen.add(r[ix["Address"]]); uniq.append(r)


def g(r, k):
    try:
        return int(float(r[ix[k]] or 0))
    except ValueError:
        return 0


tot = sum(g(r, "# Samples") for r in uniq)
inst = sum(g(r, "Instructions Executed") for r in uniq)
print(f"{len(uniq)} SASS instructions, {inst / 1e6:.2f} M warp instructions executed, {tot} stall samples")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter({h: sum(g(r, h) for r in uniq) for h in reasons})
print("by reason:", ", ".join(f"{h[6:]} {v * 100 // max(tot, 1)}%" for h, v in agg.most_common(6)))
for i, r in sorted(enumerate(uniq), key=lambda t: -g(t[1], "# Samples"))[:topn]:
    why = max(reasons, key=lambda h: g(r, h))
    print(f"  #{i:5d} {g(r, '# Samples') * 100 / max(tot, 1):5.1f}%  {why[6:]:10s} exec {g(r, 'Instructions Executed'):9d}  {r[ix['Source']][:70]}")
