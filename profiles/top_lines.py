#!/usr/bin/env python
"""Top CUDA source lines of a kernel by stall samples (report captured with --import-source on, built with -lineinfo).

    python profiles/top_lines.py gpurun_out/pass_full.ncu-rep <kernel regex> [launch_skip] [top_n]
"""
import csv, io, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, best = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 3 and r[0] == "Line No":
        hdr = r
        i_s, i_long = hdr.index("# Samples"), hdr.index("stall_long_sb")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            n, nl = int(float(r[i_s] or 0)), int(float(r[i_long] or 0))
        except ValueError:
            continue
        if n:
            best[(fname, int(r[0]))] = (n, nl, r[1].strip())
tot = sum(b[0] for b in best.values())
print(rx, "stall samples", tot)
for (f, ln), (n, nl, src) in sorted(best.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"  {n * 100 / max(tot, 1):5.1f}% (long_sb {nl * 100 // max(n, 1):3d}%)  {f}:{ln:<5d} {src[:100]}")
