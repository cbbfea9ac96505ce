#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list per kernel (developer tool for profiles/).
    python tools/ncu_launch_list.py launches.csv "header comment" > profiles/<name>.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_ms", "share_pct"])
w.writerow(["# " + (sys.argv[2] if len(sys.argv) > 2 else "") + f"; {sum(v[0] for v in agg.values())} launches, {tot:.1f} ms in the profiled window (ncu times are cold-cache and serialised: shares, not absolutes)"])
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    w.writerow([name, n, f"{ms:.4f}", f"{100 * ms / tot:.2f}"])
