#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS view) export by CUDA source line.

    nvdisasm -g -c <cubin> > dis.txt ; ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_by_line.py dis.txt sass.csv <kernel symbol substring> [top N]

Developer tool (profiles/ summaries are produced with it); not part of the product."""
import csv
import re
import sys
from collections import defaultdict


def line_map(dis, sym):
    amap, cur, active = {}, None, False
    for ln in open(dis):
        if ln.startswith("\t.section") or ln.startswith(".section"):
            active = sym in ln
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2))
    return amap


def main():
    dis, sass, sym = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(dis, sym)
    rows = list(csv.reader(open(sass)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = defaultdict(lambda: defaultdict(float))
    base = None
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        if base is None:
            base = a
        key = amap.get(a - base, (None, ""))[0]
        g = agg[key]
        g["inst"] += float(r[col["Instructions Executed"]] or 0)
        g["samples"] += float(r[col["# Samples"]] or 0)
        g["smem_wave"] += float(r[col["L1 Wavefronts Shared"]] or 0)
        for s in stall_cols:
            g[s] += float(r[col[s]] or 0)
    tot_i = sum(g["inst"] for g in agg.values())
    tot_s = sum(g["samples"] for g in agg.values())
    print(f"total warp-instructions {tot_i:.3e}, samples {tot_s:.0f}")
    print(f"{'line':>22} {'inst%':>6} {'smp%':>6} {'smemW%':>7}  top stalls")
    tot_w = sum(g["smem_wave"] for g in agg.values()) or 1
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((g[s], s) for s in stall_cols), reverse=True)[:3]
        ss = " ".join(f"{n[6:]}:{v / max(g['samples'], 1):.2f}" for v, n in st if v > 0)
        name = f"{key[0]}:{key[1]}" if key else "?"
        print(f"{name:>22} {100 * g['inst'] / tot_i:6.2f} {100 * g['samples'] / tot_s:6.2f} {100 * g['smem_wave'] / tot_w:7.2f}  {ss}")


if __name__ == "__main__":
    main()
