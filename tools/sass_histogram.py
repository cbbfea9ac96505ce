#!/usr/bin/env python
"""SASS opcode histogram of the shipped library (developer tool for profiles/): which kernels contain tcgen05 / TMA / TMEM
instructions.  Runs anywhere cuobjdump is installed (no GPU needed).
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "v-detr_b200", "lib", "libvdetr_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "HMMA", "LDSM", "REDG", "RED.", "ATOMS", "SYNCS", "FFMA2", "FMUL2",
         "MUFU.LG2", "MUFU.EX2", "BAR.SYNC", "LDS", "STS", "MEMBAR", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per, total, cur = collections.OrderedDict(), collections.Counter(), None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = [0, collections.Counter()]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        per[cur][0] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][1][w] += 1
                total[w] += 1
                break
print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a), tools/sass_histogram.py")
print("# tcgen05 = UTCHMMA (MMA) / UTCBAR (commit); TMA = UTMALDG (tensor) / UBLKCP (bulk); TMEM = LDTM / STTM; mma.sync = HMMA; ldmatrix = LDSM")
print("# whole library: " + ", ".join(f"{k}={v}" for k, v in total.items()))
print()
for fn, (n, c) in per.items():
    label = fn[:40]
    for m in re.finditer(r"\d+", fn):             # Itanium mangling: <length><identifier>; take the identifier that names a kernel
        for k in range(len(m.group(0))):
            L = int(m.group(0)[k:])
            ident = fn[m.end():m.end() + L]
            if len(ident) == L and re.fullmatch(r"[a-z_][a-z0-9_]*", ident) and "kernel" in ident:
                label = ident
    print(f"{label:36s} [{fn[:60]}] {n} instructions: " + ", ".join(f"{k}={v}" for k, v in c.items()))
