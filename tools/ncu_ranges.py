#!/usr/bin/env python
"""Instruction / sample share per source-line range: ncu_ranges.py rep name:lo-hi ..."""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iInst = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples"); iThr = hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
L = {}
ST = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[2] != "-": continue
    try:
        L[int(r[0])] = (int(r[iInst]), int(r[iSamp]), int(r[iThr]))
        ST[int(r[0])] = [int(r[i] or 0) for i in stall_cols]
    except ValueError: pass
tot = sum(v[0] for v in L.values()); tots = sum(v[1] for v in L.values())
for spec in sys.argv[2:]:
    name, rg = spec.split(":"); lo, hi_ = map(int, rg.split("-"))
    a = sum(v[0] for k, v in L.items() if lo <= k <= hi_); b = sum(v[1] for k, v in L.items() if lo <= k <= hi_)
    c = sum(v[2] for k, v in L.items() if lo <= k <= hi_)
    st = [sum(ST[k][j] for k in ST if lo <= k <= hi_) for j in range(len(stall_cols))]
    top = sorted(zip(st, [hdr[i][6:] for i in stall_cols]), reverse=True)[:5]
    tt = sum(st) or 1
    print(f"{name:14s} L{lo}-{hi_}: {100*a/tot:5.1f}% inst  {100*b/tots:5.1f}% samples  thr/inst {c/max(a,1):4.1f}  | " +
          ", ".join(f"{n} {100*v/tt:.0f}%" for v, n in top))
