#!/usr/bin/env python
"""Top source lines by instructions executed / stall samples: ncu_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iL, iS, iA = 0, 1, 2
iInst = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples"); iThr = hdr.index("Thread Instructions Executed")
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[iA] != "-": continue   # cuda lines have address "-"
    try: lines.append((int(r[iInst]), int(r[iSamp]), int(r[iThr]), r[iL], r[iS].strip()[:110]))
    except ValueError: pass
tot = sum(l[0] for l in lines) or 1; tots = sum(l[1] for l in lines) or 1
print(f"total warp-inst {tot:,}  samples {tots:,}")
for l in sorted(lines, key=lambda l: -l[1])[:top]:
    print(f"{100*l[0]/tot:5.1f}% inst {100*l[1]/tots:5.1f}% samp  thr/inst {l[2]/max(l[0],1):4.1f}  L{l[3]:>4s}  {l[4]}")
