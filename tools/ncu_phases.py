#!/usr/bin/env python
"""Instructions / stall samples per source-line range: ncu_phases.py rep a-b:name [a-b:name ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
rngs = []
for a in sys.argv[2:]:
    r, name = a.split(":")
    lo, hi = r.split("-")
    rngs.append((int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi_]
iInst = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
acc = {n: [0, 0] for _, _, n in rngs}; acc["other"] = [0, 0]
for r in rows[hi_ + 1:]:
    if len(r) < len(hdr) or r[2] != "-": continue
    try: ln, ins, sm = int(r[0]), int(r[iInst]), int(r[iSamp])
    except ValueError: continue
    for lo, hi, n in rngs:
        if lo <= ln <= hi:
            acc[n][0] += ins; acc[n][1] += sm; break
    else:
        acc["other"][0] += ins; acc["other"][1] += sm
ti = sum(v[0] for v in acc.values()) or 1; ts = sum(v[1] for v in acc.values()) or 1
for n, (i, s) in acc.items():
    print(f"{n:16s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%   ({i/1e6:8.1f} M warp-inst)")
