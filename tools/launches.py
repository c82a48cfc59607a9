#!/usr/bin/env python
"""Aggregate an ncu launch list (gpu__time_duration.sum csv) per kernel."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try: agg[row['Kernel Name'][:60]].append(float(row['Metric Value']))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{k:60s} n={len(v):4d} total={sum(v)/1e6:8.3f} ms mean={sum(v)/len(v)/1e3:8.1f} us min={min(v)/1e3:8.1f}")
