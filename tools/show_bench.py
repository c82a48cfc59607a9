#!/usr/bin/env python
"""Short summary of a bench.py JSON line."""
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as ex:
    print("bench failed:", ex); sys.exit(0)
r = lambda x: round(x, 3) if isinstance(x, float) else x
print("ms/step", r(d["ms_per_step"]), "value", round(d["value"]), "e2e ms", r(d.get("e2e", {}).get("ms_per_step", 0.0)),
      "launches", d.get("gpu_launches"), "n_gpus", d["n_gpus"], d["config"].get("config"))
if "kernels" in d:
    print("  kernels", {k: r(v["ms"]) for k, v in d["kernels"].items()})
if "roofline" in d:
    print("  roofline", {k: r(v) for k, v in d["roofline"].items() if k in ("achieved", "frac", "ms_per_launch")})
for k in ("alt_falloff", "comm", "contact"):
    if k in d:
        print(" ", k, {a: r(b) for a, b in d[k].items()})
if "sweep" in d:
    for row in d["sweep"]:
        print("  ", {a: r(b) for a, b in row.items()})
if "cpu_baseline" in d:
    print("  cpu", r(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"], "cores")
if "e2e" in d:
    print("  e2e numa", d["e2e"].get("numa"))
