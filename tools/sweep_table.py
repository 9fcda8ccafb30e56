#!/usr/bin/env python
"""Print the config-5 sweep of a run_configs.py JSON as a table (fraction of the binding roofline per cell)."""
import json
import sys

d = json.load(open(sys.argv[1]))
r5 = [r for r in d if r.get("config") == 5][0]
rows = r5["rows"]
hbm, fp = r5["hbm_peak_gbs"], r5["fp32_peak_tflops"]
print("k   P   s  algo    ms     hbm   fp32  bind | alternatives (bind)")
binds = []
for r in sorted(rows, key=lambda r: (r["s"], r["k"], r["P"])):
    b = max(r["hbm_frac"], r["fp32_frac"])
    alts = {a: b * r["ms"] / m for a, m in r.get("alt_ms", {}).items()}
    best = max([b] + list(alts.values()))
    binds.append((b, best, r))
    print(f'{r["k"]:2d} {r["P"]:4d} {r["s"]:2d} {r["algo"]:7s} {r["ms"]:6.3f} {r["hbm_frac"]:5.3f} {r["fp32_frac"]:5.3f} {b:5.3f} | '
          + " ".join(f"{a}={v:5.3f}" for a, v in alts.items()))
import statistics
print("median(auto)", round(statistics.median(b for b, _, _ in binds), 3), "median(best)", round(statistics.median(b for _, b, _ in binds), 3))
for s in (2, 4, 8):
    v = [b for b, _, r in binds if r["s"] == s]
    print("factor", s, "median", round(statistics.median(v), 3), "min", round(min(v), 3), "max", round(max(v), 3))
print("P=64 min", round(min(b for b, _, r in binds if r["P"] == 64), 3))
