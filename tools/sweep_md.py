#!/usr/bin/env python
"""Markdown table of the config-5 sweep from a run_configs.py JSON (fraction of the binding roofline per cell).
   python tools/sweep_md.py profiles/r2_configs.json"""
import json
import statistics
import sys

d = json.load(open(sys.argv[1]))
d = d["results"] if isinstance(d, dict) else d
r5 = [r for r in d if r.get("config") == 5][0]
cell = {(r["k"], r["P"], r["s"]): r for r in r5["rows"]}
print(f"HBM peak {r5['hbm_peak_gbs']:.0f} GB/s, FP32 peak {r5['fp32_peak_tflops']:.1f} TFLOP/s ({r5['fp32_peak_source']})\n")
print("| factor | k | flop/B | bound | P = 64 | P = 128 | P = 256 | P = 512 |")
print("|---|---|---|---|---|---|---|---|")
allb = []
for s in (2, 4, 8):
    for k in (11, 13, 15, 21, 31):
        row = []
        for p in (64, 128, 256, 512):
            r = cell[(k, p, s)]
            b = max(r["hbm_frac"], r["fp32_frac"])
            allb.append((b, k, p, s))
            row.append(f"{b:.2f} ({r['algo']})")
        r = cell[(k, 256, s)]
        bound = "HBM" if r["hbm_frac"] >= r["fp32_frac"] else "FP32"
        print(f"| {s} | {k} | {r['flop_per_byte']:.1f} | {bound} | " + " | ".join(row) + " |")
b = [x[0] for x in allb]
print(f"\nmedian {statistics.median(b):.3f}; cells >= 0.70: {sum(x >= 0.7 for x in b)}, >= 0.50: {sum(x >= 0.5 for x in b)}, < 0.35: {sum(x < 0.35 for x in b)} of {len(b)}")
for s in (2, 4, 8):
    v = [x[0] for x in allb if x[3] == s]
    print(f"factor {s}: median {statistics.median(v):.3f}, min {min(v):.3f}, max {max(v):.3f}")
v = [x[0] for x in allb if x[2] == 64]
print(f"64-wide patches: median {statistics.median(v):.3f}, min {min(v):.3f}")
