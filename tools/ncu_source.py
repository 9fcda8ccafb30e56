#!/usr/bin/env python
"""Per-opcode and top-instruction stall attribution from `ncu --page source --csv` of a report.
   python tools/ncu_source.py report.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(h)]
ix = {k: i for i, k in enumerate(h)}
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_samples = sum(f(r, "# Samples") for r in body)
tot_inst = sum(f(r, "Instructions Executed") for r in body)
print(f"instructions executed {tot_inst:.0f}, samples {tot_samples:.0f}")
per_op = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in body:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    e = per_op[op]
    e[0] += f(r, "Instructions Executed"); e[1] += f(r, "# Samples")
    for k in stall_cols:
        e[2][k[6:]] += f(r, k)
print("opcode      inst%  samples%  top stall reasons (share of this opcode's samples)")
for op, (ni, ns, c) in sorted(per_op.items(), key=lambda kv: -kv[1][1])[:14]:
    tops = ", ".join(f"{k} {v / max(ns, 1):.2f}" for k, v in c.most_common(4) if v)
    print(f"{op:10s} {100 * ni / tot_inst:6.1f} {100 * ns / tot_samples:8.1f}  {tops}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print("top instructions by samples:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:n]:
    c = {k[6:]: f(r, k) for k in stall_cols if f(r, k)}
    tops = ", ".join(f"{k} {v:.0f}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])[:3])
    print(f"{r[ix['Address']][-5:]} {100 * f(r, '# Samples') / tot_samples:5.2f}%  {r[ix['Source']][:70]:70s} {tops}")
