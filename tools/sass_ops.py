#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libkmsr.so (cuobjdump -sass): what proves the kernels are sm_100a code
(UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 / FADD2 = packed fp32, HMMA =
mma.sync of the selector's fallback path; UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMASTG = TMA store
of the selector's tcgen05 path).  python tools/sass_ops.py > profiles/sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "kernel-modeling-super-resolution_b200", "libkmsr.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs = re.split(r"\n\s*Function : ", txt)[1:]
print(f"# {os.path.relpath(so, ROOT)}: {len(funcs)} kernels, arch {', '.join(arch)}; opcode counts are static (per SASS listing)")
KEY = ["UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "HMMA", "LDGSTS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMASTG"]
print("# kernel | instructions | " + " ".join(KEY))
tot = collections.Counter()
for f in sorted(funcs, key=lambda f: f.split("\n")[0]):
    name = f.split("\n")[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0].replace("void ", "").replace("kmsr::", "")
    ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f, re.M))
    n = sum(ops.values())
    def cnt(k):
        return sum(v for o, v in ops.items() if o == k or (k in ("FFMA", "LDS", "STS", "LDG", "STG") and o == k))
    print(f"{dem:70s} {n:6d} | " + " ".join(f"{k}={cnt(k)}" for k in KEY if cnt(k)))
    tot.update(ops)
print("# library total: " + " ".join(f"{k}={tot[k]}" for k in KEY if tot[k]))
