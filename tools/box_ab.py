#!/usr/bin/env python
"""A/B of box-kernel variants selected by environment variables, in one process:
python tools/box_ab.py "K,P,S;K,P,S;..." "NAME:VAR=VAL,VAR=VAL;NAME2:..." [GB] [reps]
Prints per cell the time of every variant (best of reps) and whether its output equals the first variant's bitwise."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops  # noqa: E402

cells = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1].split(";")]
variants = []
for v in sys.argv[2].split(";"):
    name, _, kv = v.partition(":")
    variants.append((name, dict(x.split("=") for x in kv.split(",") if x)))
gb = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
algo = os.environ.get("AB_ALGO", "box")
allvars = sorted({k for _, d in variants for k in d})
print("cell            " + " ".join(f"{n:>12s}" for n, _ in variants))
for k, p, s in cells:
    n = max(8, int(gb * 1e9 / (4 * 5 * p * p)))
    hr = torch.randn((n, 5, p, p), device="cuda") * 3.0 + 50.0
    pb = ops.prepare_kernels(torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), s)
    ref = None
    row = []
    for name, env in variants:
        for v in allvars:
            os.environ.pop(v, None)
        os.environ.update(env)
        out = torch.empty((n, 5, p // s, p // s), device="cuda")
        best = 1e9
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.degrade_batch(hr, pb, factor=s, out=out, algo=algo)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        same = "" if ref is None else ("=" if torch.equal(out, ref) else "!")
        if ref is None:
            ref = out
        tf = 2 * 5 * (p // s) ** 2 * (k + s - 1) ** 2 * n / best / 1e9
        gbs = 4 * 5 * (p * p + (p // s) ** 2) * n / best / 1e6
        row.append(f"{best:7.3f}{same:1s}{tf:4.0f}")
    print(f"{k:2d},{p:3d},{s} {_lib.last_algo():6s}" + " ".join(f"{r:>12s}" for r in row) + f"   (ms, TFLOP/s; last: {gbs:.0f} GB/s)")
    del hr, ref, out
