#!/usr/bin/env python
"""One shape through one degrade kernel, for ncu / quick timing:  python tools/degrade_case.py K P S ALGO [GB] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops  # noqa: E402

k, p, s, algo = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
gb = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
n = max(8, int(gb * 1e9 / (4 * 5 * p * p)))
hr = torch.randn((n, 5, p, p), device="cuda") * 3.0 + 50.0
pb = ops.prepare_kernels(torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), s)
out = torch.empty((n, 5, p // s, p // s), device="cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.degrade_batch(hr, pb, factor=s, out=out, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
by = 4 * 5 * (p * p + (p // s) ** 2) * n
fma = 5 * (p // s) ** 2 * (k + s - 1) ** 2 * n
print(f"k={k} P={p} s={s} n={n} algo={_lib.last_algo()} {ms:.3f} ms {by / ms / 1e6:.0f} GB/s {2 * fma / ms / 1e9:.1f} TFLOP/s")
