#!/usr/bin/env python
"""One shape through the generic streaming kernel, for ncu (python tools/stream_case.py K P S [reps])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops  # noqa: E402

k, p, s = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n = max(8, int(4e9 / (4 * 5 * p * p)))
hr = torch.randn((n, 5, p, p), device="cuda") * 3.0 + 50.0
pb = ops.prepare_kernels(torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), s)
out = torch.empty((n, 5, p // s, p // s), device="cuda")
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.degrade_batch(hr, pb, factor=s, out=out, algo="stream")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
by = 4 * 5 * (p * p + (p // s) ** 2) * n
print(f"k={k} P={p} s={s} n={n} algo={_lib.last_algo()} {ms:.3f} ms {by / ms / 1e6:.0f} GB/s")
