#!/usr/bin/env python
"""Several shapes through the generic streaming kernel, one line each (A/B runs of kernel variants):
python tools/stream_sweep.py "11,13,15,21,31" "2,4" "64,256" [reps] [stream|reg|tiled|auto]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops  # noqa: E402

ks, ss, ps = ([int(v) for v in a.split(",")] for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
algo = sys.argv[5] if len(sys.argv) > 5 else "stream"
torch.manual_seed(1234)
for p in ps:
    n = max(8, int(2e9 / (4 * 5 * p * p)))
    hr = torch.randn((n, 5, p, p), device="cuda") * 3.0 + 50.0
    for s in ss:
        out = torch.empty((n, 5, p // s, p // s), device="cuda")
        for k in ks:
            pb = ops.prepare_kernels(torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), s)
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.degrade_batch(hr, pb, factor=s, out=out, algo=algo)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"k={k} P={p} s={s} n={n} {_lib.last_algo()} {best:.3f} ms checksum {float(out.double().sum()):.6e}", flush=True)
    del hr
