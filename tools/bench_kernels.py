#!/usr/bin/env python
"""Achieved bandwidth of every auxiliary libkmsr kernel against its algorithmic bytes.

    python tools/bench_kernels.py [--out gpurun_out/kernels.json]

CUDA events on the current stream, best of 5 after a warm-up; inputs are larger than L2 unless noted.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmsr_b200 import ops  # noqa: E402

PEAK = 6448.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def best_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def row(name, algo_bytes, ms, note=""):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    return {"kernel": name, "algorithmic_bytes": int(algo_bytes), "ms": ms, "gbs": gbs, "hbm_frac": gbs / PEAK, "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    rows = []

    # band_stats: 4096 patches [5,256,256], one pass
    n = 4096
    x = torch.randn((n, 5, 256, 256), generator=g, device=dev) + 50.0
    sums = torch.zeros(11, dtype=torch.float64, device=dev)
    ms = best_ms(lambda: ops.band_stats(x, sums))
    rows.append(row("band_stats_kernel (+ stats_reduce)", x.numel() * 4, ms, "one pass, pivot-shifted fp64 accumulators: HR read once"))

    # add_noise: 262144 LR patches (5.4 GB blurred + pool gathers + out)
    m = 131072
    b = torch.randn((m, 5, 32, 32), generator=g, device=dev)
    pool = torch.randn((4096, 5, 32, 32), generator=g, device=dev)
    nidx = torch.randint(0, 4096, (m,), generator=g, device=dev, dtype=torch.int32)
    out = torch.empty_like(b)
    ms = best_ms(lambda: ops.add_noise_batch(b, pool, nidx, out=out))
    rows.append(row("add_noise_kernel", 3 * b.numel() * 4, ms, "blurred read + pool gather (84 MB pool, L2 resident) + out write"))
    del b, out

    # crop_sub: one 5 x 4096 x 4096 file, 16384 crops of 32 x 32
    geo = torch.randn((5, 4096, 4096), generator=g, device=dev)
    den = torch.randn((5, 4096, 4096), generator=g, device=dev)
    k = 16384
    top = np.random.RandomState(1).randint(0, 4096 - 32, k).astype(np.int32)
    left = np.random.RandomState(2).randint(0, 4096 - 32, k).astype(np.int32)
    ms = best_ms(lambda: ops.crop_sub(geo, den, top, left, 32))
    rows.append(row("crop_sub_kernel", 3 * k * 5 * 32 * 32 * 4, ms, "2 gathered reads (32-float row segments) + 1 write per pool entry"))
    del geo, den

    # scene: 5 x 8192 x 8192
    scene = torch.randn((5, 8192, 8192), generator=g, device=dev) * 0.5 + 3.0
    scene[:, 1000:1400, 2000:2600] = -9999.0
    masked = torch.empty_like(scene)
    ms = best_ms(lambda: ops.water_mask(scene, 1e-6, 7.0, out=masked))
    rows.append(row("water_mask_vec_kernel", 2 * scene.numel() * 4, ms, "scene read + masked copy write; the in-place NaN write-back (CUT:102) only touches replaced fill pixels"))
    ms = best_ms(lambda: ops.keep_mask(masked, 256, 128, 0.0))
    rows.append(row("keep_mask (cell count + window sum)", scene.numel() * 4, ms, "masked scene read once; includes torch allocations of the wrapper"))

    # kernel preparation: 4096 per-patch kernels (f1: dynamic kernels)
    kb = torch.rand((4096, 5, 13, 13), generator=g, device=dev)
    ms = best_ms(lambda: ops.prepare_kernels(kb, 8))
    rows.append(row("prepare_kernels_kernel", kb.numel() * 4 + 4096 * 5 * 400 * 4, ms, "4096 x 5 kernels 13x13 -> 20x20 composites"))

    for r in rows:
        print(f"{r['kernel']:40s} {r['ms']:8.3f} ms {r['gbs']:8.0f} GB/s  {100 * r['hbm_frac']:5.1f} % of {PEAK:.0f}   {r['note']}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"hbm_peak_gbs": PEAK, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
