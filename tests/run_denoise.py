"""Measurement of the denoise stage on one B200: GPU time of estimate_sigma + nlm_kernel over
a batch of [5,256,256] patches, parity summary against the checker, and the checker's float32 integral-image
evaluation timed on one host core as the CPU baseline (skimage itself is not in the image).  Lives under tests/
because it executes the oracle.  Usage: python tests/run_denoise.py [--patches 64] [--out gpurun_out/denoise.json]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--h-factor", type=float, default=1.0)      # Landsat (README.MD:18); GOCI-2 uses 1.8
    ap.add_argument("--cpu-bands", type=int, default=2)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from kmsr_b200 import _lib, ops, synth
    from oracle import oracle_c
    from test_gpu_denoise import _check_band

    n, p = a.patches, a.size
    hr = synth.make_hr(n, 4321, "textured")[:, :, :p, :p].copy()
    hr[n // 2:] = synth.make_hr(n - n // 2, 4322, "water")[:, :, :p, :p]
    x = torch.from_numpy(hr).cuda()
    out = torch.empty_like(x)
    for _ in range(2):
        ops.denoise_nlm(x, a.h_factor, out=out)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        _, sigma = ops.denoise_nlm(x, a.h_factor, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    launches = (_lib.launch_count() - l0) // a.steps
    # sigma alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.estimate_sigma(x)
    e0.record(); ops.estimate_sigma(x); e1.record(); torch.cuda.synchronize()
    ms_sigma = e0.elapsed_time(e1)
    best = min(ms)
    px_shifts = n * 5 * p * p * 529.0
    res = {"workload": f"denoise_band_float_nlm of {n} patches [5,{p},{p}] (patch 7, distance 11, h_factor {a.h_factor})",
           "gpu_ms": best, "gpu_ms_all": ms, "gpu_ms_estimate_sigma": ms_sigma, "launches_per_step": launches,
           "patches_per_s": n / best * 1e3, "bands_per_s": 5 * n / best * 1e3,
           "pixel_shifts_per_s": px_shifts / best * 1e3,
           "hbm_gbs_algorithmic": n * 5 * p * p * 8 / best / 1e6}
    # parity on a textured and a water band
    o = out.cpu().numpy()
    sg = sigma.cpu().numpy()
    par = {}
    for name, (i, c) in {"textured": (0, 1), "water": (n - 1, 3)}.items():
        e1_, e2_, s = _check_band(o[i, c], hr[i, c], a.h_factor, name=name, sigma=float(sg[i, c]))
        par[name] = {"ours_vs_exact": e1_, "ours_vs_ref": e2_, "sigma": s, "sigma_gpu": float(sg[i, c])}
    res["parity"] = par
    # CPU baseline: the float32 integral-image algorithm (what skimage's fast mode does), one core
    t0 = time.perf_counter()
    for b in range(a.cpu_bands):
        img = hr[0, b]
        s = oracle_c.estimate_sigma(img)
        oracle_c.nlm_fast_f32(img, a.h_factor * s, s)
    dt = (time.perf_counter() - t0) / a.cpu_bands
    res["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": f"{a.cpu_bands} bands of {p}x{p}",
                           "bands_per_s": 1.0 / dt, "patches_per_s": 1.0 / (5 * dt)}
    res["speedup_vs_one_core"] = res["bands_per_s"] / res["cpu_baseline"]["bands_per_s"]
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
