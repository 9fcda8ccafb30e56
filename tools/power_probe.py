#!/usr/bin/env python
"""How the FP32 roofline denominator depends on launch duration on this B200 (power cap), and what the SM clock does
while a degrade kernel runs back to back.  python tools/power_probe.py"""
import ctypes as C
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib as L, ops  # noqa: E402

import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], [False]


def poll():
    while not stop[0]:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.002)


def clocks_during(fn, seconds):
    samples.clear(); stop[0] = False
    t = threading.Thread(target=poll, daemon=True); t.start()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        fn(); n += 1
        if n % 20 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    stop[0] = True; t.join()
    mid = [s for s in samples if s[0] - t0 > 0.3 * seconds]
    clk = sorted(c for _, c, _ in mid); pw = sorted(p for _, _, p in mid)
    return clk[len(clk) // 2], pw[len(pw) // 2], n


dev = torch.device("cuda", 0)
sink = torch.zeros(512 * 148, device=dev)
fma = C.c_double(0)


def probe(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.lib().kmsr_fp32_probe(C.c_void_p(sink.data_ptr()), iters, C.byref(fma), None))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


probe(1000)
print("FFMA2 probe, TFLOP/s by launch duration (0.2 s idle before each launch):")
for iters in (50, 100, 200, 500, 1000, 2000, 5000, 20000, 100000):
    best = 0
    for _ in range(3):
        time.sleep(0.2)
        ms = probe(iters)
        best = max(best, 2 * fma.value / ms / 1e9)
    print(f"  iters {iters:6d}  {ms:8.3f} ms  {best:6.1f} TFLOP/s")
clk, pw, n = clocks_during(lambda: probe(2000), 2.0)
print(f"probe back to back for 2 s: SM clock median {clk} MHz, power median {pw:.0f} W")

for (k, p, s, algo) in ((13, 256, 2, "box"), (31, 256, 2, "box"), (13, 256, 4, "box"), (13, 256, 8, "auto")):
    n = int(1e9 / (4 * 5 * p * p))
    hr = torch.randn((n, 5, p, p), device=dev) * 3 + 50
    pb = ops.prepare_kernels(torch.from_numpy(synth.softmax_kernels(k, 7)).to(dev), s)
    out = torch.empty((n, 5, p // s, p // s), device=dev)
    fn = lambda: ops.degrade_batch(hr, pb, factor=s, out=out, algo=algo)
    fn(); torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms_cold = e0.elapsed_time(e1)
    clk, pw, reps = clocks_during(fn, 2.0)
    e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms_hot = e0.elapsed_time(e1) / 50
    fl = 2 * 5 * (p // s) ** 2 * (k + s - 1) ** 2 * n
    print(f"({k},{p},{s}) {L.last_algo()}: single launch after idle {ms_cold:.3f} ms = {fl / ms_cold / 1e9:.1f} TFLOP/s; back to back "
          f"{ms_hot:.3f} ms = {fl / ms_hot / 1e9:.1f} TFLOP/s at SM clock {clk} MHz, {pw:.0f} W")
