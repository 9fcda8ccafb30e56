#!/usr/bin/env python
"""Small invocations of every libkmsr kernel, for `compute-sanitizer --tool memcheck python tools/sanitize_case.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmsr_b200.synth as synth  # noqa: E402
from kmsr_b200 import _lib, ops, rng  # noqa: E402
from kmsr_b200 import A_00_patch_cutter_universal as CUT  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "moe_bank.npz"))
kb, sb = torch.from_numpy(z["kernels"]).cuda(), torch.from_numpy(z["sigmas"])
n = 7
hr = torch.from_numpy(synth.make_hr(n, 1, "textured")).cuda()
pool = torch.from_numpy(synth.make_noise_pool(16, 42)).cuda()
kidx, nidx = rng.draw_multi_kernel_indices(n, 10, 16, 42)
for algo in ("tma", "stream", "tiled"):
    ops.degrade_batch(hr, kb, kidx=kidx, sigma=sb, pool=pool, nidx=nidx, factor=8, algo=algo)
    print(algo, _lib.last_algo())
sums = torch.zeros(11, dtype=torch.float64, device="cuda")
ops.degrade_batch_stats(hr, kb[0], pool=pool, nidx=nidx, factor=8, noise_mode="add", sums=sums)
for k, p, s in ((11, 64, 2), (31, 128, 8), (21, 512, 4), (15, 256, 2), (13, 128, 4)):
    x = torch.from_numpy(synth.make_hr(3, 2, "textured", size=p)).cuda()
    for pad in ("replicate", "zero"):
        ops.degrade_batch(x, torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), factor=s, pad_mode=pad, algo="stream")
    print("stream", k, p, s)
scene = synth.make_scene(3, 640, 896, n_fill=2, n_cloud=2)
total, kept, ij, offs, dev_scene = CUT.create_patches_from_raw(scene)
ops.degrade_batch(dev_scene, kb[2], factor=8, patch_offsets=offs, patch_hw=(256, 256), strides=(640 * 896, 896), scene_hw=(640, 896),
                  x_multiple=128, pool=pool, nidx=rng.draw_noise_indices(kept, 16, 42), noise_mode="add")
print("windows", kept, _lib.last_algo())
work = dev_scene.clone()
m = ops.water_mask(work, 1e-6, 7.0)
ops.keep_mask(m, 256, 128, 0.0)
ops.band_stats(hr)
ops.add_noise_batch(torch.zeros(n, 5, 32, 32, device="cuda"), pool, nidx)
ops.crop_sub(dev_scene, m, np.array([0, 600], dtype=np.int32), np.array([0, 860], dtype=np.int32), 32)
for k, h, w, s in ((11, 64, 64, 2), (31, 72, 200, 2), (13, 100, 36, 4), (21, 256, 256, 2)):
    x = torch.from_numpy(synth.make_hr(3, 2, "textured", size=256)).cuda()[:, :, :h, :w]
    for pad in ("replicate", "zero"):
        ops.degrade_batch(x, torch.from_numpy(synth.softmax_kernels(k, 7)).cuda(), factor=s, pad_mode=pad, algo="reg")
    print("reg", k, h, w, s)
xd = hr[:2, :, :72, :100].clone()
xd[0, 1, 3:9, 4:20] = float("nan")
xd[1, 4] = float("nan")
ops.denoise_nlm(xd, 1.8)
ops.estimate_sigma(xd)
print("denoise")
torch.cuda.synchronize()
print("done")
