"""Drop-in for kernel_from_lr_gan/data_mean_std.py (per-band radiance statistics).

`analyze_radiance_stats(patch_dir, num_samples=100)` keeps the reference contract: first
`num_samples` of sorted('*.npy') (S:10-18), per patch NaN-skipping mean / population std over
(H,W) (S:32-33), mean over patches (S:45-46), a printed table and the scalar mean of the five
stds (S:60); it returns None.  `radiance_stats` is the additive array-level form and returns the
numbers; with torch.distributed initialised it all-reduces the 2C+1 sums.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from . import ops


def radiance_stats(patches, distributed: bool = False):
    """patches [N,C,H,W] (numpy or tensor) -> dict(mean [N,C], std [N,C], avg_mean [C], avg_std [C], global_avg_std, count)."""
    ops.require_cuda()
    t = patches if isinstance(patches, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(patches, dtype=np.float32))
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = t.to(device=dev, dtype=torch.float32)
    c = x.shape[1]
    sums = torch.zeros(2 * c + 1, dtype=torch.float64, device=dev)
    mean, std = ops.band_stats(x, sums)
    if distributed:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)           # 88 bytes at C=5
    s = sums.cpu().numpy()
    count = s[2 * c]
    avg_mean = s[:c] / count
    avg_std = s[c:2 * c] / count
    return {"mean": mean.cpu().numpy(), "std": std.cpu().numpy(), "avg_mean": avg_mean, "avg_std": avg_std,
            "global_avg_std": float(np.mean(avg_std)), "count": int(count)}


def analyze_radiance_stats(patch_dir, num_samples=100):
    patch_files = sorted(glob.glob(os.path.join(patch_dir, "*.npy")))
    if len(patch_files) == 0:
        print(f"error: no .npy files under {patch_dir}.")
        return
    num_samples = min(num_samples, len(patch_files))
    print(f"analysing {num_samples} patch files...")
    loaded = []
    for f in patch_files[:num_samples]:
        try:
            loaded.append(np.load(f))
        except Exception as e:  # noqa: BLE001   S:37-38
            print(f"skipped {f}: {e}")
    # patches of one shape go to the device as one batch; ragged sets are grouped by shape
    groups: dict = {}
    for i, a in enumerate(loaded):
        groups.setdefault(a.shape, []).append(i)
    nb = loaded[0].shape[0]
    means = np.empty((len(loaded), nb)); stds = np.empty((len(loaded), nb))
    for shape, idxs in groups.items():
        r = radiance_stats(np.stack([loaded[i] for i in idxs]).astype(np.float32))
        means[idxs] = r["mean"]; stds[idxs] = r["std"]
    avg_mean = means.mean(axis=0)
    avg_std = stds.mean(axis=0)
    print("\n" + "=" * 50)
    print(f"radiance statistics ({num_samples} samples):")
    print("=" * 50)
    print(f"{'band':<10} | {'mean radiance (Mean)':<20} | {'mean std (Noise Std)':<20}")
    print("-" * 55)
    for i in range(nb):
        print(f"Band {i:<7} | {avg_mean[i]:<20.6f} | {avg_std[i]:<20.6f}")
    print("-" * 55)
    print(f"all-band mean Std (suggested initial sigma): {np.mean(avg_std):.6f}")
    print("=" * 50)
