"""Drop-in for the arithmetic of kernel_from_lr_gan/D_build_noise_pool.py (noise pool).

`random_crop` keeps the reference signature (numpy in, list of numpy crops out, offsets from the
global CPython `random` stream, top before left, D:49-50).  `build_noise_pool_arrays` is the
D:80-110 loop on in-memory (geophysical_data, denoised) pairs: `noise = geo - den` (D:88) and the
crop are one fused kernel per file, so the full-size noise image is never materialised.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch

from . import ops, patch_io, rng

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # D:23


def random_crop(data: np.ndarray, crop_size: int, n_samples: int) -> list:
    """D:41-53 on the GPU: crops of `data` (C,H,W) at host-drawn offsets (gather kernel, den = 0)."""
    _, h, w = data.shape
    top, left = rng.draw_crop_offsets(h, w, crop_size, n_samples)        # raises ValueError like D:44-45
    if n_samples == 0:
        return []
    ops.require_cuda()
    x = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    out = ops.crop_sub(x, torch.zeros_like(x), top, left, crop_size).cpu().numpy()
    return [out[i] for i in range(n_samples)]


def build_noise_pool_arrays(geo_list, den_list, samples_per_file: int = 1, patch_size: int = 32,
                            seed: int = 42, return_offsets: bool = False):
    """D:65-66 seeds, D:80-92 per-file loop, D:110 stack -> pool (N, C, patch, patch) float32."""
    random.seed(seed)
    np.random.seed(seed)
    ops.require_cuda()
    crops, offsets = [], []
    for geo, den in zip(geo_list, den_list):
        try:
            _, h, w = geo.shape
            top, left = rng.draw_crop_offsets(h, w, patch_size, samples_per_file)
            g = torch.as_tensor(np.ascontiguousarray(geo, dtype=np.float32)).cuda()
            d = torch.as_tensor(np.ascontiguousarray(den, dtype=np.float32)).cuda()
            crops.append(ops.crop_sub(g, d, top, left, patch_size))
            offsets.extend(zip(top.tolist(), left.tolist()))
        except Exception as e:  # noqa: BLE001   per-file skip, D:102-104
            print(f"skipped a file: {e}")
            continue
    if not crops:
        raise RuntimeError("no noise patches extracted")                      # D:106-107
    pool = torch.cat(crops, dim=0).cpu().numpy()
    return (pool, np.array(offsets, dtype=np.int32)) if return_offsets else pool


def load_group_bands(nc_path: str, group_name: str) -> np.ndarray:
    """D:26-38."""
    return patch_io.read_group_bands(nc_path, group_name, BAND_NAMES)


def build_noise_pool(goci_dir: str, output_file: str, metadata_file: str, samples_per_file: int = 1,
                     patch_size: int = 32, seed: int = 42):
    """D:56-132: noise = geophysical_data - denoised of every patch file (os.listdir order, D:71), random 32x32
    crops at CPython-`random` offsets, pool (N,5,32,32) float32 saved with np.save + metadata list."""
    random.seed(seed)
    np.random.seed(seed)
    if not os.path.isdir(goci_dir):
        raise FileNotFoundError(f"GOCI directory does not exist: {goci_dir}")
    names = patch_io.list_patch_files(goci_dir, sort=False)
    if not names:
        raise FileNotFoundError(f"no patch files (.nc / .npz) in {goci_dir}")
    ops.require_cuda()
    crops, metadata = [], []
    print(f"processing {len(names)} files, {samples_per_file} x {patch_size}x{patch_size} noise crops each...")
    for fname in names:
        pth = os.path.join(goci_dir, fname)
        try:
            geo = load_group_bands(pth, "geophysical_data")
            den = load_group_bands(pth, "denoised")
            _, h, w = geo.shape
            top, left = rng.draw_crop_offsets(h, w, patch_size, samples_per_file)     # ValueError before any draw, D:44-45
            crops.append(ops.crop_sub(torch.from_numpy(geo).cuda(), torch.from_numpy(den).cuda(), top, left, patch_size))
            for i in range(samples_per_file):
                metadata.append({"source_file": fname, "patch_id": i, "patch_size": patch_size})
        except Exception as e:  # noqa: BLE001   D:102-104
            print(f"\nfailed {fname}: {e}")
            continue
    if not crops:
        raise RuntimeError("no noise patches extracted")                               # D:106-107
    noise_pool = torch.cat(crops, dim=0).cpu().numpy()
    if os.path.dirname(output_file):
        os.makedirs(os.path.dirname(output_file), exist_ok=True)
    np.save(output_file, noise_pool)
    np.save(metadata_file, metadata)
    print(f"\nnoise pool built: {noise_pool.shape[0]} samples, shape {noise_pool.shape}, saved to {output_file}")
    for i, band in enumerate(BAND_NAMES):
        b = noise_pool[:, i]
        print(f"   {band:12s}: mean={np.nanmean(b):+.6f}, std={np.nanstd(b):.6f}, min={np.nanmin(b):+.6f}, max={np.nanmax(b):+.6f}")
    return noise_pool
