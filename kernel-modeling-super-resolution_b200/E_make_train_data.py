"""Drop-in for the arithmetic of kernel_from_lr_gan/E_make_train_data.py (pair assembly).

`add_noise` keeps the reference signature and its side effect on the global numpy stream
(E:72: one np.random.randint per call); `make_pairs` is the batched form of the E:223-250 loop
without NetCDF: shape gates skip a file WITHOUT drawing (E:239-247), accepted files draw one
index each in order, and blur + downsample + noise run as ONE fused launch (C_30 -> E fused;
the intermediate 'blurred' group is lossless f4, so fusing is value-preserving).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, rng

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # E:28

_pool_cache: dict = {}


def _device_pool(noise_pool) -> torch.Tensor:
    """Keep the last host pool resident on the device (it is replicated per GPU)."""
    if isinstance(noise_pool, torch.Tensor) and noise_pool.is_cuda:
        return noise_pool
    arr = np.asarray(noise_pool)
    key = (arr.__array_interface__["data"][0], arr.shape, str(arr.dtype), torch.cuda.current_device())
    hit = _pool_cache.get("pool")
    if hit is None or hit[0] != key:
        _pool_cache["pool"] = (key, torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda())
    return _pool_cache["pool"][1]


def add_noise(blurred: np.ndarray, noise_pool: np.ndarray) -> np.ndarray:
    """E:65-74: idx = np.random.randint(0, len(pool)); return blurred + pool[idx]  (gather + add on the GPU)."""
    idx = np.random.randint(0, len(noise_pool))
    ops.require_cuda()
    pool = _device_pool(noise_pool)
    b = torch.from_numpy(np.ascontiguousarray(blurred, dtype=np.float32)).cuda().unsqueeze(0)
    out = ops.add_noise_batch(b, pool, np.array([idx], dtype=np.int32))
    return out[0].cpu().numpy()


def make_pairs(hr, kernel, noise_pool, seed: int | None = 42, downscale_factor: int = 8,
               hr_size: int = 256, lr_size: int = 32, with_stats: bool = False, sums=None):
    """Batched E.process_files arithmetic: hr [N,5,256,256] -> (hr, lr [N,5,32,32], nidx).

    Equivalent to C_30.apply_kernel_degradation on every patch followed by E.add_noise in file
    order with `np.random.seed(seed)` first (E:190); `seed=None` continues the global stream.
    """
    ops.require_cuda()
    t = hr if isinstance(hr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(hr, dtype=np.float32))
    if t.shape[-2] != hr_size or t.shape[-1] != hr_size:                       # E:239-242
        raise ValueError(f"HR patches must be {hr_size}x{hr_size}, got {tuple(t.shape[-2:])}")
    n = t.shape[0]
    nidx = rng.draw_noise_indices(n, len(noise_pool), seed)
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    pool = _device_pool(noise_pool)
    k = kernel if isinstance(kernel, torch.Tensor) else torch.as_tensor(np.asarray(kernel, dtype=np.float32))
    if k.ndim == 2:
        k = k.unsqueeze(0).repeat(t.shape[1], 1, 1)
    if with_stats:
        # data_mean_std.py:32-33 over the HR patches in the same pass: returns (hr, lr, nidx, mean, std)
        lr, mean, std = ops.degrade_batch_stats(t.to(device=dev, dtype=torch.float32), k.to(dev), pool=pool, nidx=nidx,
                                                factor=int(downscale_factor), noise_mode="add", sums=sums)
    else:
        lr = ops.degrade_batch(t.to(device=dev, dtype=torch.float32), k.to(dev), pool=pool, nidx=nidx,
                               factor=int(downscale_factor), noise_mode="add")
    if lr.shape[-1] != lr_size or lr.shape[-2] != lr_size:                     # E:244-247
        raise ValueError(f"LR patches must be {lr_size}x{lr_size}, got {tuple(lr.shape[-2:])}")
    if with_stats:
        return t, (lr if t.is_cuda else lr.cpu()), nidx, mean, std
    return t, (lr if t.is_cuda else lr.cpu()), nidx
