"""Drop-in for the tiling arithmetic of kernel_from_lr_gan/A_00_patch_cutter_universal.py.

`apply_water_mask` keeps the reference signature and its in-place side effect (-9999 -> NaN on the
argument, CUT:102).  `patch_grid_keep` is the CUT:152-183 raster loop as two kernels (NaN count per
stride cell, then per window) and returns the keep mask; `create_patches` yields the kept windows
in raster (i, j) order -- as zero-copy device views -- without writing NetCDF (writer: SURVEY 8f).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

PATCH_SIZE = 256            # CUT:29
STRIDE_RATIO = 0.5          # CUT:30
NAN_THRESHOLD = 0.0         # CUT:31
THRESHOLD_MIN = 0.000001    # CUT:32
THRESHOLD_MAX = 7.0         # CUT:33
NIR_BAND_INDEX = 4          # CUT:35
INVALID_VALUE = -9999.0     # CUT:36


def apply_water_mask(data: np.ndarray, threshold_min: float, threshold_max: float) -> np.ndarray:
    """CUT:89-123.  numpy [C,H,W] in (mutated: invalid -> NaN), masked numpy copy out."""
    ops.require_cuda()
    d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda()
    masked = ops.water_mask(d, float(threshold_min), float(threshold_max), NIR_BAND_INDEX, INVALID_VALUE)
    data[...] = d.cpu().numpy()                         # the reference mutates its argument
    out = masked.cpu().numpy()
    nir = data[NIR_BAND_INDEX]
    total_valid = int(np.sum(~np.isnan(nir)))
    water = int(np.sum(~np.isnan(out[NIR_BAND_INDEX])))
    ratio = water / total_valid * 100 if total_valid > 0 else 0
    print(f"  valid pixels: {total_valid:,}")
    print(f"  water pixels: {water:,} ({ratio:.2f}%)")
    return out


def patch_grid(height: int, width: int, patch_size: int = PATCH_SIZE, stride_ratio: float = STRIDE_RATIO):
    """CUT:152-155."""
    stride = int(patch_size * stride_ratio)
    return (height - patch_size) // stride + 1, (width - patch_size) // stride + 1, stride


def patch_grid_keep(masked: torch.Tensor, patch_size: int = PATCH_SIZE, stride_ratio: float = STRIDE_RATIO,
                    nan_threshold: float = NAN_THRESHOLD):
    """keep [hp, wp] bool (device) and the NaN count per window for a masked scene [C,H,W] on the device."""
    stride = int(patch_size * stride_ratio)
    return ops.keep_mask(masked, patch_size, stride, nan_threshold)


def create_patches(data, patch_size: int = PATCH_SIZE, stride_ratio: float = STRIDE_RATIO,
                   nan_threshold: float = NAN_THRESHOLD):
    """CUT:126-197 without the writer: returns (total, kept, ij [kept,2], offsets int64 [kept], scene_device).

    `offsets` are element offsets of each kept window's top-left pixel in the scene tensor; pass them
    to ops.degrade_batch(patch_offsets=...) to degrade the windows in place, clamped to the window
    (patches are cut first, then blurred: neighbouring scene pixels never leak into a patch's halo).
    """
    ops.require_cuda()
    t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
    scene = t.cuda().contiguous() if not t.is_cuda else t.contiguous()
    _, h, w = scene.shape
    hp, wp, stride = patch_grid(h, w, patch_size, stride_ratio)
    total = max(hp, 0) * max(wp, 0)
    keep, _ = ops.keep_mask(scene, patch_size, stride, nan_threshold)
    ij = torch.nonzero(keep)                                  # raster order == reference loop order
    offsets = (ij[:, 0] * stride * w + ij[:, 1] * stride).to(torch.int64)
    return total, int(ij.shape[0]), ij.cpu().numpy(), offsets, scene
