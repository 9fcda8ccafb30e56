"""Drop-in for the tiling arithmetic of kernel_from_lr_gan/A_00_patch_cutter_universal.py.

`apply_water_mask` keeps the reference signature and its in-place side effect (-9999 -> NaN on the
argument, CUT:102).  `patch_grid_keep` is the CUT:152-183 raster loop as two kernels (NaN count per
stride cell, then per window) and returns the keep mask; `create_patches` yields the kept windows
in raster (i, j) order -- as zero-copy device views; `create_patches_nc` adds the reference's writer loop
(CUT:126-197) on top of it through patch_io.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops, patch_io

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # CUT:34
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


def create_patches_from_raw(raw, threshold_min: float = THRESHOLD_MIN, threshold_max: float = THRESHOLD_MAX,
                            patch_size: int = PATCH_SIZE, stride_ratio: float = STRIDE_RATIO):
    """process_single_nc's mask + tiling (CUT:89-123, :152-183) fused for the reference's nan_threshold = 0: the keep
    grid comes from one read of the RAW scene and the kept windows -- which contain no masked pixel, hence equal the
    raw pixels -- are returned as offsets into the raw scene itself: (total, kept, ij, offsets, scene_device).
    Nothing is written, the masked copy is never materialised."""
    ops.require_cuda()
    t = raw if isinstance(raw, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw, dtype=np.float32))
    scene = t.cuda().contiguous() if not t.is_cuda else t.contiguous()
    _, h, w = scene.shape
    hp, wp, stride = patch_grid(h, w, patch_size, stride_ratio)
    keep, _ = ops.scene_keep_mask(scene, threshold_min, threshold_max, patch_size, stride, 0.0, NIR_BAND_INDEX, INVALID_VALUE)
    ij = torch.nonzero(keep)
    offsets = (ij[:, 0] * stride * w + ij[:, 1] * stride).to(torch.int64)
    return max(hp, 0) * max(wp, 0), int(ij.shape[0]), ij.cpu().numpy(), offsets, scene


def save_patch_as_nc(patch: np.ndarray, output_path: str, metadata: dict, grid_i: int, grid_j: int, h_offset: int,
                     w_offset: int) -> None:
    """CUT:200-260: one patch file with groups geophysical_data (one variable per band) and navigation_data
    (2-D navigation arrays cropped to the patch window) plus the grid attributes."""
    band_names = metadata.get("band_names", BAND_NAMES)
    _, h, w = patch.shape
    groups = {"geophysical_data": {b: patch[i] for i, b in enumerate(band_names)}}
    nav = metadata.get("navigation_data", {})
    if nav:
        groups["navigation_data"] = {k: v[h_offset:h_offset + h, w_offset:w_offset + w] for k, v in nav.items()
                                     if getattr(v, "ndim", 0) == 2}
    attrs = {"source_file": metadata.get("source_file", "unknown"), "invalid_value": metadata.get("invalid_value", INVALID_VALUE),
             "grid_i": grid_i, "grid_j": grid_j, "h_offset": h_offset, "w_offset": w_offset, "patch_size": h,
             "description": "Patch extracted from Landsat/GOCI-2 L1B data"}
    patch_io.write_groups(output_path, groups, attrs)


def create_patches_nc(data: np.ndarray, patch_size: int, stride_ratio: float, nan_threshold: float, output_dir: str,
                      prefix: str, metadata: dict, ext: str = ".nc"):
    """CUT:126-197: tile `data` [C,H,W] (stride = int(patch_size * stride_ratio)), keep a window iff its NaN ratio is
    <= nan_threshold, write `<prefix>_<i:03d>_<j:03d><ext>` per kept window in raster order; returns (total, kept).
    The NaN census of all windows is one GPU pass (ops.keep_mask); only kept windows are copied out and written."""
    total, kept, ij, _, scene = create_patches(data, patch_size, stride_ratio, nan_threshold)
    _, h, w = scene.shape
    hp, wp, stride = patch_grid(h, w, patch_size, stride_ratio)
    print(f"  data size: {tuple(scene.shape)}")
    print(f"  patch size: {patch_size}x{patch_size}, stride: {stride} ({int(stride_ratio * 100)}% overlap)")
    print(f"  patch grid: {hp}x{wp} = {hp * wp}")
    os.makedirs(output_dir, exist_ok=True)
    host = data if isinstance(data, np.ndarray) else scene.cpu().numpy()
    for i, j in ij:
        hs, ws = int(i) * stride, int(j) * stride
        save_patch_as_nc(host[:, hs:hs + patch_size, ws:ws + patch_size], os.path.join(output_dir, f"{prefix}_{int(i):03d}_{int(j):03d}{ext}"),
                         metadata, int(i), int(j), hs, ws)
    print(f"  generated: {total} windows, kept: {kept} (dropped {total - kept}), saved to {output_dir}")
    return total, kept
