"""Drop-in for kernel_from_lr_gan/C_30apply_kernel_to_landsat.py (single-kernel apply).

Same function names, argument order, defaults, return layouts and exception types as the
reference; the arithmetic runs in libkmsr's fused blur + downsample kernel on the GPU.
CPU tensors are copied to the current CUDA device and the result is copied back (the reference
returns CPU tensors, C_30:124); CUDA tensors stay on their device.  No CPU arithmetic fallback.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops, patch_io

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # C_30:49


def load_kernel(kernel_path):
    """C_30:18-33: np.load -> float32 -> tensor, shape untouched ([C,kH,kW] or [kH,kW])."""
    kernel = torch.from_numpy(np.load(kernel_path).astype(np.float32))
    print(f"kernel: {os.path.basename(kernel_path)}")
    print(f"  shape: {kernel.shape}")
    print(f"  sum: {kernel.sum().item():.6f}")
    return kernel


def _band_kernel(kernel: torch.Tensor, bands: int, strict_ndim: bool) -> torch.Tensor:
    """C_30:83-88 / C_31:64-69: 2-D kernels are shared by all bands, 3-D ones must match the band count."""
    if kernel.ndim == 2:
        return kernel.unsqueeze(0).repeat(bands, 1, 1)
    if kernel.ndim == 3:
        assert kernel.shape[0] == bands, f"kernel bands ({kernel.shape[0]}) != image bands ({bands})"
        return kernel
    if strict_ndim:
        raise ValueError(f"unsupported kernel shape: {tuple(kernel.shape)}")       # C_31:68-69
    # C_30 has no ndim check: other ranks run into the per-band loop (C_30:94-97), which indexes
    # kernel[i] for i < C (IndexError when the leading dim is shorter), or into conv2d (RuntimeError)
    if kernel.ndim == 0 or kernel.shape[0] < bands:
        raise IndexError(f"index {kernel.shape[0] if kernel.ndim else 0} is out of bounds for dimension 0 "
                         f"of a kernel of shape {tuple(kernel.shape)}")
    raise RuntimeError(f"kernel of shape {tuple(kernel.shape)} cannot be applied as a per-band blur")


def _apply(img, kernel, downscale_factor, strict_ndim):
    if not isinstance(img, torch.Tensor):
        img = torch.as_tensor(np.asarray(img))
    if not isinstance(kernel, torch.Tensor):
        kernel = torch.as_tensor(np.asarray(kernel))
    if img.ndim != 3:
        raise ValueError(f"expected an image of shape [C,H,W], got {tuple(img.shape)}")   # C_30:80 unpack
    bands = img.shape[0]
    k3 = _band_kernel(kernel, bands, strict_ndim)
    ops.require_cuda()
    dev = img.device if img.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = img.to(device=dev, dtype=torch.float32).unsqueeze(0)
    lr = ops.degrade_batch(x, k3.to(device=dev, dtype=torch.float32), factor=int(downscale_factor))
    lr = lr[0]
    return lr if img.is_cuda else lr.cpu()


def apply_kernel_degradation(img, kernel, downscale_factor: int = 8):
    """C_30:68-124: per-band normalised replicate-padded blur, then int(log2(f)) 2x2 mean pools.

    img [C,H,W], kernel [C,kH,kW] or [kH,kW] -> [C, H/f', W/f'] with f' = 2**int(log2(f)).
    """
    return _apply(img, kernel, downscale_factor, strict_ndim=False)


def degrade_patches(patches, kernel, downscale_factor: int = 8) -> torch.Tensor:
    """Additive batched form: patches [N,C,H,W] -> [N,C,H/f',W/f'] in one launch."""
    if not isinstance(patches, torch.Tensor):
        patches = torch.as_tensor(np.asarray(patches))
    k3 = _band_kernel(kernel if isinstance(kernel, torch.Tensor) else torch.as_tensor(kernel), patches.shape[1], True)
    ops.require_cuda()
    dev = patches.device if patches.is_cuda else torch.device("cuda", torch.cuda.current_device())
    lr = ops.degrade_batch(patches.to(device=dev, dtype=torch.float32), k3.to(dev), factor=int(downscale_factor))
    return lr if patches.is_cuda else lr.cpu()


# ---------------------------------------------------------------------------------------------------------
# folder driver (C_30:127-213)
# ---------------------------------------------------------------------------------------------------------
_CHUNK = 256      # files per launch


def load_landsat_nc(nc_path, group: str = "denoised"):
    """C_30:36-65: the five bands of group 'denoised' as a [C,H,W] float32 tensor (masked -> NaN) + band names."""
    img = torch.from_numpy(patch_io.read_group_bands(nc_path, group, BAND_NAMES))
    print(f"image shape: {img.shape}")
    print(f"value range: [{img.min().item():.2f}, {img.max().item():.2f}]")
    return img, list(BAND_NAMES)


def _degrade_files(paths, kernel, downscale_factor, group, strict_ndim):
    """Read `paths` (per-file failures are reported and skipped), degrade all readable patches of a chunk in one
    launch per image shape, yield (path, img, lr) in file order."""
    for a in range(0, len(paths), _CHUNK):
        loaded = []
        for pth in paths[a:a + _CHUNK]:
            try:
                loaded.append((pth, load_landsat_nc(pth, group)[0]))
            except Exception as e:  # noqa: BLE001   C_30:205-209
                print(f"failed: {os.path.basename(pth)}: {e}")
        by_shape: dict = {}
        for i, (_, img) in enumerate(loaded):
            by_shape.setdefault(tuple(img.shape), []).append(i)
        out = [None] * len(loaded)
        for shape, idxs in by_shape.items():
            try:
                k3 = _band_kernel(kernel, shape[0], strict_ndim)
                ops.require_cuda()
                x = torch.stack([loaded[i][1] for i in idxs]).cuda()
                lr = ops.degrade_batch(x, k3.to(device=x.device, dtype=torch.float32), factor=int(downscale_factor)).cpu()
                for j, i in enumerate(idxs):
                    out[i] = lr[j]
            except Exception as e:  # noqa: BLE001   a bad kernel/image pairing fails every file of that shape
                for i in idxs:
                    print(f"failed: {os.path.basename(loaded[i][0])}: {e}")
        for (pth, img), lr in zip(loaded, out):
            if lr is not None:
                yield pth, img, lr


def process_landsat_folder(landsat_dir, kernel_path, output_dir):
    """C_30:127-213: every patch file of `landsat_dir` (sorted) -> `<name>_blurred.<ext>` in `output_dir`, a copy of
    the source plus group 'blurred' (five f4 bands, 8x downsampled).  The blur + downsample of all files of a chunk
    is one GPU launch; the QA plots of C_30:201-203 are out of scope (no matplotlib on the hot path)."""
    kernel = load_kernel(kernel_path)
    names = patch_io.list_patch_files(landsat_dir, sort=True)
    if len(names) == 0:
        print(f"no patch files (.nc / .npz) found in {landsat_dir}")
        return
    print(f"\nfound {len(names)} Landsat patch files")
    os.makedirs(output_dir, exist_ok=True)
    done = 0
    for pth, img, lr in _degrade_files([os.path.join(landsat_dir, f) for f in names], kernel, 8, "denoised", False):
        try:
            base, ext = os.path.splitext(os.path.basename(pth))
            new_path = os.path.join(output_dir, f"{base}_blurred{ext}")
            patch_io.add_group(new_path, "blurred", lr.numpy(), BAND_NAMES, dims=("y_blurred", "x_blurred"), src=pth,
                               history="Original HR patch with added blurred group (applied blur kernel, 8x downsampled)",
                               long_name="Blurred TOA Radiance at {wl} nm")
            print(f"blur + downsample: {tuple(img.shape)} -> {tuple(lr.shape)}; saved {new_path}")
            done += 1
        except Exception as e:  # noqa: BLE001
            print(f"failed: {os.path.basename(pth)}: {e}")
    print(f"\ndone: {done} of {len(names)} files, results in {output_dir}")
