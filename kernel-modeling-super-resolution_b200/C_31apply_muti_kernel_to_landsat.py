"""Drop-in for kernel_from_lr_gan/C_31apply_muti_kernel_to_landsat.py (multi-kernel apply).

C_31's arithmetic is bit-identical to C_30's; what differs is the kernel loader (4-D batch kernels
are averaged over the batch, 2-D ones become [1,kH,kW]: C_31:22-37) and the ValueError for other
kernel ranks (C_31:68-69).  `degrade_multi_kernel` is the additive batched entry point for the
"random pick of the 10 moe_kernels + per-band sigma noise" composition (BASELINE config 2; kernel
bank and sigma semantics from muti_kernel/train_gemini.py:107-115, :137).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops, patch_io, rng
from .C_30apply_kernel_to_landsat import BAND_NAMES, _apply, _degrade_files  # noqa: F401


def load_kernel(kernel_path: str) -> torch.Tensor:
    """C_31:22-37: float32; [B,C,kH,kW] -> mean over B; [kH,kW] -> [1,kH,kW]."""
    kernel = torch.from_numpy(np.load(kernel_path).astype(np.float32))
    if kernel.ndim == 4:
        kernel = kernel.mean(dim=0)
    if kernel.ndim == 2:
        kernel = kernel.unsqueeze(0)
    print(f"kernel: {os.path.basename(kernel_path)}")
    print(f"  shape: {kernel.shape}")
    print(f"  sum: {kernel.sum().item():.6f}")
    return kernel


def apply_kernel_degradation(img: torch.Tensor, kernel: torch.Tensor, downscale_factor: int = 8) -> torch.Tensor:
    """C_31:59-97 (== C_30:68-124 plus ValueError on kernels that are neither 2-D nor 3-D)."""
    return _apply(img, kernel, downscale_factor, strict_ndim=True)


def load_kernel_bank(kernel_dir: str, n_kernels: int = 10):
    """moe_kernels/kernel_{i}.npy [5,13,13] + sigma_{i}.npy [5] (written by train_gemini.py:241-249)."""
    ks = [np.load(os.path.join(kernel_dir, f"kernel_{i}.npy")).astype(np.float32) for i in range(n_kernels)]
    ss = [np.load(os.path.join(kernel_dir, f"sigma_{i}.npy")).astype(np.float32) for i in range(n_kernels)]
    return torch.from_numpy(np.stack(ks)), torch.from_numpy(np.stack(ss))


def degrade_multi_kernel(patches: torch.Tensor, kernel_bank: torch.Tensor, sigma_bank, noise_pool,
                         seed: int = 42, downscale_factor: int = 8, kidx=None, nidx=None,
                         pad_mode: str = "replicate", down_mode: str = "boxmean",
                         noise_mode: str = "sigma"):
    """lr[n,c] = degrade(hr[n], K[kidx[n]])[c] + sigma[kidx[n],c] * pool[nidx[n],c]; returns (lr, kidx, nidx).

    Indices come from one RandomState(seed): kernel picks first, then noise picks (rng.py).
    """
    n = patches.shape[0]
    if kidx is None or nidx is None:
        kidx, nidx = rng.draw_multi_kernel_indices(n, kernel_bank.shape[0], len(noise_pool), seed)
    ops.require_cuda()
    dev = patches.device if patches.is_cuda else torch.device("cuda", torch.cuda.current_device())
    lr = ops.degrade_batch(patches.to(device=dev, dtype=torch.float32), kernel_bank.to(dev), kidx=kidx,
                           sigma=sigma_bank if noise_mode == "sigma" else None,
                           pool=torch.as_tensor(noise_pool).to(dev), nidx=nidx, factor=int(downscale_factor),
                           pad_mode=pad_mode, down_mode=down_mode, noise_mode=noise_mode)
    return (lr if patches.is_cuda else lr.cpu()), kidx, nidx


def load_landsat_nc(nc_path: str):
    """C_31:40-56: the five bands of group 'hr' as a [C,H,W] float32 tensor + band names."""
    return torch.from_numpy(patch_io.read_group_bands(nc_path, "hr", BAND_NAMES)), list(BAND_NAMES)


def process_landsat_folder(landsat_dir: str, kernel_path: str, output_dir: str, downscale_factor: int = 8,
                           visualize_top_n: int = 5) -> None:
    """C_31:136-195: read group 'hr' of every patch file (sorted), write / overwrite group 'lr' IN THE SAME FILE.
    `visualize_top_n` is accepted for signature compatibility; the QA plots are out of scope."""
    kernel = load_kernel(kernel_path)
    names = patch_io.list_patch_files(landsat_dir, sort=True)
    if len(names) == 0:
        print(f"no patch files (.nc / .npz) found in {landsat_dir}")
        return
    print(f"\nfound {len(names)} Landsat patch files")
    os.makedirs(output_dir, exist_ok=True)
    done = 0
    paths = [os.path.abspath(os.path.join(landsat_dir, f)) for f in names]
    for pth, img, lr in _degrade_files(paths, kernel, downscale_factor, "hr", True):
        try:
            patch_io.add_group(pth, "lr", lr.numpy(), BAND_NAMES, dims=("y_lr", "x_lr"),
                               history="Added lr group by applying learned blur kernel and downsampling",
                               long_name="TOA Radiance at {wl} nm (LR)")
            print(f"degraded: {tuple(img.shape)} -> {tuple(lr.shape)}; updated {pth}")
            done += 1
        except Exception as e:  # noqa: BLE001   C_31:185-189
            print(f"failed: {os.path.basename(pth)}: {e}")
    print(f"\ndone: {done} of {len(names)} files; kernel {kernel_path}; source {landsat_dir}")
