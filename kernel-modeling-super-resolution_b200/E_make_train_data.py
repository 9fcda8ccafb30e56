"""Drop-in for the arithmetic of kernel_from_lr_gan/E_make_train_data.py (pair assembly).

`add_noise` keeps the reference signature and its side effect on the global numpy stream
(E:72: one np.random.randint per call); `make_pairs` is the batched form of the E:223-250 loop
without NetCDF: shape gates skip a file WITHOUT drawing (E:239-247), accepted files draw one
index each in order, and blur + downsample + noise run as ONE fused launch (C_30 -> E fused;
the intermediate 'blurred' group is lossless f4, so fusing is value-preserving).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops, patch_io, rng

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # E:28

_pool_cache: dict = {}


def _pool_checksum(arr: np.ndarray) -> tuple:
    """Cheap content fingerprint: catches in-place edits of a pool that is still the same object."""
    flat = arr.reshape(-1)
    step = max(1, flat.size // 4096)
    probe = np.ascontiguousarray(flat[::step][:4096])
    return (float(probe.astype(np.float64).sum()), float(flat[-1]) if flat.size else 0.0)


def _device_pool(noise_pool) -> torch.Tensor:
    """Keep the last host pool resident on the device (it is replicated per GPU).  The cache holds a
    reference to the source array and compares by identity (an address can be reused by a different pool once the
    old one is freed) plus a sampled checksum (in-place edits)."""
    if isinstance(noise_pool, torch.Tensor) and noise_pool.is_cuda:
        return noise_pool
    arr = np.asarray(noise_pool)
    dev = torch.cuda.current_device()
    hit = _pool_cache.get("pool")
    chk = _pool_checksum(arr)
    if hit is None or hit[0] is not arr or hit[1] != (arr.shape, str(arr.dtype), dev, chk):
        _pool_cache["pool"] = (arr, (arr.shape, str(arr.dtype), dev, chk),
                               torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda())
    return _pool_cache["pool"][2]


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


def load_group_bands(nc_path: str, group_name: str) -> np.ndarray:
    """E:31-42."""
    return patch_io.read_group_bands(nc_path, group_name, BAND_NAMES)


def load_navigation_data(nc_path: str) -> dict:
    """E:45-58."""
    return patch_io.read_navigation(nc_path)


def save_training_sample(output_path: str, hr: np.ndarray, lr: np.ndarray, nav_data: dict):
    """E:77-117."""
    patch_io.write_training_sample(output_path, hr, lr, nav_data, BAND_NAMES)


def process_files(input_dir: str, noise_pool_path: str, output_dir: str, vis_dir: str = None, max_vis: int = 30,
                  seed: int = 42, chunk: int = 256):
    """E:187-272: for every patch file (os.listdir order) read 'denoised' (HR), 'blurred' and navigation data, gate the
    shapes (a rejected or unreadable file draws nothing, E:239-247), lr = blurred + noise_pool[randint] with the
    global numpy stream seeded once (E:190, E:72), write `<name>_train.<ext>` with groups hr / lr / navigation_data.
    The gather + add of a chunk of files is one GPU launch; `vis_dir` / `max_vis` are accepted for signature
    compatibility (plots are out of scope).  Returns (success_count, fail_count)."""
    np.random.seed(seed)
    if not os.path.isdir(input_dir):
        raise FileNotFoundError(f"input directory does not exist: {input_dir}")
    if not os.path.isfile(noise_pool_path):
        raise FileNotFoundError(f"noise pool file does not exist: {noise_pool_path}")
    noise_pool = np.load(noise_pool_path)
    print(f"noise pool {noise_pool_path}: {noise_pool.shape}")
    os.makedirs(output_dir, exist_ok=True)
    names = patch_io.list_patch_files(input_dir, sort=False)
    if not names:
        raise FileNotFoundError(f"no patch files (.nc / .npz) in {input_dir}")
    ops.require_cuda()
    pool_dev = _device_pool(noise_pool)
    ok = fail = 0
    for a in range(0, len(names), chunk):
        accepted = []
        for fname in names[a:a + chunk]:
            pth = os.path.join(input_dir, fname)
            try:
                hr = load_group_bands(pth, "denoised")
                blurred = load_group_bands(pth, "blurred")
                nav = load_navigation_data(pth)
                if hr.shape[1] != 256 or hr.shape[2] != 256:
                    print(f"\n{fname}: HR shape {hr.shape} is not (5,256,256), skipped")
                    fail += 1
                    continue
                if blurred.shape[1] != 32 or blurred.shape[2] != 32:
                    print(f"\n{fname}: blurred shape {blurred.shape} is not (5,32,32), skipped")
                    fail += 1
                    continue
                accepted.append((fname, hr, blurred, nav))
            except Exception as e:  # noqa: BLE001   E:264-267
                print(f"\nfailed {fname}: {e}")
                fail += 1
        if not accepted:
            continue
        nidx = rng.draw_noise_indices(len(accepted), len(noise_pool), None)     # continues the stream seeded above
        b = torch.from_numpy(np.stack([x[2] for x in accepted])).cuda()
        lr = ops.add_noise_batch(b, pool_dev, nidx).cpu().numpy()
        for (fname, hr, _, nav), l in zip(accepted, lr):
            ext = os.path.splitext(fname)[1]
            base = fname.replace(f"_denoised_blurred{ext}", f"_train{ext}")
            if base == fname:
                base = fname.replace(ext, f"_train{ext}")
            try:
                save_training_sample(os.path.join(output_dir, base), hr, l, nav)
                ok += 1
            except Exception as e:  # noqa: BLE001
                print(f"\nfailed {fname}: {e}")
                fail += 1
    print(f"\ndone: {ok} written, {fail} failed, output {output_dir}")
    return ok, fail
