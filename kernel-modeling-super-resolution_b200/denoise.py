"""Drop-in for denoise/denoise.py (the upstream NLM denoise stage).

`denoise_band_float_nlm(img_float, h_factor=1.15, patch_size=7, patch_distance=11, verbose=True)` keeps the
reference contract (denoise.py:34-68): one 2-D band in, `(denoised, estimated_sigma)` out, NaN pixels filled with
the band's nanmean for the computation and restored afterwards, an all-NaN band returned as it is with sigma 0.0.
`process_nc_file(file_path, output_dir, h_factor=1.8, plot=False, verbose=True)` (and `batch_denoise`, the loop of
batch_denoise.py) keeps the folder-level contract
(denoise.py:150-262): reads the five `geophysical_data` bands (zeros -> NaN, :31), denoises each, copies the file to
`<stem>_denoised.<ext>` and adds a `denoised` group with the per-band `<band>_sigma` / `<band>_h` attributes and
their averages; returns `(success, output_path, error_msg)` and never raises.  `denoise_bands` is the additive
batched form.  skimage's estimate_sigma / denoise_nl_means arithmetic runs in libkmsr's kernels
(csrc/denoise.cu); there is no CPU fallback.  The plotting helpers of the reference are out of scope.
"""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch

from . import BAND_NAMES, ops, patch_io


def denoise_bands(bands, h_factor: float = 1.15, patch_size: int = 7, patch_distance: int = 11):
    """bands [N,C,H,W] / [C,H,W] (numpy or tensor, CPU or CUDA) -> (denoised, sigma [N,C] / [C] float64), same
    container type and device as the input."""
    ops.require_cuda()
    is_np = not isinstance(bands, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(bands, dtype=np.float32)) if is_np else bands
    squeeze = t.ndim == 3
    if squeeze:
        t = t.unsqueeze(0)
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    out, sigma = ops.denoise_nlm(t.to(device=dev, dtype=torch.float32), h_factor, patch_size, patch_distance)
    if squeeze:
        out, sigma = out[0], sigma[0]
    if is_np:
        return out.cpu().numpy(), sigma.cpu().numpy()
    if not t.is_cuda:
        return out.cpu(), sigma.cpu()
    return out, sigma


def denoise_band_float_nlm(img_float, h_factor=1.15, patch_size=7, patch_distance=11, verbose=True):
    img = np.asarray(img_float)
    valid_mask = ~np.isnan(img)
    if not valid_mask.any():                              # denoise.py:40-41
        return img_float, 0.0
    out, sigma = denoise_bands(img[None, None].astype(np.float32), h_factor, patch_size, patch_distance)
    estimated_sigma = float(sigma[0, 0])
    if verbose:
        print(f"    -> Sigma: {estimated_sigma:.6f} | h: {h_factor * estimated_sigma:.6f}")
    return out[0, 0], estimated_sigma


def process_nc_file(file_path, output_dir, h_factor=1.8, plot=False, verbose=True):
    try:
        file_path = str(file_path)
        if verbose:
            print(f"Loading: {file_path}")
        arr = patch_io.read_group_bands(file_path, "geophysical_data", BAND_NAMES)
        arr = np.where(arr != 0, arr, np.nan).astype(np.float32)          # denoise.py:31
        os.makedirs(str(output_dir), exist_ok=True)
        stem, ext = os.path.splitext(os.path.basename(file_path))
        output_path = os.path.join(str(output_dir), f"{stem}_denoised{ext}")
        denoised, sigma = denoise_bands(arr, h_factor, 7, 11)             # denoise.py:204
        all_nan = np.isnan(arr).all(axis=(1, 2))
        attrs = {"h_factor": h_factor, "denoising_method": "Non-Local Means (NLM)", "patch_size": 7, "patch_distance": 11}
        for i, name in enumerate(BAND_NAMES):
            s = 0.0 if all_nan[i] else float(sigma[i])
            if verbose:
                print(f"\n--- Processing {name} ---\n    -> Sigma: {s:.6f} | h: {h_factor * s:.6f}")
            attrs[f"{name}_sigma"] = s
            attrs[f"{name}_h"] = h_factor * s
        avg_sigma = float(np.mean([attrs[f"{n}_sigma"] for n in BAND_NAMES]))
        attrs["average_sigma"] = avg_sigma
        attrs["average_h"] = h_factor * avg_sigma
        shutil.copy2(file_path, output_path)
        patch_io.add_group(output_path, "denoised", denoised, BAND_NAMES, group_attrs=attrs)
        if verbose:
            print(f"denoised data saved in group 'denoised' of {output_path}")
            print(f"  -> Average Sigma: {avg_sigma:.6f}, Average h: {h_factor * avg_sigma:.6f}")
        return True, output_path, None
    except Exception as e:  # noqa: BLE001   denoise.py:258-262
        error_msg = f"Error: {str(e)}"
        if verbose:
            print(error_msg)
        return False, None, error_msg


def batch_denoise(input_dir, output_dir=None, h_factor=1.8, pattern="*.nc", verbose=False):
    """The loop of denoise/batch_denoise.py:16-93 as a function: every file of `input_dir` matching `pattern` goes
    through process_nc_file into `output_dir` (default `<input_dir>_denoised`, :39-42); failures are collected, not
    raised (:60-93).  Returns (success_count, failed_files) with failed_files = [(file name, error message)].
    `*.npz` group containers are picked up next to `*.nc` when the pattern is the default one."""
    import glob
    input_dir = str(input_dir)
    if not os.path.isdir(input_dir):
        print(f"error: input directory does not exist: {input_dir}")
        return 0, []
    if output_dir is None:
        output_dir = os.path.join(os.path.dirname(os.path.abspath(input_dir)), f"{os.path.basename(os.path.abspath(input_dir))}_denoised")
    os.makedirs(str(output_dir), exist_ok=True)
    files = sorted(glob.glob(os.path.join(input_dir, pattern)))
    if pattern == "*.nc":
        files += sorted(glob.glob(os.path.join(input_dir, "*.npz")))
    if not files:
        print(f"error: no file matching '{pattern}' in {input_dir}")
        return 0, []
    print(f"batch denoise: {len(files)} files, h_factor={h_factor}, patch_size=7, patch_distance=11 -> {output_dir}")
    success_count, failed_files = 0, []
    for i, f in enumerate(files, 1):
        if verbose:
            print(f"\n[{i}/{len(files)}]")
        ok, _, err = process_nc_file(f, output_dir, h_factor=h_factor, plot=False, verbose=verbose)
        if ok:
            success_count += 1
        else:
            failed_files.append((os.path.basename(f), err))
    print(f"done: {success_count}/{len(files)} succeeded, {len(failed_files)} failed")
    return success_count, failed_files
