"""CPU oracle for the LR/HR pair-synthesis hot path  --  TEST INFRASTRUCTURE ONLY.

This module restates, in our own words, what the reference scripts compute on the
hot path.  It exists to *check* the CUDA product path; nothing in the product
package imports it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Where the arithmetic lives: the reference (Python) delegates every flop to
third-party libraries that are neither vendored nor version-pinned by it
(no requirements file anywhere in the tree): PyTorch CPU `F.pad / F.conv2d /
F.avg_pool2d` (C_30:109,112,122; C_31:87,89,95), NumPy legacy `np.random.seed /
randint` (E:190, E:72), `np.nanmean / np.nanstd / np.mean` (data_mean_std:32-33,
45-46) and CPython `random.seed / randint` (D:65, D:49-50).  The oracle therefore
issues *the same library calls at the same call sites* with the container's
torch 2.11.0 / numpy 2.3.5 / CPython 3.12.3 -- that is the reference's CPU path.
A second, library-free restatement in plain C (fp32 in reference operation order
and an fp64 "truth") lives in oracle/oracle.c.

Parity pin: PINNED.  `tests/golden/make_golden.py` imports the real reference
functions from /root/reference in the build container and stores their outputs
in `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
below against those vectors (bit-exact for indices/offsets/masks and for the
float results on the machine that generated them, 2e-6*range otherwise because
ATen's CPU convolution picks ISA-specific kernels).

Citations are `file:line` relative to /root/reference/kernel_from_lr_gan/.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np
import torch
import torch.nn.functional as F

INVALID_VALUE = -9999.0          # A_00_patch_cutter_universal.py:36
NIR_BAND_INDEX = 4               # A_00_patch_cutter_universal.py:35


# --------------------------------------------------------------------------
# a1  load_kernel  (C_30:18-33, C_31:22-37) -- array form (no file, no prints)
# --------------------------------------------------------------------------
def kernel_from_array_c30(arr: np.ndarray) -> torch.Tensor:
    """C_30:26-27: f32 cast, shape untouched."""
    return torch.from_numpy(np.asarray(arr).astype(np.float32))


def kernel_from_array_c31(arr: np.ndarray) -> torch.Tensor:
    """C_31:24-32: f32 cast; 4-D [B,C,k,k] -> mean over B; 2-D -> [1,k,k]."""
    t = torch.from_numpy(np.asarray(arr).astype(np.float32))
    if t.ndim == 4:
        t = t.mean(dim=0)
    if t.ndim == 2:
        t = t.unsqueeze(0)
    return t


# --------------------------------------------------------------------------
# a2  apply_kernel_degradation  (C_30:68-124 == C_31:59-97)
# --------------------------------------------------------------------------
def normalize_kernel(kernel: torch.Tensor, bands: int, strict_ndim: bool = False) -> torch.Tensor:
    """C_30:83-97 / C_31:64-78: band broadcast, band-count assert, per-band sum>0 normalise."""
    if kernel.ndim == 2:
        kernel = kernel.unsqueeze(0).repeat(bands, 1, 1)
    elif kernel.ndim == 3:
        assert kernel.shape[0] == bands, (
            f"kernel bands ({kernel.shape[0]}) != image bands ({bands})")
    elif strict_ndim:
        raise ValueError(f"unsupported kernel ndim: {tuple(kernel.shape)}")   # C_31:68-69
    kn = kernel.clone()
    for c in range(bands):
        s = kernel[c].sum()
        if s > 0:
            kn[c] = kernel[c] / s
    return kn


def apply_kernel_degradation(img: torch.Tensor, kernel: torch.Tensor,
                             downscale_factor: int = 8, strict_ndim: bool = False) -> torch.Tensor:
    """Replicate-pad depthwise cross-correlation, then int(log2(f)) 2x2 mean pools.

    C_30:104-124: pad k//2 both sides (replicate), conv2d(groups=C, padding=0),
    `for _ in range(int(np.log2(f))): avg_pool2d(2,2)`.
    """
    bands = img.shape[0]
    kn = normalize_kernel(kernel, bands, strict_ndim)
    kh, kw = kn.shape[-2:]
    x = F.pad(img.unsqueeze(0), (kw // 2, kw // 2, kh // 2, kh // 2), mode="replicate")
    y = F.conv2d(x, kn.unsqueeze(1), padding=0, groups=bands)
    for _ in range(int(np.log2(downscale_factor))):
        y = F.avg_pool2d(y, kernel_size=2, stride=2)
    return y.squeeze(0)


def degrade_gem(x: torch.Tensor, batch_kernels: torch.Tensor, step: int = 4) -> torch.Tensor:
    """muti_kernel/train_gemini.py:124-134: zero-pad conv with per-sample kernels, `[::step]` decimation.

    x [B,C,H,W], batch_kernels [B,C,k,k] already effective (sum to one); no normalisation here.
    """
    b, c, h, w = x.shape
    k = batch_kernels.shape[-1]
    out = F.conv2d(x.reshape(1, b * c, h, w), batch_kernels.reshape(b * c, 1, k, k),
                   padding=k // 2, groups=b * c)
    return out.view(b, c, h, w)[:, :, ::step, ::step]


# --------------------------------------------------------------------------
# a5/a6  add_noise + pair assembly  (E:65-74, E:187-272)
# --------------------------------------------------------------------------
def add_noise(blurred: np.ndarray, noise_pool: np.ndarray) -> np.ndarray:
    """E:72-74: one draw from the *global* legacy numpy stream, gather, add."""
    idx = np.random.randint(0, len(noise_pool))
    return blurred + noise_pool[idx]


def draw_noise_indices(n_valid: int, pool_len: int, seed: int = 42) -> np.ndarray:
    """Index stream E produces: seed once (E:190), one scalar randint per accepted file (E:250->72)."""
    np.random.seed(seed)
    return np.array([np.random.randint(0, pool_len) for _ in range(n_valid)], dtype=np.int64)


def make_pairs(hr_list, blurred_list, noise_pool: np.ndarray, seed: int = 42):
    """E:190 + loop E:223-250 without NetCDF: shape gates skip *without* drawing (E:239-247)."""
    np.random.seed(seed)
    pairs = []
    for hr, blurred in zip(hr_list, blurred_list):
        if hr.shape[1] != 256 or hr.shape[2] != 256:
            pairs.append(None)
            continue
        if blurred.shape[1] != 32 or blurred.shape[2] != 32:
            pairs.append(None)
            continue
        pairs.append((hr, add_noise(blurred, noise_pool)))
    return pairs


# --------------------------------------------------------------------------
# a3  multi-kernel + sigma composition (build-defined)
# --------------------------------------------------------------------------
def draw_multi_kernel_indices(n: int, n_kernels: int, pool_len: int, seed: int = 42):
    """rs=RandomState(seed); kidx=rs.randint(0,nK,N); nidx=rs.randint(0,Npool,N)."""
    rs = np.random.RandomState(seed)
    kidx = rs.randint(0, n_kernels, n)
    nidx = rs.randint(0, pool_len, n)
    return kidx.astype(np.int32), nidx.astype(np.int32)


def multi_kernel_pairs(hr: np.ndarray, kbank: np.ndarray, sigma, pool, kidx, nidx,
                       factor: int = 8, pad_mode: str = "replicate", down_mode: str = "boxmean",
                       noise_mode: str = "sigma") -> np.ndarray:
    """lr[n,c] = degrade(hr[n], K[kidx[n]])[c] + scale[n,c] * pool[nidx[n],c].

    degrade = C_31:59-97 (replicate + box mean) or train_gemini.py:124-134 (zero + decimate);
    scale = sigma[kidx[n],c] (train_gemini.py:137 semantics), 1 (E:74) or no noise.
    The scaled add is evaluated as one fp32 fused multiply-add (exact product, one
    rounding) which for scale==1 equals the plain add of E:74 bit for bit.
    """
    n = hr.shape[0]
    out = []
    for i in range(n):
        k = torch.from_numpy(np.ascontiguousarray(kbank[kidx[i]] if kidx is not None else kbank[0]))
        x = torch.from_numpy(np.ascontiguousarray(hr[i]))
        if pad_mode == "replicate" and down_mode == "boxmean":
            y = apply_kernel_degradation(x, k, factor).numpy()
        elif pad_mode == "zero" and down_mode == "decimate":
            kn = normalize_kernel(k, x.shape[0])
            y = degrade_gem(x.unsqueeze(0), kn.unsqueeze(0), factor)[0].numpy()
        else:
            raise ValueError("oracle covers the two reference combinations only")
        if noise_mode != "none":
            nz = pool[nidx[i]].astype(np.float64)
            if noise_mode == "sigma":
                sc = sigma[kidx[i] if kidx is not None else 0].astype(np.float64)[:, None, None]
            else:
                sc = 1.0
            y = (y.astype(np.float64) + sc * nz).astype(np.float32)   # == fp32 FMA (single rounding)
        out.append(y)
    return np.stack(out, axis=0)


# --------------------------------------------------------------------------
# a4  noise pool  (D:41-53, D:56-132)
# --------------------------------------------------------------------------
def random_crop(data: np.ndarray, crop_size: int, n_samples: int):
    """D:43-53: top then left from CPython `random.randint` (inclusive bounds)."""
    _, h, w = data.shape
    if h < crop_size or w < crop_size:
        raise ValueError(f"image {h}x{w} smaller than crop {crop_size}")
    patches = []
    for _ in range(n_samples):
        top = _pyrandom.randint(0, h - crop_size)
        left = _pyrandom.randint(0, w - crop_size)
        patches.append(data[:, top:top + crop_size, left:left + crop_size])
    return patches


def draw_crop_offsets(shapes, crop_size: int, samples_per_file: int, seed: int = 42) -> np.ndarray:
    """(top,left) stream of D: seed once (D:65), per file per sample top then left (D:49-50)."""
    _pyrandom.seed(seed)
    offs = []
    for (h, w) in shapes:
        for _ in range(samples_per_file):
            top = _pyrandom.randint(0, h - crop_size)
            left = _pyrandom.randint(0, w - crop_size)
            offs.append((top, left))
    return np.array(offs, dtype=np.int32).reshape(-1, 2)


def build_noise_pool(geo_list, den_list, samples_per_file: int = 1, patch_size: int = 32,
                     seed: int = 42) -> np.ndarray:
    """D:65-66, D:80-92, D:110 on in-memory arrays: noise = geo - den, random crops, stack."""
    _pyrandom.seed(seed)
    np.random.seed(seed)
    crops = []
    for geo, den in zip(geo_list, den_list):
        noise = geo - den
        crops.extend(random_crop(noise, patch_size, samples_per_file))
    if not crops:
        raise RuntimeError("no noise patches extracted")           # D:106-107
    return np.stack(crops, axis=0)


# --------------------------------------------------------------------------
# a7  per-band statistics  (data_mean_std.py:5-62)
# --------------------------------------------------------------------------
def radiance_stats(patches, num_samples: int = 100):
    """S:17-18 (first num_samples), S:32-33 per patch, S:41-46 mean over patches, S:60 scalar.

    Returns (per_patch_mean [N,C], per_patch_std [N,C], avg_mean [C], avg_std [C], global_avg_std).
    """
    sel = patches[:min(num_samples, len(patches))]
    means, stds = [], []
    for data in sel:
        means.append(np.nanmean(data, axis=(1, 2)))
        stds.append(np.nanstd(data, axis=(1, 2)))
    all_means = np.array(means)
    all_stds = np.array(stds)
    avg_mean = np.mean(all_means, axis=0)
    avg_std = np.mean(all_stds, axis=0)
    return all_means, all_stds, avg_mean, avg_std, np.mean(avg_std)


def radiance_stats_f64(patches):
    """fp64 two-pass truth for the same statistics (used for the <=1e-6 relative bar)."""
    x = np.asarray(patches, dtype=np.float64)
    m = np.nanmean(x, axis=(2, 3))
    s = np.nanstd(x, axis=(2, 3))
    return m, s, m.mean(axis=0), s.mean(axis=0)


# --------------------------------------------------------------------------
# a8  scene mask + tiling  (A_00_patch_cutter_universal.py:89-123, 126-197)
# --------------------------------------------------------------------------
def apply_water_mask(data: np.ndarray, threshold_min: float, threshold_max: float) -> np.ndarray:
    """CUT:102-113: -9999 -> NaN *in place*, NIR window test, NaN every band outside it."""
    data[data == INVALID_VALUE] = np.nan
    nir = data[NIR_BAND_INDEX].copy()
    with np.errstate(invalid="ignore"):
        water = (nir >= threshold_min) & (nir <= threshold_max)
    masked = data.copy()
    for c in range(data.shape[0]):
        masked[c][~water] = np.nan
    return masked


def patch_grid(height: int, width: int, patch_size: int = 256, stride_ratio: float = 0.5):
    """CUT:152-155."""
    stride = int(patch_size * stride_ratio)
    return (height - patch_size) // stride + 1, (width - patch_size) // stride + 1, stride


def keep_mask(masked: np.ndarray, patch_size: int = 256, stride_ratio: float = 0.5,
              nan_threshold: float = 0.0) -> np.ndarray:
    """CUT:166-183: raster (i,j) loop; keep iff nan_ratio > nan_threshold is False."""
    _, h, w = masked.shape
    hp, wp, stride = patch_grid(h, w, patch_size, stride_ratio)
    keep = np.zeros((max(hp, 0), max(wp, 0)), dtype=bool)
    for i in range(hp):
        for j in range(wp):
            p = masked[:, i * stride:i * stride + patch_size, j * stride:j * stride + patch_size]
            ratio = np.sum(np.isnan(p)) / p.size
            keep[i, j] = not (ratio > nan_threshold)
    return keep


# --------------------------------------------------------------------------
# tolerance helpers shared by the parity tests
# --------------------------------------------------------------------------
def band_range(hr: np.ndarray) -> np.ndarray:
    """Per-patch per-band dynamic range max-min over (H,W); hr [..., C, H, W] -> [..., C, 1, 1]."""
    return (np.nanmax(hr, axis=(-2, -1)) - np.nanmin(hr, axis=(-2, -1)))[..., None, None]


def rel_err(a: np.ndarray, b: np.ndarray, rng: np.ndarray) -> float:
    """max |a-b| / per-band range."""
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / rng))


# --------------------------------------------------------------------------
# f4  denoise_band_float_nlm  (denoise/denoise.py:34-68) -- PARITY UNPINNED
# --------------------------------------------------------------------------
# skimage.restoration.{estimate_sigma, denoise_nl_means} and PyWavelets are absent from this image, so this
# call-site port cannot issue the reference's library calls; it strings the reference's own numpy lines around
# the restated algorithms of oracle/oracle_nlm.c (see its header for what is restated and why it is unpinned).
def denoise_band_float_nlm(img_float, h_factor=1.15, patch_size=7, patch_distance=11, exact=False, eps=0.0):
    """Returns (denoised, sigma) like denoise.py:34-68; with exact=True the fp64 evaluation and the per-pixel
    cut-off sensitivity `flip` are returned as (denoised64, sigma, flip)."""
    from oracle import oracle_c
    img_float = np.asarray(img_float)
    valid_mask = ~np.isnan(img_float)                                             # :38
    if not valid_mask.any():                                                       # :39-40
        return (img_float, 0.0, None) if exact else (img_float, 0.0)
    fill_value = np.nanmean(img_float)                                             # :42
    img_filled = np.nan_to_num(img_float, nan=fill_value).astype(np.float32)       # :43
    estimated_sigma = oracle_c.estimate_sigma(img_filled)                          # :46
    h_val = h_factor * estimated_sigma                                             # :49
    if exact:
        den, flip = oracle_c.nlm_exact_f64(img_filled, h_val, estimated_sigma, patch_size, patch_distance, eps)
        return np.where(valid_mask, den, np.nan), estimated_sigma, flip
    den = oracle_c.nlm_fast_f32(img_filled, h_val, estimated_sigma, patch_size, patch_distance)   # :55-62
    return np.where(valid_mask, den, np.nan), estimated_sigma                      # :65
