"""Batched device entry points: thin tensor-level wrappers over the C ABI (include/kmsr.h).

PyTorch supplies device memory and the current stream; every flop happens in libkmsr's CUDA
kernels.  All tensors are float32 CHW (reference layout, C_30apply_kernel_to_landsat.py:59-60);
indices are int32 and are drawn on the host (rng.py) so they match the reference bit for bit.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("kmsr_b200 computes on a CUDA device (B200, sm_100a); none is visible and "
                           "there is no CPU fallback")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32)


def _i32(t, device) -> torch.Tensor | None:
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return t.to(device=device, dtype=torch.int32).contiguous()


def _check_indices(what: str, idx, n: int, upper: int, device, validate_device: bool = False) -> None:
    """kidx < nK, nidx < nPool, crop offsets: a bad index is an out-of-bounds device read in the kernels, so host arrays
    are range-checked here for free (IndexError, as noise_pool[idx] raises in the reference, E:72-74) and device tensors
    through kmsr_validate_indices when the caller opts in (`validate=True`; it synchronises the stream)."""
    if idx is None:
        return
    count = int(idx.numel()) if isinstance(idx, torch.Tensor) else int(np.asarray(idx).size)
    if count != n:
        raise ValueError(f"{what}: {count} indices for {n} patches")
    if n == 0:
        return
    if isinstance(idx, torch.Tensor) and idx.is_cuda:
        if validate_device:
            t = idx.to(torch.int32).contiguous()
            scratch = torch.zeros(1, dtype=torch.int32, device=t.device)
            with torch.cuda.device(t.device):
                rc = L.lib().kmsr_validate_indices(_ptr(t), n, int(upper), _ptr(scratch), what.encode(), _stream(t.device))
            if rc == L.E_INVALID:
                raise IndexError(L.last_error())
            L.check(rc)
        return
    a = idx.numpy() if isinstance(idx, torch.Tensor) else np.asarray(idx)
    lo, hi = int(a.min()), int(a.max())
    if lo < 0 or hi >= upper:
        raise IndexError(f"{what}: values span [{lo}, {hi}], valid range is [0, {upper})")


@dataclass
class PreparedBank:
    """Normalised + box-folded kernel bank on the device (kmsr_prepare_kernels)."""
    comp: torch.Tensor      # [nK, C, KH, KWp]
    dsum: torch.Tensor      # [nK, C]
    nK: int
    C: int
    kh: int
    kw: int
    factor: int
    down_mode: int


def prepare_kernels(kbank: torch.Tensor, factor: int = 8, down_mode: str | int = "boxmean") -> PreparedBank:
    """kbank [nK, C, kh, kw] (or [C, kh, kw]) -> PreparedBank.  C_30:93-97 normalisation + box fold."""
    require_cuda()
    dm = L.DOWN_MODES[down_mode] if isinstance(down_mode, str) else int(down_mode)
    if kbank.ndim == 3:
        kbank = kbank.unsqueeze(0)
    if kbank.ndim != 4:
        raise ValueError(f"kernel bank must be [nK,C,kh,kw] or [C,kh,kw], got {tuple(kbank.shape)}")
    dev = kbank.device if kbank.is_cuda else torch.device("cuda", torch.cuda.current_device())
    kb = _f32(kbank, dev).contiguous()
    nK, Cb, kh, kw = kb.shape
    KH, KWp, _ = L.composite_size(kh, kw, factor, dm)
    comp = torch.empty((nK, Cb, KH, KWp), dtype=torch.float32, device=dev)
    dsum = torch.empty((nK, Cb), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_prepare_kernels(_ptr(kb), nK, Cb, kh, kw, factor, dm, _ptr(comp), _ptr(dsum),
                                             _stream(dev)))
    return PreparedBank(comp, dsum, nK, Cb, kh, kw, int(factor), dm)


def degrade_batch(hr: torch.Tensor, kbank, *, kidx=None, sigma=None, pool=None, nidx=None,
                  factor: int = 8, pad_mode: str = "replicate", down_mode: str = "boxmean",
                  noise_mode: str | None = None, out: torch.Tensor | None = None, algo: str = "auto",
                  patch_offsets: torch.Tensor | None = None, patch_hw: tuple[int, int] | None = None,
                  strides: tuple[int, int, int] | None = None, n_patches: int | None = None,
                  scene_hw: tuple[int, int] | None = None, x_multiple: int = 1, validate: bool = False) -> torch.Tensor:
    """lr[n,c] = degrade(hr[n], K[kidx[n]])[c] (+ scale * pool[nidx[n], c]) on the device.

    hr: CUDA float32 [N, C, H, W] with contiguous rows (any N / C / row strides), or -- with
    `patch_offsets` (int64 element offsets), `patch_hw`, `strides=(sC, sH)` -- a base tensor that the
    patches are windows of (scene-scale path, A_00_patch_cutter_universal.py:176).  With `scene_hw` (extents
    of that base tensor) and `x_multiple` (every window's left column is a multiple of it; >= 4 for the
    TMA kernel) the windows stream straight from the scene through kmsr_degrade_windows.
    kbank: tensor [nK,C,kh,kw] / [C,kh,kw] or a PreparedBank.
    """
    require_cuda()
    if not hr.is_cuda or hr.dtype != torch.float32:
        raise TypeError("degrade_batch expects a CUDA float32 tensor (drop-in wrappers copy host inputs)")
    dev = hr.device
    dm = L.DOWN_MODES[down_mode]
    pm = L.PAD_MODES[pad_mode]
    if noise_mode is None:
        noise_mode = "none" if nidx is None else ("sigma" if sigma is not None else "add")
    nm = L.NOISE_MODES[noise_mode]
    bank = kbank if isinstance(kbank, PreparedBank) else prepare_kernels(kbank.to(dev), factor, dm)
    if bank.factor != factor or bank.down_mode != dm:
        raise ValueError("PreparedBank was built for a different factor / down_mode")

    if patch_offsets is None:
        if hr.ndim != 4:
            raise ValueError(f"hr must be [N,C,H,W], got {tuple(hr.shape)}")
        N, Cc, H, W = hr.shape
        if hr.stride(3) != 1 and W > 1:
            hr = hr.contiguous()
        sN, sC, sH = hr.stride(0), hr.stride(1), hr.stride(2)
        po = None
    else:
        Cc = bank.C
        H, W = patch_hw
        sC, sH = strides
        sN = 0
        po = patch_offsets.to(device=dev, dtype=torch.int64).contiguous()
        N = int(po.numel()) if n_patches is None else int(n_patches)
    assert Cc == bank.C, f"kernel bands ({bank.C}) != image bands ({Cc})"      # C_30:88
    Ho, Wo = L.degrade_out_size(H, W, bank.kh, bank.kw, factor, dm)
    if out is None:
        out = torch.empty((N, Cc, Ho, Wo), dtype=torch.float32, device=dev)
    else:
        assert out.is_cuda and out.is_contiguous() and tuple(out.shape) == (N, Cc, Ho, Wo)
    _check_indices("kidx", kidx, N, bank.nK, dev, validate)
    kidx_d = _i32(kidx, dev)
    nidx_d = _i32(nidx, dev)
    sig, pl, npool = _noise_args(nm, sigma, pool, nidx, N, bank, (Cc, Ho, Wo), dev, validate)
    if po is not None and scene_hw is not None:
        with torch.cuda.device(dev):
            L.check(L.lib().kmsr_degrade_windows(
                _ptr(hr), Cc, int(scene_hw[0]), int(scene_hw[1]), sC, sH, _ptr(po), N, H, W, int(x_multiple),
                _ptr(bank.comp), _ptr(bank.dsum), bank.nK, bank.kh, bank.kw, _ptr(kidx_d),
                _ptr(sig), _ptr(pl), npool, _ptr(nidx_d),
                factor, pm, dm, nm, _ptr(out), L.ALGOS[algo], _stream(dev)))
        return out
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_degrade_prepared(
            _ptr(hr), N, Cc, H, W, sN, sC, sH, _ptr(po),
            _ptr(bank.comp), _ptr(bank.dsum), bank.nK, bank.kh, bank.kw, _ptr(kidx_d),
            _ptr(sig), _ptr(pl), npool, _ptr(nidx_d),
            factor, pm, dm, nm, _ptr(out), L.ALGOS[algo], _stream(dev)))
    return out


def _noise_args(nm, sigma, pool, nidx, N, bank, lr_shape, dev, validate):
    """sigma [nK,C], pool [nPool,C,Ho,Wo] and the nidx range, validated once for degrade_batch and degrade_batch_stats."""
    sig = None if sigma is None else _f32(torch.as_tensor(sigma), dev).contiguous()
    if nm == L.NOISE_NONE:
        return sig, None, 0
    if pool is None or nidx is None:
        raise ValueError("noise requested without pool / nidx")
    pl = pool if (pool.is_cuda and pool.dtype == torch.float32 and pool.is_contiguous()) else _f32(pool, dev).contiguous()
    if pl.ndim != 4 or tuple(pl.shape[1:]) != tuple(lr_shape):
        raise ValueError(f"noise pool {tuple(pl.shape)} does not match LR patches {tuple(lr_shape)}")
    if nm == L.NOISE_SIGMA:
        if sig is None:
            raise ValueError("noise_mode 'sigma' without sigma")
        if tuple(sig.shape) != (bank.nK, bank.C):
            raise ValueError(f"sigma {tuple(sig.shape)} does not match the kernel bank {(bank.nK, bank.C)}")
    _check_indices("nidx", nidx, N, int(pl.shape[0]), dev, validate)
    return sig, pl, int(pl.shape[0])


def degrade_batch_stats(hr: torch.Tensor, kbank, *, kidx=None, sigma=None, pool=None, nidx=None, factor: int = 8,
                        pad_mode: str = "replicate", down_mode: str = "boxmean", noise_mode: str | None = None,
                        out: torch.Tensor | None = None, sums: torch.Tensor | None = None, algo: str = "auto",
                        validate: bool = False):
    """degrade_batch + band_stats of the HR patches in one pass (E_make_train_data.py:223-250 with
    data_mean_std.py:32-33 fused): returns (lr, mean [N,C] f64, std [N,C] f64); `sums` (f64 [2C+1]) is
    accumulated as in band_stats.  hr: contiguous CUDA float32 [N, C, H, W]."""
    require_cuda()
    if not hr.is_cuda or hr.dtype != torch.float32 or hr.ndim != 4:
        raise TypeError("degrade_batch_stats expects a CUDA float32 tensor [N,C,H,W]")
    dev = hr.device
    dm, pm = L.DOWN_MODES[down_mode], L.PAD_MODES[pad_mode]
    if noise_mode is None:
        noise_mode = "none" if nidx is None else ("sigma" if sigma is not None else "add")
    nm = L.NOISE_MODES[noise_mode]
    bank = kbank if isinstance(kbank, PreparedBank) else prepare_kernels(kbank.to(dev), factor, dm)
    if bank.factor != factor or bank.down_mode != dm:
        raise ValueError("PreparedBank was built for a different factor / down_mode")
    N, Cc, H, W = hr.shape
    if not hr[0].is_contiguous() or (N > 1 and hr.stride(0) < Cc * H * W):
        hr = hr.contiguous()
    assert Cc == bank.C, f"kernel bands ({bank.C}) != image bands ({Cc})"
    Ho, Wo = L.degrade_out_size(H, W, bank.kh, bank.kw, factor, dm)
    if out is None:
        out = torch.empty((N, Cc, Ho, Wo), dtype=torch.float32, device=dev)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (N, Cc, Ho, Wo)):
        raise ValueError(f"out must be a contiguous CUDA float32 tensor {(N, Cc, Ho, Wo)}")
    if sums is not None and not (sums.is_cuda and sums.is_contiguous() and sums.dtype == torch.float64
                                 and sums.numel() == 2 * Cc + 1):
        raise ValueError(f"sums must be a contiguous CUDA float64 tensor of {2 * Cc + 1} elements")
    _check_indices("kidx", kidx, N, bank.nK, dev, validate)
    kidx_d, nidx_d = _i32(kidx, dev), _i32(nidx, dev)
    sig, pl, npool = _noise_args(nm, sigma, pool, nidx, N, bank, (Cc, Ho, Wo), dev, validate)
    mean = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    std = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    wsb = int(L.check(L.lib().kmsr_degrade_stats_workspace_bytes(N, Cc)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_degrade_stats_prepared(
            _ptr(hr), N, Cc, H, W, hr.stride(0) if N > 1 else Cc * H * W,
            _ptr(bank.comp), _ptr(bank.dsum), bank.nK, bank.kh, bank.kw, _ptr(kidx_d),
            _ptr(sig), _ptr(pl), npool, _ptr(nidx_d), factor, pm, dm, nm, _ptr(out), _ptr(mean), _ptr(std),
            _ptr(sums), _ptr(ws), wsb, L.ALGOS[algo], _stream(dev)))
    return out, mean, std


def add_noise_batch(blurred: torch.Tensor, pool: torch.Tensor, nidx, *, sigma=None, kidx=None,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """out[n] = blurred[n] + scale * pool[nidx[n]]  (E_make_train_data.py:72-74; scale 1 unless sigma)."""
    require_cuda()
    dev = blurred.device
    b = blurred.contiguous()
    N, Cc = b.shape[0], b.shape[1]
    hw = int(np.prod(b.shape[2:]))
    pl = pool if pool.is_cuda else pool.to(dev)
    pl = pl.contiguous()
    if tuple(pl.shape[1:]) != tuple(b.shape[1:]):
        raise ValueError(f"noise pool {tuple(pl.shape)} does not match patches {tuple(b.shape)}")
    if out is None:
        out = torch.empty_like(b)
    sig = None if sigma is None else _f32(torch.as_tensor(sigma), dev).contiguous()
    _check_indices("nidx", nidx, N, int(pl.shape[0]), dev)
    if sig is not None:
        _check_indices("kidx", kidx, N, int(sig.shape[0]), dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_add_noise(_ptr(b), N, Cc, hw, _ptr(pl), pl.shape[0], _ptr(_i32(nidx, dev)),
                                       _ptr(sig), _ptr(_i32(kidx, dev)), _ptr(out), _stream(dev)))
    return out


def crop_sub(geo: torch.Tensor, den: torch.Tensor, top, left, crop: int) -> torch.Tensor:
    """pool[m] = (geo - den)[:, top[m]:top[m]+crop, left[m]:left[m]+crop]  (D_build_noise_pool.py:88, :51)."""
    require_cuda()
    dev = geo.device
    if geo.ndim != 3 or tuple(den.shape) != tuple(geo.shape):
        # the reference's `geo - den` (D:88) raises on a shape mismatch before any crop offset is drawn
        raise ValueError(f"geophysical_data {tuple(geo.shape)} and denoised {tuple(den.shape)} must be the same [C,H,W]")
    if geo.dtype != torch.float32 or den.dtype != torch.float32:
        raise TypeError("crop_sub expects float32 tensors")
    g = geo.contiguous()
    d = den.to(dev).contiguous()
    Cc, H, W = g.shape
    if crop > H or crop > W:
        raise ValueError(f"crop {crop} is larger than the image {H}x{W}")          # D:44-45
    ta, la = np.asarray(top), np.asarray(left)
    if ta.shape != la.shape:
        raise ValueError("top / left must have the same length")
    _check_indices("crop top", ta, int(ta.size), H - crop + 1, dev)
    _check_indices("crop left", la, int(la.size), W - crop + 1, dev)
    t = _i32(ta, dev)
    l = _i32(la, dev)
    n = int(t.numel())
    out = torch.empty((n, Cc, crop, crop), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_crop_sub(_ptr(g), _ptr(d), Cc, H, W, _ptr(t), _ptr(l), n, crop, _ptr(out),
                                      _stream(dev)))
    return out


def band_stats(x: torch.Tensor, sums: torch.Tensor | None = None):
    """Per-patch per-band NaN-skipping mean / population std (data_mean_std.py:32-33).

    x [N, C, ...] CUDA float32 -> (mean [N,C] f64, std [N,C] f64).  `sums` (f64 [2C+1], device) is
    accumulated in place: sum of means, sum of stds, patch count -- the vector the ranks all-reduce.
    """
    require_cuda()
    dev = x.device
    N, Cc = x.shape[0], x.shape[1]
    hw = int(np.prod(x.shape[2:]))
    flat = x.reshape(N, Cc, hw)
    if flat.stride(2) != 1 or flat.stride(1) != hw:
        flat = flat.contiguous()
    mean = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    std = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_band_stats(_ptr(flat), N, Cc, hw, flat.stride(0) if N > 1 else Cc * hw,
                                        _ptr(mean), _ptr(std), _ptr(sums), _stream(dev)))
    return mean, std


def water_mask(data: torch.Tensor, tmin: float, tmax: float, nir: int = 4, invalid: float = -9999.0,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """In place invalid->NaN on `data`, returns the masked copy (A_00_patch_cutter_universal.py:102-113)."""
    require_cuda()
    assert data.is_cuda and data.is_contiguous() and data.dtype == torch.float32
    Cc = data.shape[0]
    hw = int(np.prod(data.shape[1:]))
    if out is None:
        out = torch.empty_like(data)
    with torch.cuda.device(data.device):
        L.check(L.lib().kmsr_water_mask(_ptr(data), Cc, hw, nir, invalid, tmin, tmax, _ptr(out),
                                        _stream(data.device)))
    return out


def keep_mask(masked: torch.Tensor, patch_size: int = 256, stride: int = 128, nan_threshold: float = 0.0):
    """keep[i,j] / NaN count of every window (A_00_patch_cutter_universal.py:152-183)."""
    require_cuda()
    assert masked.is_cuda and masked.is_contiguous() and masked.dtype == torch.float32
    dev = masked.device
    Cc, H, W = masked.shape
    hp = (H - patch_size) // stride + 1 if H >= patch_size else 0
    wp = (W - patch_size) // stride + 1 if W >= patch_size else 0
    keep = torch.zeros((max(hp, 0), max(wp, 0)), dtype=torch.uint8, device=dev)
    cnt = torch.zeros((max(hp, 0), max(wp, 0)), dtype=torch.int32, device=dev)
    if hp > 0 and wp > 0:
        wsb = int(L.lib().kmsr_keep_mask_workspace_bytes(H, W, patch_size, stride))
        ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().kmsr_keep_mask(_ptr(masked), Cc, H, W, patch_size, stride, float(nan_threshold),
                                           _ptr(keep), _ptr(cnt), _ptr(ws), wsb, _stream(dev)))
    return keep.bool(), cnt


def scene_keep_mask(scene: torch.Tensor, tmin: float, tmax: float, patch_size: int = 256, stride: int = 128,
                    nan_threshold: float = 0.0, nir: int = 4, invalid: float = -9999.0):
    """keep[i,j] / NaN count of every window of the masked scene, computed from the RAW scene [C,H,W] in one read
    (no masked copy, no in-place fill replacement).  With nan_threshold == 0 the kept windows can be degraded
    straight from `scene` (A_00_patch_cutter_universal.py:89-123 + :152-183 fused)."""
    require_cuda()
    assert scene.is_cuda and scene.is_contiguous() and scene.dtype == torch.float32
    dev = scene.device
    Cc, H, W = scene.shape
    hp = (H - patch_size) // stride + 1 if H >= patch_size else 0
    wp = (W - patch_size) // stride + 1 if W >= patch_size else 0
    keep = torch.zeros((max(hp, 0), max(wp, 0)), dtype=torch.uint8, device=dev)
    cnt = torch.zeros((max(hp, 0), max(wp, 0)), dtype=torch.int32, device=dev)
    if hp > 0 and wp > 0:
        wsb = int(L.lib().kmsr_keep_mask_workspace_bytes(H, W, patch_size, stride))
        ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().kmsr_scene_keep_mask(_ptr(scene), Cc, H, W, nir, float(invalid), float(tmin), float(tmax),
                                                 patch_size, stride, float(nan_threshold), _ptr(keep), _ptr(cnt), _ptr(ws),
                                                 wsb, _stream(dev)))
    return keep.bool(), cnt


def _bands_nchw(x: torch.Tensor):
    if x.ndim != 4:
        raise ValueError(f"expected [N,C,H,W], got {tuple(x.shape)}")
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("expected a CUDA float32 tensor (drop-in wrappers copy host inputs)")
    N, Cc, H, W = x.shape
    if x.stride(3) != 1 or x.stride(2) != W or x.stride(1) != H * W:
        x = x.contiguous()
    return x, N, Cc, H, W, (x.stride(0) if N > 1 else Cc * H * W)


def estimate_sigma(x: torch.Tensor) -> torch.Tensor:
    """skimage.restoration.estimate_sigma of every band (denoise/denoise.py:47) -> [N,C] float64 on the device.
    NaN pixels are read as the band's nanmean (denoise.py:43-44); an all-NaN band reports 0.0 (:40-41)."""
    require_cuda()
    x, N, Cc, H, W, sn = _bands_nchw(x)
    dev = x.device
    sigma = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    wsb = L.check(int(L.lib().kmsr_denoise_workspace_bytes(N, Cc, H, W)))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_estimate_sigma(_ptr(x), N, Cc, H, W, sn, _ptr(sigma), _ptr(ws), wsb, _stream(dev)))
    return sigma


def denoise_nlm(x: torch.Tensor, h_factor: float = 1.15, patch_size: int = 7, patch_distance: int = 11,
                out: torch.Tensor | None = None):
    """denoise_band_float_nlm (denoise/denoise.py:34-65) of every band of x [N,C,H,W] -> (denoised [N,C,H,W] f32,
    sigma [N,C] f64): nanmean fill, estimate_sigma, fast-mode non-local means with h = h_factor * sigma, NaN restore."""
    require_cuda()
    x, N, Cc, H, W, sn = _bands_nchw(x)
    dev = x.device
    if out is None:
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (N, Cc, H, W)
    sigma = torch.empty((N, Cc), dtype=torch.float64, device=dev)
    wsb = L.check(int(L.lib().kmsr_denoise_workspace_bytes(N, Cc, H, W)))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().kmsr_denoise_nlm(_ptr(x), N, Cc, H, W, sn, float(h_factor), int(patch_size), int(patch_distance),
                                         _ptr(out), _ptr(sigma), _ptr(ws), wsb, _stream(dev)))
    return out, sigma


def fp32_peak_tflops(device=None, sustained_ms: float = 150.0) -> dict:
    """What this device does in packed FFMA2 (kmsr_fp32_probe), in TFLOP/s (2 flops per multiply-add): the FP32-FMA
    roofline denominator.  {"burst": best of several ~2 ms launches with pauses between them -- the figure for a kernel
    timed alone --, "sustained": one launch of ~sustained_ms, where the 1 kW power cap has pulled the SM clock down --
    the figure for a kernel inside a long step}.  A measurement aid for bench.py / the sweep, not part of the
    reference path."""
    import time
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    info = L.device_info(dev.index or 0)
    sink = torch.zeros(512 * info["sm_count"], dtype=torch.float32, device=dev)
    fma = C.c_double(0.0)

    def run(iters: int) -> float:
        with torch.cuda.device(dev):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.lib().kmsr_fp32_probe(_ptr(sink), int(iters), C.byref(fma), _stream(dev)))
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1)

    ms = run(2000)                                      # sizes the launches; also the warm-up
    per_iter = max(ms, 1e-3) / 2000
    burst = 0.0
    for _ in range(5):
        time.sleep(0.05)
        it = max(200, int(2.0 / per_iter))
        m = run(it)
        burst = max(burst, 2.0 * fma.value / (m * 1e-3) / 1e12)
    it = max(2000, int(sustained_ms / per_iter))
    m = run(it)
    sustained = 2.0 * fma.value / (m * 1e-3) / 1e12
    return {"burst": burst, "sustained": sustained}
