"""Deterministic synthetic "Landsat-shaped" inputs.

Everything here is numpy-only elementwise arithmetic on legacy MT19937 draws, so
the same seed gives the same bytes on the build container and on the GPU box
(no BLAS, no libm transcendental in the value path beyond RandomState's own).
Used by tests/, bench.py and __graft_entry__.smoke(); it is input generation,
not part of the product path.

Shapes follow the reference: HR patch `[5, 256, 256] f32` CHW with band order
443/490/555/660/865 nm (C_30:49), noise pool `(N, 5, 32, 32) f32` (D:110).
"""
from __future__ import annotations

import numpy as np

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]
BAND_BASE = np.array([80.0, 70.0, 50.0, 25.0, 8.0], dtype=np.float64)
# per-band GOCI noise scale, muti_kernel/train.py:212
POOL_SIGMA = np.array([0.55, 0.72, 0.83, 0.63, 0.19], dtype=np.float64)

REGIMES = {
    "textured": (5.0, 0.5),   # amplitude of the smooth field, white-noise sigma
    "water": (0.3, 0.03),     # low dynamic range: the fp32 accuracy stress case
}


def _bilinear_up(coarse: np.ndarray, size: int) -> np.ndarray:
    """Bilinear upsample [..., g, g] -> [..., size, size] with explicit f64 arithmetic."""
    g = coarse.shape[-1]
    pos = (np.arange(size, dtype=np.float64) + 0.5) * (g / size) - 0.5
    pos = np.clip(pos, 0.0, g - 1.0)
    i0 = np.floor(pos).astype(np.int64)
    i1 = np.minimum(i0 + 1, g - 1)
    t = pos - i0
    rows = coarse[..., i0, :] * (1.0 - t)[:, None] + coarse[..., i1, :] * t[:, None]
    return rows[..., :, i0] * (1.0 - t) + rows[..., :, i1] * t


def make_hr(n: int, seed: int, regime: str = "textured", size: int = 256,
            bands: int = 5, grid: int = 32) -> np.ndarray:
    """`[n, bands, size, size] f32` NaN-free radiance-like patches."""
    amp, sn = REGIMES[regime]
    rs = np.random.RandomState(seed)
    g = min(grid, size)
    out = np.empty((n, bands, size, size), dtype=np.float32)
    base = BAND_BASE[np.arange(bands) % 5]
    for i in range(n):
        coarse = rs.standard_normal((bands, g, g))
        white = rs.standard_normal((bands, size, size))
        field = _bilinear_up(coarse, size)
        out[i] = (base[:, None, None] + amp * field + sn * white).astype(np.float32)
    return out


def make_noise_pool(n: int = 4096, seed: int = 42, size: int = 32, bands: int = 5) -> np.ndarray:
    """`(n, bands, size, size) f32` zero-mean pool with the per-band GOCI sigmas."""
    rs = np.random.RandomState(seed)
    z = rs.standard_normal((n, bands, size, size))
    return (z * POOL_SIGMA[np.arange(bands) % 5][None, :, None, None]).astype(np.float32)


def softmax_kernels(k: int, seed: int = 7, bands: int = 5, n: int | None = None) -> np.ndarray:
    """Sweep kernels `softmax(randn(bands,k,k))` (sum to one per band), f32."""
    rs = np.random.RandomState(seed)
    shape = (bands, k * k) if n is None else (n, bands, k * k)
    z = rs.standard_normal(shape)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    p = e / e.sum(axis=-1, keepdims=True)
    return p.reshape(shape[:-1] + (k, k)).astype(np.float32)


def make_scene(seed: int, height: int, width: int, bands: int = 5,
               n_fill: int = 6, n_cloud: int = 6) -> np.ndarray:
    """Scene `[bands, H, W] f32` with -9999 fill blobs and NIR>7 "cloud" blobs (config 4).

    NIR (band 4) sits around 3 so that it passes the 1e-6..7.0 water test
    (A_00_patch_cutter_universal.py:32-33) except inside the cloud discs.
    """
    rs = np.random.RandomState(seed)
    g = 16
    coarse = rs.standard_normal((bands, g, g))
    field = np.empty((bands, height, width), dtype=np.float64)
    # separable bilinear to a non-square target
    py = np.clip((np.arange(height) + 0.5) * (g / height) - 0.5, 0, g - 1.0)
    px = np.clip((np.arange(width) + 0.5) * (g / width) - 0.5, 0, g - 1.0)
    y0 = np.floor(py).astype(np.int64); y1 = np.minimum(y0 + 1, g - 1); ty = py - y0
    x0 = np.floor(px).astype(np.int64); x1 = np.minimum(x0 + 1, g - 1); tx = px - x0
    rows = coarse[:, y0, :] * (1 - ty)[None, :, None] + coarse[:, y1, :] * ty[None, :, None]
    field[:] = rows[:, :, x0] * (1 - tx) + rows[:, :, x1] * tx
    base = np.array([80.0, 70.0, 50.0, 25.0, 3.0])[np.arange(bands) % 5]
    scene = (base[:, None, None] + 0.4 * field).astype(np.float32)
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(n_fill):
        cy, cx = rs.randint(0, height), rs.randint(0, width)
        r = rs.randint(max(2, height // 40), max(3, height // 12))
        scene[:, (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = -9999.0
    for _ in range(n_cloud):
        cy, cx = rs.randint(0, height), rs.randint(0, width)
        r = rs.randint(max(2, height // 40), max(3, height // 12))
        scene[bands - 1, (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 9.5
    return scene
