"""Host-side index draws.  The reference takes every random decision on the host with two
MT19937 front-ends; reproducing them bit for bit means issuing the same library calls in the
same order, never regenerating on the device:

* noise index per accepted file: `np.random.seed(seed)` once (E_make_train_data.py:190) then one
  `np.random.randint(0, len(pool))` per file (E:72).  A single vector draw is bit-identical to the
  scalar draws one by one.
* crop offsets: `random.seed(seed)` once (D_build_noise_pool.py:65) then per sample
  `random.randint(0, H - crop)` and `random.randint(0, W - crop)` (inclusive, D:49-50).
* multi-kernel composition (build-defined): one `RandomState(seed)`;
  kernel picks first, noise picks second.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np


def draw_noise_indices(n: int, pool_len: int, seed: int | None = None) -> np.ndarray:
    """n draws from the GLOBAL legacy numpy stream, as E.add_noise does; seeds it first if given."""
    if pool_len <= 0:
        raise ValueError("empty noise pool")
    if seed is not None:
        np.random.seed(seed)
    if n == 0:
        return np.zeros(0, dtype=np.int32)
    return np.random.randint(0, pool_len, size=n).astype(np.int32)


def draw_multi_kernel_indices(n: int, n_kernels: int, pool_len: int, seed: int = 42):
    rs = np.random.RandomState(seed)
    kidx = rs.randint(0, n_kernels, n)
    nidx = rs.randint(0, pool_len, n)
    return kidx.astype(np.int32), nidx.astype(np.int32)


def draw_crop_offsets(height: int, width: int, crop_size: int, n_samples: int):
    """(top, left) pairs from the GLOBAL CPython `random` stream, top first (D:49-50)."""
    if height < crop_size or width < crop_size:
        raise ValueError(f"image {height}x{width} is smaller than the crop size {crop_size}")   # D:44-45
    top = np.empty(n_samples, dtype=np.int32)
    left = np.empty(n_samples, dtype=np.int32)
    for i in range(n_samples):
        top[i] = _pyrandom.randint(0, height - crop_size)
        left[i] = _pyrandom.randint(0, width - crop_size)
    return top, left


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous patch range of a rank: [floor(r*n/G), floor((r+1)*n/G))  (SURVEY.md 8e)."""
    return (rank * n) // world, ((rank + 1) * n) // world
