"""ctypes front-end of oracle/liboracle.so (plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_nlm.c")]
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def normalize_kernel(k: np.ndarray) -> np.ndarray:
    k = np.ascontiguousarray(k, dtype=np.float32)
    out = np.empty_like(k)
    lib().orc_normalize_kernel(_p(k, C.c_float), k.shape[0], k.shape[1], k.shape[2], _p(out, C.c_float))
    return out


def _out_hw(h, w, factor, decimate):
    if decimate:
        return (h + factor - 1) // factor, (w + factor - 1) // factor
    steps = int(np.log2(factor))
    for _ in range(steps):
        h, w = h // 2, w // 2
    return h, w


def degrade(img: np.ndarray, kn: np.ndarray, factor: int = 8, zero_pad=False, decimate=False, f64=False):
    img = np.ascontiguousarray(img, dtype=np.float32)
    kn = np.ascontiguousarray(kn, dtype=np.float32)
    c, h, w = img.shape
    ho, wo = _out_hw(h, w, factor, decimate)
    out = np.empty((c, ho, wo), dtype=np.float64 if f64 else np.float32)
    fn = lib().orc_degrade_f64 if f64 else lib().orc_degrade_f32
    rc = fn(_p(img, C.c_float), c, h, w, _p(kn, C.c_float), kn.shape[1], kn.shape[2], factor,
            int(zero_pad), int(decimate), _p(out, C.c_double if f64 else C.c_float))
    assert rc == 0
    return out


def numpy_randint_stream(seed: int, high: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().orc_numpy_randint_stream(C.c_uint32(seed), C.c_int64(high), C.c_int64(n), _p(out, C.c_int64))
    return out


def numpy_two_randint_vectors(seed: int, nk: int, npool: int, n: int):
    a = np.empty(n, dtype=np.int32); b = np.empty(n, dtype=np.int32)
    lib().orc_numpy_two_randint_vectors(C.c_uint32(seed), C.c_int64(nk), C.c_int64(npool), C.c_int64(n),
                                        _p(a, C.c_int32), _p(b, C.c_int32))
    return a, b


def python_crop_offsets(seed: int, shapes, crop: int, spf: int) -> np.ndarray:
    hw = np.ascontiguousarray(np.array(shapes, dtype=np.int32).reshape(-1, 2))
    out = np.empty((hw.shape[0] * spf, 2), dtype=np.int32)
    lib().orc_python_crop_offsets(C.c_uint32(seed), _p(hw, C.c_int32), C.c_int64(hw.shape[0]), crop, spf,
                                  _p(out, C.c_int32))
    return out


def crop_sub(geo, den, top, left, crop):
    geo = np.ascontiguousarray(geo, dtype=np.float32); den = np.ascontiguousarray(den, dtype=np.float32)
    c, h, w = geo.shape
    out = np.empty((c, crop, crop), dtype=np.float32)
    lib().orc_crop_sub(_p(geo, C.c_float), _p(den, C.c_float), c, h, w, int(top), int(left), crop, _p(out, C.c_float))
    return out


def add_noise(blurred, noise, scale=None):
    blurred = np.ascontiguousarray(blurred, dtype=np.float32); noise = np.ascontiguousarray(noise, dtype=np.float32)
    c = blurred.shape[0]; hw = blurred[0].size
    out = np.empty_like(blurred)
    sp = None if scale is None else _p(np.ascontiguousarray(scale, dtype=np.float32), C.c_float)
    lib().orc_add_noise(_p(blurred, C.c_float), _p(noise, C.c_float), sp, c, hw, _p(out, C.c_float))
    return out


def band_stats_f64(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = x.shape[0]; hw = x[0].size
    m = np.empty(c); s = np.empty(c)
    lib().orc_band_stats_f64(_p(x, C.c_float), c, C.c_int64(hw), _p(m, C.c_double), _p(s, C.c_double))
    return m, s


def water_mask(data, tmin, tmax, nir=4):
    data = np.ascontiguousarray(data, dtype=np.float32)
    c = data.shape[0]; hw = data[0].size
    out = np.empty_like(data)
    lib().orc_water_mask(_p(data, C.c_float), c, C.c_int64(hw), nir, C.c_float(tmin), C.c_float(tmax), _p(out, C.c_float))
    return out


def keep_mask(masked, patch=256, stride=128, nan_thr=0.0):
    masked = np.ascontiguousarray(masked, dtype=np.float32)
    c, h, w = masked.shape
    hp, wp = (h - patch) // stride + 1, (w - patch) // stride + 1
    keep = np.zeros((max(hp, 0), max(wp, 0)), dtype=np.uint8)
    if hp > 0 and wp > 0:
        lib().orc_keep_mask(_p(masked, C.c_float), c, h, w, patch, stride, C.c_double(nan_thr), _p(keep, C.c_uint8), hp, wp)
    return keep.astype(bool)


# ---- f4: upstream denoise stage (denoise/denoise.py:34-65), oracle_nlm.c -- PARITY UNPINNED (see its header)
def estimate_sigma(img: np.ndarray, return_dd: bool = False, f64: bool = False):
    """skimage.restoration.estimate_sigma of one 2-D float32 image (db2 'dd' coefficients, MAD / 0.6745).
    f64=True evaluates the transform in float64 (the value the float32 transform approximates)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    if f64:
        fn = lib().orc_estimate_sigma_f64
        fn.restype = C.c_double
        return float(fn(_p(img, C.c_float), h, w))
    dd = np.empty(((h + 3) // 2, (w + 3) // 2), dtype=np.float32) if return_dd else None
    fn = lib().orc_estimate_sigma
    fn.restype = C.c_double
    s = fn(_p(img, C.c_float), h, w, None if dd is None else _p(dd, C.c_float))
    return (float(s), dd) if return_dd else float(s)


def nlm_fast_f32(img: np.ndarray, h: float, sigma: float, patch_size: int = 7, patch_distance: int = 11) -> np.ndarray:
    """skimage denoise_nl_means(fast_mode=True) of one float32 band, integral-image algorithm in float32."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.empty_like(img)
    rc = lib().orc_nlm_fast_f32(_p(img, C.c_float), img.shape[0], img.shape[1], patch_size, patch_distance,
                                C.c_float(np.float32(h)), C.c_float(np.float32(sigma * sigma)), _p(out, C.c_float))
    assert rc == 0
    return out


def nlm_exact_f64(img: np.ndarray, h: float, sigma: float, patch_size: int = 7, patch_distance: int = 11,
                  eps: float = 0.0):
    """The same formula in float64.  Returns (out, flip): flip bounds what the distance cut-off can change per pixel
    when a distance moves by less than eps (None when eps == 0)."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.empty(img.shape, dtype=np.float64)
    flip = np.empty(img.shape, dtype=np.float64) if eps > 0 else None
    rc = lib().orc_nlm_exact_f64(_p(img, C.c_float), img.shape[0], img.shape[1], patch_size, patch_distance,
                                 C.c_float(np.float32(h)), C.c_float(np.float32(sigma * sigma)), C.c_double(eps),
                                 _p(out, C.c_double), None if flip is None else _p(flip, C.c_double))
    assert rc == 0
    return out, flip
