"""ctypes binding of libkmsr.so (include/kmsr.h).  Fails loudly: no library -> ImportError-like
RuntimeError on first use; a non-zero return code -> KmsrError carrying kmsr_last_error()."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KMSR_LIB selects another build of the same library (libkmsr_debug.so: device-side asserts; libkmsr_bench.so:
# measurement switches) -- never a different implementation
LIB_PATH = os.environ.get("KMSR_LIB") or os.path.join(_HERE, "libkmsr.so")

# enums of include/kmsr.h
PAD_REPLICATE, PAD_ZERO = 0, 1
DOWN_BOXMEAN, DOWN_DECIMATE = 0, 1
NOISE_NONE, NOISE_ADD, NOISE_SIGMA = 0, 1, 2
ALGO_AUTO, ALGO_TILED, ALGO_TMA, ALGO_STREAM, ALGO_REG, ALGO_BOX = 0, 1, 2, 3, 4, 5
E_INVALID, E_UNSUPPORTED, E_CUDA, E_ALIGN = -1, -2, -3, -4

PAD_MODES = {"replicate": PAD_REPLICATE, "zero": PAD_ZERO}
DOWN_MODES = {"boxmean": DOWN_BOXMEAN, "decimate": DOWN_DECIMATE}
NOISE_MODES = {"none": NOISE_NONE, "add": NOISE_ADD, "sigma": NOISE_SIGMA}
ALGOS = {"auto": ALGO_AUTO, "tiled": ALGO_TILED, "tma": ALGO_TMA, "stream": ALGO_STREAM, "reg": ALGO_REG,
         "box": ALGO_BOX}


class KmsrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libkmsr error {code}: {message}")
        self.code = code
        self.message = message


_i32, _i64, _f32, _f64, _vp = C.c_int, C.c_int64, C.c_float, C.c_double, C.c_void_p
_pi = C.POINTER(C.c_int)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); device pointers travel as c_void_p
SIGNATURES = {
    "kmsr_version": (_i32, []),
    "kmsr_last_error": (C.c_char_p, []),
    "kmsr_device_info": (_i32, [_i32, _pi, _pi, _pi, _pi64, _pi64]),
    "kmsr_validate_indices": (_i32, [_vp, _i64, _i64, _vp, C.c_char_p, _vp]),
    "kmsr_fp32_probe": (_i32, [_vp, _i32, C.POINTER(C.c_double), _vp]),
    "kmsr_degrade_out_size": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _pi, _pi]),
    "kmsr_composite_size": (_i32, [_i32, _i32, _i32, _i32, _pi, _pi, _pi]),
    "kmsr_degrade_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "kmsr_prepare_kernels": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "kmsr_degrade_prepared": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64, _i64, _i64, _vp,
                                     _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                                     _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "kmsr_degrade_windows": (_i32, [_vp, _i32, _i32, _i32, _i64, _i64, _vp, _i64, _i32, _i32, _i32,
                                    _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                                    _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "kmsr_degrade_stats_workspace_bytes": (_i64, [_i64, _i32]),
    "kmsr_degrade_stats_prepared": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64,
                                           _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                                           _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "kmsr_degrade_batch": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64, _i64, _i64, _vp,
                                  _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                                  _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _vp]),
    "kmsr_add_noise": (_i32, [_vp, _i64, _i32, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "kmsr_crop_sub": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _vp, _vp]),
    "kmsr_band_stats": (_i32, [_vp, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _vp]),
    "kmsr_water_mask": (_i32, [_vp, _i32, _i64, _i32, _f32, _f32, _f32, _vp, _vp]),
    "kmsr_keep_mask_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "kmsr_keep_mask": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _i64, _vp]),
    "kmsr_scene_keep_mask": (_i32, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _i32, _f64, _vp, _vp, _vp, _i64, _vp]),
    "kmsr_denoise_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "kmsr_estimate_sigma": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64, _vp, _vp, _i64, _vp]),
    "kmsr_denoise_nlm": (_i32, [_vp, _i64, _i32, _i32, _i32, _i64, _f64, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "kmsr_selector_weight_floats": (_i64, [_i32, _i32]),
    "kmsr_selector_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "kmsr_selector_logits": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kmsr_selector_umma_supported": (_i32, [_i32, _i32]),
    "kmsr_selector_umma_weight_floats": (_i64, [_i32, _i32]),
    "kmsr_selector_umma_workspace_bytes": (_i64, [_i64, _i32, _i32]),
    "kmsr_selector_logits_umma": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kmsr_launch_count": (_i64, []),
    "kmsr_last_algo": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load libkmsr.so once.  The product path never falls back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C kernel-modeling-super-resolution_b200/csrc`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)        # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return (lib().kmsr_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> int:
    if rc < 0:
        raise KmsrError(int(rc), last_error())
    return rc


def launch_count() -> int:
    return int(lib().kmsr_launch_count())


def last_algo() -> str:
    return (lib().kmsr_last_algo() or b"").decode()


def degrade_out_size(H, W, kh, kw, factor, down_mode=DOWN_BOXMEAN):
    ho, wo = C.c_int(), C.c_int()
    check(lib().kmsr_degrade_out_size(H, W, kh, kw, factor, down_mode, C.byref(ho), C.byref(wo)))
    return ho.value, wo.value


def composite_size(kh, kw, factor, down_mode=DOWN_BOXMEAN):
    a, b, s = C.c_int(), C.c_int(), C.c_int()
    check(lib().kmsr_composite_size(kh, kw, factor, down_mode, C.byref(a), C.byref(b), C.byref(s)))
    return a.value, b.value, s.value


def device_info(device: int = 0) -> dict:
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    l2, sh = C.c_int64(), C.c_int64()
    check(lib().kmsr_device_info(device, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2), C.byref(sh)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value, "smem_optin": sh.value}
