"""kmsr_b200 -- B200-native LR/HR training-pair synthesis.

Drop-in for one hot path of Zhiyyeah/Kernel-Modeling-Super-Resolution: degrade Landsat HR patches
with estimated blur kernels, inject noise-pool patches, assemble pairs, per-band statistics.
The arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of `libkmsr.so`
(include/kmsr.h); this package is the host side: the reference's Python entry points with their
signatures and layouts, bound to that library with ctypes.  There is no CPU arithmetic fallback:
calling a compute entry point without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]
