"""The reference arm of bench.py  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product).

`bench.py --impl reference` and the `cpu_baseline` leg time the reference's own CPU implementation of the path.
The reference is pure Python: its hot-path arithmetic is `apply_kernel_degradation`
(kernel_from_lr_gan/C_31apply_muti_kernel_to_landsat.py:59-97).  When the VERBATIM reference file is available it
is loaded by path and called unmodified (`kind = "reference"`); it is searched for in

    1. $KMSR_REFERENCE_ROOT            (a checkout of the reference)
    2. <repo>/baseline/_ref            (git-ignored copy made by __graft_entry__.build() in the build container;
                                        it travels to the GPU box with the snapshot, the reference tree does not)
    3. /root/reference                 (build container)

and otherwise the call-site port in oracle/kmsr_oracle.py runs (`kind = "port"`: the same torch / numpy calls at the
same call sites).  netCDF4 / matplotlib are absent from the image and never touched by the hot-path functions, so
empty stand-in modules satisfy the module-level imports (C_31:18-19).

BASELINE config 2 is a composition the reference has no single function for: per patch
`lr = apply_kernel_degradation(hr[n], K[kidx[n]], f) + sigma[kidx[n]][:, None, None] * pool[nidx[n]]`
(kernel pick and sigma semantics: muti_kernel/train_gemini.py:107-115, :137; injection: E_make_train_data.py:72-74),
one patch per call exactly as the reference's folder loop does (C_31:147-150).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REL = "kernel_from_lr_gan/C_31apply_muti_kernel_to_landsat.py"
_cache: dict = {}


def _candidates():
    env = os.environ.get("KMSR_REFERENCE_ROOT")
    return [p for p in (env, os.path.join(ROOT, "baseline", "_ref"), "/root/reference") if p]


def _stub_missing():
    for name in ("netCDF4", "matplotlib", "matplotlib.pyplot", "tqdm"):
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "netCDF4":
                m.Dataset = object
            if name == "tqdm":
                m.tqdm = lambda it=None, *a, **k: it
            sys.modules[name] = m
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])


def load_apply():
    """-> (apply_kernel_degradation, kind, where): the verbatim reference function when a copy exists, else the port."""
    if "apply" in _cache:
        return _cache["apply"]
    for root in _candidates():
        path = os.path.join(root, _REL)
        if os.path.isfile(path):
            try:
                _stub_missing()
                spec = importlib.util.spec_from_file_location("_kmsr_ref_C31", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _cache["apply"] = (mod.apply_kernel_degradation, "reference", path)
                return _cache["apply"]
            except Exception as e:                      # an unloadable copy must not take the arm down
                print(f"refarm: could not load {path}: {e!r}; using the port", file=sys.stderr)
    from oracle import kmsr_oracle as orc
    _cache["apply"] = (orc.apply_kernel_degradation, "port", "oracle/kmsr_oracle.py")
    return _cache["apply"]


def multi_kernel_pairs(hr, kbank, sbank, pool, kidx, nidx, factor):
    """BASELINE config 2 on the CPU, one patch per call (C_31:147-150): returns lr [N,C,Ho,Wo] float32."""
    apply_fn, _, _ = load_apply()
    out = []
    for i in range(hr.shape[0]):
        k = int(kidx[i])
        lr = apply_fn(torch.from_numpy(hr[i]), torch.from_numpy(kbank[k]), factor).numpy()
        # the scaled add as one fp32 fused multiply-add (exact product, one rounding), as oracle/kmsr_oracle.py does
        out.append((lr.astype(np.float64) + sbank[k].astype(np.float64)[:, None, None] * pool[int(nidx[i])]).astype(np.float32))
    return np.stack(out)
