"""Load the *real* reference modules by path (build container only).

/root/reference is read-only and absent on the GPU box, so this helper is used
only by tests/golden/make_golden.py (fixture generation) and by the optional
`-m "not gpu"` cross-checks that skip when the reference tree is absent.
netCDF4 / matplotlib are not installed here; the hot-path functions never touch
them, so empty stand-in modules are enough to satisfy the module-level imports
(C_30:12-13, C_31:18-19, D:15, E:18-19).
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("KMSR_REFERENCE_ROOT", "/root/reference")
_REL = {
    "C30": "kernel_from_lr_gan/C_30apply_kernel_to_landsat.py",
    "C31": "kernel_from_lr_gan/C_31apply_muti_kernel_to_landsat.py",
    "D": "kernel_from_lr_gan/D_build_noise_pool.py",
    "E": "kernel_from_lr_gan/E_make_train_data.py",
    "S": "kernel_from_lr_gan/data_mean_std.py",
    "CUT": "kernel_from_lr_gan/A_00_patch_cutter_universal.py",
}


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, _REL["C30"]))


def _stub_missing():
    for name in ("netCDF4", "matplotlib", "matplotlib.pyplot"):
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "netCDF4":
                m.Dataset = object
            sys.modules[name] = m
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])


def load(tag: str):
    """Return the reference module for tag in {C30,C31,D,E,S,CUT}."""
    _stub_missing()
    path = os.path.join(REF_ROOT, _REL[tag])
    spec = importlib.util.spec_from_file_location(f"_kmsr_ref_{tag}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
