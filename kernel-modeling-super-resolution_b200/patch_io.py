"""Patch-file I/O of the folder drivers: the reference's NetCDF4 group layout, with a portable stand-in.

The reference keeps every patch in a NetCDF4 file with groups `geophysical_data`, `navigation_data`, `denoised`,
`blurred`, `hr`, `lr` and one 2-D `f4` variable per band (README.MD:1-10, C_30:174-196, C_31:158-178, E:84-117).
libnetcdf / the `netCDF4` module is not part of this image, so two backends sit behind one interface:

* `*.nc`  -- through `netCDF4` when it is importable, issuing the same calls as the reference's readers/writers;
             without the module a `.nc` file fails with ImportError, which the folder drivers treat like any other
             per-file failure (print, continue: C_30:205-209, E:264-267);
* `*.npz` -- a group container with keys "<group>/<variable>" (plus "__attrs__/<name>"), lossless float32 like the
             reference's zlib `f4` variables.  `nc_to_npz` / `npz_to_nc` convert when netCDF4 is present.

I/O is not on the hot path; these helpers only carry arrays between files and the GPU path.
"""
from __future__ import annotations

import os
import shutil

import numpy as np

BAND_NAMES = ["L_TOA_443", "L_TOA_490", "L_TOA_555", "L_TOA_660", "L_TOA_865"]   # C_30:49, D:23, E:28
PATCH_EXTS = (".nc", ".npz")


def _nc():
    try:
        import netCDF4  # noqa: PLC0415
    except Exception as e:  # noqa: BLE001
        raise ImportError("reading/writing .nc patch files needs the netCDF4 module (absent here); "
                          "use the .npz group container or convert with patch_io.nc_to_npz") from e
    return netCDF4


def is_patch_file(name: str, exts=PATCH_EXTS) -> bool:
    return name.endswith(exts)


def _filled(arr) -> np.ndarray:
    if isinstance(arr, np.ma.MaskedArray):                    # C_30:54-55, E:39-40
        arr = arr.filled(np.nan)
    return np.array(arr, dtype=np.float32)


def _load_npz(path: str) -> dict:
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def read_group_bands(path: str, group: str, band_names=BAND_NAMES) -> np.ndarray:
    """(C, H, W) float32 of one group; ValueError when the group or a band is missing (E:33-38)."""
    if path.endswith(".npz"):
        z = _load_npz(path)
        if not any(k.startswith(group + "/") for k in z):
            raise ValueError(f"group {group} does not exist in {path}")
        bands = []
        for b in band_names:
            key = f"{group}/{b}"
            if key not in z:
                raise ValueError(f"band {b} does not exist in group {group}")
            bands.append(_filled(z[key]))
        return np.stack(bands, axis=0)
    with _nc().Dataset(path, "r") as ds:
        if group not in ds.groups:
            raise ValueError(f"group {group} does not exist in {path}")
        grp = ds.groups[group]
        bands = []
        for b in band_names:
            if b not in grp.variables:
                raise ValueError(f"band {b} does not exist in group {group}")
            bands.append(_filled(grp.variables[b][:]))
        return np.stack(bands, axis=0)


def read_navigation(path: str) -> dict:
    """{'latitude': f32 array, 'longitude': f32 array} of group navigation_data (E:45-58)."""
    out = {}
    if path.endswith(".npz"):
        z = _load_npz(path)
        if not any(k.startswith("navigation_data/") for k in z):
            raise ValueError(f"navigation_data group does not exist in {path}")
        for v in ("latitude", "longitude"):
            if f"navigation_data/{v}" in z:
                out[v] = _filled(z[f"navigation_data/{v}"])
        return out
    with _nc().Dataset(path, "r") as ds:
        if "navigation_data" not in ds.groups:
            raise ValueError(f"navigation_data group does not exist in {path}")
        grp = ds.groups["navigation_data"]
        for v in ("latitude", "longitude"):
            if v in grp.variables:
                out[v] = _filled(grp.variables[v][:])
    return out


def write_groups(path: str, groups: dict, attrs: dict | None = None) -> None:
    """Create a patch file from {group: {variable: 2-D array}} (used by tests, converters and the cutter)."""
    if path.endswith(".npz"):
        flat = {f"{g}/{v}": np.asarray(a, dtype=np.float32) for g, vs in groups.items() for v, a in vs.items()}
        for k, v in (attrs or {}).items():
            flat[f"__attrs__/{k}"] = np.array(str(v))
        np.savez_compressed(path, **flat)
        return
    with _nc().Dataset(path, "w", format="NETCDF4") as ds:
        for g, vs in groups.items():
            grp = ds.createGroup(g)
            for v, a in vs.items():
                a = np.asarray(a, dtype=np.float32)
                dims = []
                for j, n in enumerate(a.shape):
                    name = f"{v}_dim_{j}" if g == "navigation_data" else ("y", "x")[j] if a.ndim == 2 else f"d{j}"
                    if name not in grp.dimensions:
                        grp.createDimension(name, n)
                    dims.append(name)
                var = grp.createVariable(v, "f4", tuple(dims), zlib=True, complevel=4)
                var[:] = a
        for k, v in (attrs or {}).items():
            setattr(ds, k, v)


def add_group(path: str, group: str, bands: np.ndarray, band_names=BAND_NAMES, dims=("y", "x"),
              history: str | None = None, long_name: str | None = None, src: str | None = None,
              group_attrs: dict | None = None) -> None:
    """Write `bands` [C,h,w] as group `group` of `path`; with `src` the file is first copied from it
    (C_30:171: shutil.copy then open in append mode); existing variables are overwritten (C_31:170-173).
    `group_attrs` become attributes of the group (denoise/denoise.py:236-252)."""
    if src is not None:
        shutil.copy(src, path)
    if path.endswith(".npz"):
        z = _load_npz(path)
        for c, b in enumerate(band_names[:bands.shape[0]]):
            z[f"{group}/{b}"] = np.asarray(bands[c], dtype=np.float32)
        if history is not None:
            z["__attrs__/history"] = np.array(history)
        for k, v in (group_attrs or {}).items():
            z[f"__attrs__/{group}/{k}"] = np.array(v)
        np.savez_compressed(path, **z)
        return
    with _nc().Dataset(path, "a", format="NETCDF4") as ds:
        for name, n in zip(dims, bands.shape[1:]):
            if name not in ds.dimensions:
                ds.createDimension(name, n)
        grp = ds.groups[group] if group in ds.groups else ds.createGroup(group)
        for c, b in enumerate(band_names[:bands.shape[0]]):
            var = grp.variables[b] if b in grp.variables else grp.createVariable(b, "f4", dims, zlib=True)
            var[:] = bands[c]
            if long_name:
                var.long_name = long_name.format(wl=b.split("_")[-1])
            var.units = "W m-2 sr-1 um-1"
        for k, v in (group_attrs or {}).items():
            setattr(grp, k, v)
        if history is not None:
            ds.history = history


def write_training_sample(path: str, hr: np.ndarray, lr: np.ndarray, nav: dict, band_names=BAND_NAMES) -> None:
    """E:77-117: groups hr, lr (one f4 variable per band) and navigation_data."""
    if path.endswith(".npz"):
        groups = {"hr": {b: hr[i] for i, b in enumerate(band_names)}, "lr": {b: lr[i] for i, b in enumerate(band_names)}}
        if nav:
            groups["navigation_data"] = {k: v for k, v in nav.items() if v is not None and v.size > 0}
        write_groups(path, groups)
        return
    with _nc().Dataset(path, "w", format="NETCDF4") as ds:
        for gname, arr in (("hr", hr), ("lr", lr)):
            grp = ds.createGroup(gname)
            grp.createDimension("band", arr.shape[0])
            grp.createDimension("y", arr.shape[1])
            grp.createDimension("x", arr.shape[2])
            for i, b in enumerate(band_names):
                var = grp.createVariable(b, "f4", ("y", "x"), zlib=True, complevel=4)
                var[:] = arr[i]
        if nav:
            grp = ds.createGroup("navigation_data")
            for key, value in nav.items():
                if value is not None and value.size > 0:
                    dims = []
                    for j, n in enumerate(value.shape):
                        name = f"{key}_dim_{j}"
                        if name not in grp.dimensions:
                            grp.createDimension(name, n)
                        dims.append(name)
                    var = grp.createVariable(key, "f4", tuple(dims), zlib=True, complevel=4)
                    var[:] = value


def nc_to_npz(nc_path: str, npz_path: str) -> None:
    """Flatten every group / variable of a NetCDF4 patch file into the .npz container (needs netCDF4)."""
    flat = {}
    with _nc().Dataset(nc_path, "r") as ds:
        for g, grp in ds.groups.items():
            for v, var in grp.variables.items():
                flat[f"{g}/{v}"] = _filled(var[:])
    np.savez_compressed(npz_path, **flat)


def npz_to_nc(npz_path: str, nc_path: str) -> None:
    z = _load_npz(npz_path)
    groups: dict = {}
    for k, a in z.items():
        if k.startswith("__attrs__/"):
            continue
        g, v = k.split("/", 1)
        groups.setdefault(g, {})[v] = a
    write_groups(nc_path, groups)


def list_patch_files(directory: str, sort: bool, exts=PATCH_EXTS) -> list:
    """File NAMES of a folder: sorted like `sorted(glob('*.nc'))` (C_30:140, C_31:139) or in os.listdir order
    (D:71, E:208 -- the order the reference's RNG draws are bound to)."""
    names = [f for f in os.listdir(directory) if f.endswith(exts)]
    return sorted(names) if sort else names


if __name__ == "__main__":
    # python -m kmsr_b200.patch_io nc2npz <in.nc|dir> <out.npz|dir>   (or npz2nc); needs the netCDF4 module
    import sys
    if len(sys.argv) != 4 or sys.argv[1] not in ("nc2npz", "npz2nc"):
        raise SystemExit("usage: python -m kmsr_b200.patch_io nc2npz|npz2nc <file or folder> <file or folder>")
    fn, ext_in, ext_out = (nc_to_npz, ".nc", ".npz") if sys.argv[1] == "nc2npz" else (npz_to_nc, ".npz", ".nc")
    src, dst = sys.argv[2], sys.argv[3]
    if os.path.isdir(src):
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            if name.endswith(ext_in):
                fn(os.path.join(src, name), os.path.join(dst, name[:-len(ext_in)] + ext_out))
    else:
        fn(src, dst)
