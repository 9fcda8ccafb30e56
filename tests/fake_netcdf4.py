"""A stand-in for the `netCDF4` module, for tests only: the subset of its API that the reference's readers / writers
(and patch_io's NetCDF4 branch) call -- Dataset(path, mode, format=), createGroup / groups, createDimension /
dimensions, createVariable(name, "f4", dims, zlib=, complevel=) / variables, var[:] get / set, var.filters(),
attributes by setattr -- kept in a pickle at `path` so that shutil.copy and append mode behave like files.
It checks what libnetcdf would check (dimension exists and has the array's length, no duplicate variable, file must
exist for "r" / "a") so a wrong call sequence fails here as it would against the real library."""
import os
import pickle

import numpy as np


class _Var:
    def __init__(self, name, dtype, dims, shape, zlib, complevel):
        self.__dict__["_d"] = {"name": name, "dtype": np.dtype(dtype), "dims": tuple(dims), "zlib": bool(zlib),
                               "complevel": complevel, "data": np.full(shape, np.nan, np.dtype(dtype)), "attrs": {}}

    dtype = property(lambda self: self._d["dtype"])
    dimensions = property(lambda self: self._d["dims"])
    shape = property(lambda self: self._d["data"].shape)

    def filters(self):
        return {"zlib": self._d["zlib"], "complevel": self._d["complevel"]}

    def __getitem__(self, key):
        return np.ma.masked_invalid(self._d["data"][key])          # netCDF4 returns masked arrays (C_30:54-55)

    def __setitem__(self, key, value):
        value = np.asarray(value)
        if self._d["data"][key].shape != value.shape:
            raise ValueError(f"shape mismatch writing {self._d['name']}: {self._d['data'][key].shape} vs {value.shape}")
        self._d["data"][key] = value

    def __setattr__(self, k, v):
        self._d["attrs"][k] = v

    def __getattr__(self, k):
        try:
            return self.__dict__["_d"]["attrs"][k]
        except KeyError:
            raise AttributeError(k) from None


class _Group:
    def __init__(self, parent=None):
        self.__dict__.update(groups={}, variables={}, dimensions={}, _attrs={}, _parent=parent)

    def createGroup(self, name):
        return self.groups.setdefault(name, _Group(self))

    def createDimension(self, name, size):
        if name in self.dimensions:
            raise RuntimeError(f"NetCDF: String match to name in use ({name})")
        self.dimensions[name] = int(size)

    def _dim(self, name):
        g = self
        while g is not None:
            if name in g.dimensions:
                return g.dimensions[name]
            g = g._parent
        raise ValueError(f"cannot find dimension {name} in this group or parent groups")

    def createVariable(self, name, datatype, dimensions=(), zlib=False, complevel=4, **_):
        if name in self.variables:
            raise RuntimeError(f"NetCDF: String match to name in use ({name})")
        dims = (dimensions,) if isinstance(dimensions, str) else tuple(dimensions)
        var = _Var(name, {"f4": "f4", "f8": "f8"}.get(datatype, datatype), dims, tuple(self._dim(d) for d in dims),
                   zlib, complevel)
        self.variables[name] = var
        return var

    def __setattr__(self, k, v):
        self._attrs[k] = v

    def __getattr__(self, k):
        try:
            return self.__dict__["_attrs"][k]
        except KeyError:
            raise AttributeError(k) from None


class Dataset(_Group):
    def __init__(self, path, mode="r", format="NETCDF4"):
        super().__init__()
        if mode in ("r", "a"):
            if not os.path.isfile(path):
                raise FileNotFoundError(path)
            with open(path, "rb") as f:
                magic = f.read(8)
                if magic != b"FAKENC4\n":
                    raise OSError("NetCDF: Unknown file format")
                self.__dict__.update(pickle.load(f))
        elif mode != "w":
            raise ValueError(mode)
        self.__dict__.update(_path=path, _mode=mode)

    def close(self):
        if self._mode in ("w", "a"):
            state = {k: self.__dict__[k] for k in ("groups", "variables", "dimensions", "_attrs", "_parent")}
            with open(self._path, "wb") as f:
                f.write(b"FAKENC4\n")
                pickle.dump(state, f)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
