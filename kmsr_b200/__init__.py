"""Importable alias of the hyphen-named package directory `kernel-modeling-super-resolution_b200/`.

`import kmsr_b200` executes that directory's __init__.py under this module name, so submodules
resolve as `kmsr_b200.<name>` while the sources stay where the layout contract puts them.
"""
import importlib.util as _u
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "kernel-modeling-super-resolution_b200")
_spec = _u.spec_from_file_location("kmsr_b200", _os.path.join(_real, "__init__.py"),
                                   submodule_search_locations=[_real])
_mod = _u.module_from_spec(_spec)
_sys.modules["kmsr_b200"] = _mod
_spec.loader.exec_module(_mod)
