"""Importable alias of the ``gnn-tumor-seg_b200/`` package directory.

The package directory carries the repository's hyphenated name, which is not
a Python identifier; this shim points ``gnn_tumor_seg_b200`` at it so that
``import gnn_tumor_seg_b200`` (and its sub-modules) resolve to the single copy
of the code that lives under ``gnn-tumor-seg_b200/``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gnn-tumor-seg_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
