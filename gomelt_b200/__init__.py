"""gomelt_b200 — B200 (sm_100a) implementation of GO-MELT's multilevel explicit FE thermal
time-step behind the reference's computeFunctions-level entry points.

Sub-modules: ``_lib`` (ctypes binding of the C ABI), ``ops`` (launchers), ``build`` (nvcc),
``computeFunctions`` (the drop-in namespace), ``driver`` (the go_melt loop), ``slab`` (Level-1 z-slabs), ``dist`` (the drop-in on a slab-decomposed Level 1).
"""
from . import build, _lib, ops, schema  # noqa: F401
from ._lib import GomeltError, load  # noqa: F401

__all__ = ["build", "ops", "schema", "load", "GomeltError"]

_LAZY = ("slab", "dist", "hostpipe", "computeFunctions", "driver", "levels", "output", "toolpath")


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
