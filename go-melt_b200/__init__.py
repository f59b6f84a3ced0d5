"""gomelt_b200 — B200 (sm_100a) implementation of GO-MELT's multilevel explicit FE thermal
time-step behind the reference's computeFunctions-level entry points.

The directory is ``go-melt_b200/`` (not an importable identifier); ``import gomelt_b200`` works
through the alias module at the repo root, or use ``importlib.import_module("go-melt_b200")``.
Sub-modules: ``_lib`` (ctypes binding of the C ABI), ``ops`` (launchers), ``build`` (nvcc).
"""
from . import build, _lib, ops, schema  # noqa: F401
from ._lib import GomeltError, load  # noqa: F401

__all__ = ["build", "ops", "schema", "load", "GomeltError"]

_LAZY = ("slab", "hostpipe")  # sub-modules that import torch at module level


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
