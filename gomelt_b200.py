"""Import alias: ``import gomelt_b200`` -> the package in ``go-melt_b200/`` (a directory name the
``import`` statement cannot spell).  Reach sub-modules as attributes (``gomelt_b200.ops``)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("go-melt_b200")
