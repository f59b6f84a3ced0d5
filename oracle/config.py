"""Oracle-wide dtype switch (test infrastructure; see oracle/__init__.py)."""
import numpy as np

# The reference never enables jax_enable_x64 -> every array is float32 / int32 (cF:15 "TFSP").
FDT = np.float32


def f(x):
    """Cast a scalar/array to the oracle float dtype (JAX weak-type promotion of Python floats)."""
    return np.asarray(x, dtype=FDT) if np.ndim(x) else FDT(x)


class use_dtype:
    """Context manager: run the oracle in another float dtype (float64 for round-off estimates)."""

    def __init__(self, dt):
        self.dt = dt

    def __enter__(self):
        global FDT
        self.prev = FDT
        FDT = self.dt
        return self

    def __exit__(self, *a):
        global FDT
        FDT = self.prev
