"""JAX-side registration of the XLA FFI handlers of libgomelt_sm100.so (include/gomelt_xla_ffi.h).

Nothing here runs in this image (no jax / jaxlib, and the reference pins jax 0.4.16, which predates both the typed FFI and
Blackwell): the handler layer itself is exercised by tests/ffi/ffi_callframe_test.cu, which plays the XLA runtime.  With a
Blackwell-capable jaxlib: rebuild the library against THAT jaxlib's headers (``GOMELT_XLA_INCLUDE=<jaxlib>/include python
gomelt_b200/build.py --force``), then

    import gomelt_b200.xla_ffi as gx
    gx.register_all()                      # jax.ffi.register_ffi_target("gomelt_level_step_f32", ..., platform="CUDA") ...
    T_new, = jax.ffi.ffi_call("gomelt_level_step_f32", [jax.ShapeDtypeStruct(T0.shape, T0.dtype)])(
        T0, S1, args=gx.slots(step_args, T0=1, S1=2, T_out=-1), props=gx.blob(props))

``slots`` / ``blob`` build the attributes of the calling convention: the argument struct of gomelt_abi.h as a u8 array in
which every device-pointer field holds a slot code (k > 0: operand k-1, k < 0: result -k-1, 0: NULL).
"""
import ctypes as C
import re

import numpy as np

from . import _lib


def target_name(handler):
    """GomeltLevelStepFfi -> gomelt_level_step_f32 (the name of the C-ABI entry point it forwards to)."""
    base = re.sub(r"(?<!^)(?=[A-Z])", "_", handler[len("Gomelt"):-len("Ffi")]).lower()
    base = base.replace("min_max", "minmax").replace("l3_substeps", "l3_substeps")
    return "gomelt_" + base + ("" if base == "box_copy" else "_f32")


def handlers():
    lib = C.CDLL(_lib.lib_path())
    lib.gomelt_xla_ffi_handler_count.restype = C.c_int
    lib.gomelt_xla_ffi_handler_name.restype = C.c_char_p
    return [(lib.gomelt_xla_ffi_handler_name(i).decode(), lib) for i in range(lib.gomelt_xla_ffi_handler_count())]


def register_all(platform="CUDA"):
    import jax

    out = []
    for name, lib in handlers():
        jax.ffi.register_ffi_target(target_name(name), jax.ffi.pycapsule(getattr(lib, name)), platform=platform)
        out.append(target_name(name))
    return out


def blob(struct):
    """A ctypes structure (gomelt_b200._lib.Props, StepArgs, ...) as the u8 array attribute the handlers expect."""
    return np.frombuffer(bytes(struct), dtype=np.uint8).copy()


def slots(struct, **codes):
    """``struct`` with the named pointer fields set to slot codes -> u8 array attribute.  Nested fields use ``__``
    (``L1__T0=3``); array elements an index (``src__0__coords=1``)."""
    for path, code in codes.items():
        obj, parts = struct, path.split("__")
        for p in parts[:-1]:
            obj = obj[int(p)] if p.isdigit() else getattr(obj, p)
        setattr(obj, parts[-1], C.c_void_p(int(code) & 0xFFFFFFFFFFFFFFFF))
    return blob(struct)
