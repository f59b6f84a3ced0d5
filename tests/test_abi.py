"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/gomelt_abi.h declares; argument validation returns the documented error codes."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gomelt_abi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gomelt_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol(gm):
    lib = gm.load()
    syms = _header_symbols()
    assert "gomelt_level_step_f32" in syms and len(syms) >= 6
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gomelt_abi.h but not exported"
    # and the python binding table covers the header
    assert set(syms) == set(gm._lib.SIGNATURES), set(syms) ^ set(gm._lib.SIGNATURES)
    assert lib.gomelt_abi_version() == 2


def test_bad_arguments_are_rejected_without_a_gpu(gm):
    lib = gm.load()
    props = gm._lib.Props()
    a = gm._lib.StepArgs()
    assert lib.gomelt_level_step_f32(C.byref(props), C.byref(a), None) == -1  # GOMELT_E_NULL
    assert b"NULL" in lib.gomelt_last_error()
    a.T0, a.S1, a.T_out = 256, 512, 1024  # never dereferenced on the host
    a.grid = gm._lib.make_grid((1, 5, 5), (0.1, 0.1, 0.1))
    a.dt, a.nz_active = 1e-5, 5
    assert lib.gomelt_level_step_f32(C.byref(props), C.byref(a), None) == -2  # GOMELT_E_SIZE
    a.grid = gm._lib.make_grid((5, 5, 5), (0.1, 0.1, 0.1))
    a.T_out = a.T0
    assert lib.gomelt_level_step_f32(C.byref(props), C.byref(a), None) == -3  # GOMELT_E_FLAGS
    a.T_out = 1024
    a.flags = gm._lib.STEP_WRITE_S1  # without S1_out
    assert lib.gomelt_level_step_f32(C.byref(props), C.byref(a), None) == -3


def test_no_cpu_fallback(gm):
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gm.GomeltError):
        gm.ops.diag_fp32_rate(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gomelt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_step_flag_values_match_the_header(gm):
    """The GOMELT_STEP_* bits of include/gomelt_abi.h and the STEP_* constants of the Python binding are one table."""
    txt = open(os.path.join(ROOT, "include", "gomelt_abi.h")).read()
    header = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define\s+GOMELT_STEP_([A-Z0-9_]+)\s+(0x[0-9a-fA-F]+)", txt)}
    assert {"CLAMP", "WRITE_S1", "SKIP_FACES", "GENERAL_KERNEL", "NO_COLD_PLANES"} <= set(header)
    for name, value in header.items():
        assert getattr(gm._lib, "STEP_" + name) == value, name
    assert len(set(header.values())) == len(header)  # no two flags share a bit
