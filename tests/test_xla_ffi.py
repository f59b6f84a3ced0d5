"""The XLA FFI custom-call layer (include/gomelt_xla_ffi.h, gomelt_b200/csrc/xla_ffi.cc): every handler is exported, and a
C++ program that builds XLA_FFI_CallFrames by hand - as the XLA runtime does - gets the registration metadata, the error
objects and (on the GPU) bit-identical results to the direct C-ABI calls.  No jaxlib is needed (or available here): the
layer is written against the FFI *C* API, tests/ffi/ffi_callframe_test.cu plays the runtime."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _handlers():
    txt = open(os.path.join(ROOT, "include", "gomelt_xla_ffi.h")).read()
    return re.findall(r"X\((Gomelt\w+Ffi)\)", txt)


def _build(tmp_path, gm):
    exe = str(tmp_path / "ffi_callframe_test")
    inc = gm.build.xla_include_dirs()
    cmd = [gm.build.NVCC, "-std=c++17", "-O1", *gm.build.ARCH, "-I", os.path.join(ROOT, "include"), *[a for d in inc for a in ("-I", d)],
           os.path.join(ROOT, "tests", "ffi", "ffi_callframe_test.cu"), "-o", exe, gm.build.lib_path(),
           "-Xlinker", "-rpath", "-Xlinker", os.path.dirname(gm.build.lib_path())]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_every_handler_is_exported(gm):
    lib = gm.load()
    names = _handlers()
    assert len(names) == 22 and lib.gomelt_xla_ffi_available() == 1
    raw = C.CDLL(gm._lib.lib_path())
    raw.gomelt_xla_ffi_handler_count.restype = C.c_int
    raw.gomelt_xla_ffi_handler_name.restype = C.c_char_p
    assert raw.gomelt_xla_ffi_handler_count() == len(names)
    assert [raw.gomelt_xla_ffi_handler_name(i).decode() for i in range(len(names))] == names
    for n in names:
        assert hasattr(raw, n), f"{n} declared in gomelt_xla_ffi.h but not exported"


def test_call_frames_metadata_and_errors(gm, tmp_path):
    r = subprocess.run([_build(tmp_path, gm), "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_call_frames_on_the_gpu_match_the_c_abi(gm, tmp_path):
    r = subprocess.run([_build(tmp_path, gm)], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
