"""Build libgomelt_sm100.so in-tree with nvcc (sm_100a only, -lineinfo for ncu source pages).

The built library lives in ``go-melt_b200/lib/`` (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  ``python go-melt_b200/build.py [--force] [--ptxas-v]``.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libgomelt_sm100.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    """Every .cu (sm_100a kernels + C ABI) and .cc (the compile-guarded XLA-FFI shim)."""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cc")))


def xla_include_dirs():
    """jaxlib's header tree (xla/ffi/api/ffi.h) when a jaxlib exists; [] in this image."""
    inc = os.environ.get("GOMELT_XLA_INCLUDE")
    if inc:
        return [inc]
    try:
        import importlib.util

        spec = importlib.util.find_spec("jaxlib")
        if spec and spec.submodule_search_locations:
            d = os.path.join(list(spec.submodule_search_locations)[0], "include")
            if os.path.exists(os.path.join(d, "xla", "ffi", "api", "ffi.h")):
                return [d]
    except Exception:
        pass
    return []


def _digest():
    h = hashlib.sha256()
    files = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    h.update(repr(xla_include_dirs()).encode())
    files.append(os.path.join(ROOT, "include", "gomelt_abi.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build_library(force=False, verbose=False, ptxas_v=False):
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = lib_path()
    stamp = out + ".sha256"
    dig = _digest()
    if not force and os.path.exists(out) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == dig:
                return out
    cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", *ARCH,
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, *[a for d in xla_include_dirs() for a in ("-I", d)],
           "-o", out, *sources()]
    if ptxas_v:
        cmd.insert(1, "-Xptxas=-v")
    if verbose or ptxas_v:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if ptxas_v:
        print(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return out


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv)
    print("built", p)
