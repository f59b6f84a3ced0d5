"""Build libgomelt_sm100.so in-tree with nvcc (sm_100a only, -lineinfo for ncu source pages).

The built library lives in ``gomelt_b200/lib/`` (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  ``python gomelt_b200/build.py [--force] [--ptxas-v]``.
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
EXTRA_FLAGS = [f for f in os.environ.get("GOMELT_NVCC_FLAGS", "").split() if f]


def sources():
    """Every .cu (sm_100a kernels + C ABI) and .cc (the compile-guarded XLA-FFI shim)."""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cc")))


XLA_FFI_MIN = os.path.join(ROOT, "third_party", "xla_ffi_min")


def xla_include_dirs():
    """Where xla/ffi/api/c_api.h comes from: GOMELT_XLA_INCLUDE, else an installed jaxlib's header tree, else the minimal
    subset of the C API in third_party/xla_ffi_min (this image has no jaxlib)."""
    return _jaxlib_include_dirs() or [XLA_FFI_MIN]


def _jaxlib_include_dirs():
    inc = os.environ.get("GOMELT_XLA_INCLUDE")
    if inc:
        return [inc]
    try:
        import importlib.util

        spec = importlib.util.find_spec("jaxlib")
        if spec and spec.submodule_search_locations:
            d = os.path.join(list(spec.submodule_search_locations)[0], "include")
            if os.path.exists(os.path.join(d, "xla", "ffi", "api", "c_api.h")):
                return [d]
    except Exception:
        pass
    return []


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
        sorted(os.path.join(dp, f) for d in ("include", os.path.join("third_party", "xla_ffi_min"))
               for dp, _, fs in os.walk(os.path.join(ROOT, d)) for f in fs if f.endswith(".h"))


def _digest(files=None):
    """sha256 over (path relative to the repo, content) of the sources, so that the stamp is checkout-independent."""
    h = hashlib.sha256()
    h.update(repr([os.path.relpath(d, ROOT) for d in xla_include_dirs()]).encode())
    h.update(repr(EXTRA_FLAGS).encode())
    for f in (sources() if files is None else files) + _headers():
        with open(f, "rb") as fh:
            h.update(os.path.relpath(f, ROOT).encode())
            h.update(fh.read())
    return h.hexdigest()


def lib_path():
    # GOMELT_LIB_PATH: load another build of the library (development A/B of kernel variants on one GPU box)
    return os.environ.get("GOMELT_LIB_PATH") or os.path.join(LIBDIR, LIBNAME)


def _read(path):
    try:
        with open(path) as fh:
            return fh.read().strip()
    except OSError:
        return None


def build_library(force=False, verbose=False, ptxas_v=False):
    """Compile every source under csrc/ (one object per file, in parallel) and link one shared library.
    Returns its path.  Objects and stamps live in lib/obj/ (git-ignored)."""
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    out = lib_path()
    stamp = out + ".sha256"
    dig = _digest()
    if not force and os.path.exists(out) and _read(stamp) == dig:
        return out
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC, *[a for d in xla_include_dirs() for a in ("-I", d)]]
    base = [NVCC, "-c", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", *ARCH, *EXTRA_FLAGS, *inc]
    if ptxas_v:
        base.insert(1, "-Xptxas=-v")

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        d = _digest([src])
        if not force and os.path.exists(obj) and _read(obj + ".sha256") == d:
            return obj, ""
        cmd = base + ["-o", obj, src]
        if verbose or ptxas_v:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        with open(obj + ".sha256", "w") as fh:
            fh.write(d)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, sources()))
    if ptxas_v:
        for _, err in res:
            print(err)
    link = [NVCC, "-shared", *ARCH, "-o", out, *[o for o, _ in res]]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return out


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv)
    print("built", p)
