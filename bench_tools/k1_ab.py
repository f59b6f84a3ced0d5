"""A/B of builds of libgomelt_sm100.so on ONE GPU box: the Level-3 substep shape of bench.py (in-place state, source tables,
fused flux, clamp, Dirichlet faces left) timed through raw ctypes for every library path given on the command line.
    python bench_tools/k1_ab.py gomelt_b200/lib/libgomelt_sm100.so gomelt_b200/lib/ab/other.so ..."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gomelt_b200 import _lib, schema  # noqa: E402


def main():
    torch.cuda.set_device(0)
    P = schema.SetupProperties(bench.EXAMPLE_PROPS)
    props = _lib.make_props(P)
    ex, ey, ez = bench.L3_ELEMENTS
    nodes = (ex + 1, ey + 1, ez + 1)
    nx, ny, nz = nodes
    nn = nx * ny * nz
    grid = _lib.make_grid(nodes, (bench.L3_H,) * 3)
    blk = bench.L3Block()  # same fields as the bench
    T0, S1 = torch.as_tensor(blk.T0_host).cuda(), torch.as_tensor(blk.S1_host).cuda()
    coords = blk.coords
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    out = {}
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for path in sys.argv[1:]:
        lib = C.CDLL(os.path.abspath(path))
        lib.gomelt_level_step_f32.restype = C.c_int
        lib.gomelt_source_tables_f32.restype = C.c_int
        lib.gomelt_last_error.restype = C.c_char_p
        tx, ty, tz = (torch.empty(n, device="cuda") for n in nodes)
        coef = C.c_float(0)
        v = (C.c_float * 3)(0.3 * ex * bench.L3_H, 0.5 * ey * bench.L3_H, 0.0)
        rc = lib.gomelt_source_tables_f32(C.byref(props), C.byref(grid), C.c_void_p(coords[0].data_ptr()),
                                          C.c_void_p(coords[1].data_ptr()), C.c_void_p(coords[2].data_ptr()), C.byref(v),
                                          C.c_float(P["laser_power"]), C.c_void_p(tx.data_ptr()), C.c_void_p(ty.data_ptr()),
                                          C.c_void_p(tz.data_ptr()), C.byref(coef), stream)
        assert rc == 0, lib.gomelt_last_error()
        A, B, S = T0.clone(), T0.clone(), S1.clone()  # (faces of both buffers initialised: the step leaves them)
        a = _lib.StepArgs()
        a.grid = grid
        a.src_x, a.src_y, a.src_z, a.src_coef = tx.data_ptr(), ty.data_ptr(), tz.data_ptr(), coef.value
        a.dt, a.nz_active, a.n_substrate = bench.DT, nz, 0
        a.flags = _lib.STEP_SKIP_FACES | _lib.STEP_CLAMP | _lib.STEP_WRITE_S1 | _lib.STEP_FUSED_FLUX
        res = {}
        for mode in ("cold", "warm"):
            ts = []
            for it in range(24):
                if mode == "cold":
                    flush.zero_()
                a.T0, a.S1, a.T_out, a.S1_out = A.data_ptr(), S.data_ptr(), B.data_ptr(), S.data_ptr()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.gomelt_level_step_f32(C.byref(props), C.byref(a), stream)
                e1.record()
                assert rc == 0, lib.gomelt_last_error()
                torch.cuda.synchronize()
                if it >= 4:
                    ts.append(e0.elapsed_time(e1) * 1e3)
                A, B = B, A
            ts.sort()
            res[mode + "_us_median"] = ts[len(ts) // 2]
            res[mode + "_us_min"] = ts[0]
        res["checksum"] = float(A.double().sum())
        out[os.path.basename(path)] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
