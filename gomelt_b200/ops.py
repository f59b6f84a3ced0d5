"""Python launchers over the C ABI: torch CUDA tensors own the device memory, the kernels run on
torch's current stream.  One function per ``extern "C"`` entry point of include/gomelt_abi.h.
"""
import ctypes as C

from . import _lib
from ._lib import (STEP_ACCUM, STEP_BC_CONST, STEP_CLAMP, STEP_FUSED_FLUX, STEP_GENERAL_KERNEL,  # noqa: F401
                   STEP_NO_COLD_PLANES,
                   STEP_SKIP_FACES,
                   STEP_WRITE_S1, STEP_WRITE_S2)

LAUNCHES = 0  # kernels launched by the library in this process (bench.py reports the difference as gpu_launches)
GRAPH_LAUNCHES = 0  # kernels executed through CUDA-graph replays (computeFunctions.dwellRows)


def _count(n=1):
    """Refresh LAUNCHES from the library's own counter (gomelt_launch_count: every <<< >>> site counts itself, so
    calls that issue several kernels - the one-call inner scan, a Level-1 step + its constant faces - are exact)."""
    global LAUNCHES
    LAUNCHES = int(_lib.load().gomelt_launch_count())


def _chk_f32(t, n, name):
    torch = _lib.require_cuda()
    if t is None:
        return
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= n):
        raise _lib.GomeltError(f"{name}: need a contiguous float32 CUDA tensor with >= {n} elements")


def level_step(props, grid, T0, S1, T_out, dt, *, rhs=None, src=None, topflux=None, nz_active=None,
               n_substrate=0, flags=0, bc5=None, S1_out=None, S2_out=None, S2_prev=None, accum=None,
               max_accum=None, z_chunk=0, z_range=None, peer_lo=None, peer_hi=None, halo=None, bk_queue=None, bk_queue_force=False):
    """K1 (gomelt_level_step_f32): one explicit sweep of one level.  ``src`` = (tx, ty, tz, coef).
    ``peer_lo`` / ``peer_hi`` = raw device addresses (int) of the z-neighbours' ghost planes in peer-mapped
    memory.  ``halo`` = (sync, sync_lo, sync_hi, seq): the fused halo protocol of gomelt_abi.h (raw addresses of this
    rank's and the neighbours' counter blocks, sweep sequence number); without it the boundary planes are pushed after
    the step and the caller orders the sweeps."""
    lib = _lib.load()
    nn = grid.nx * grid.ny * grid.nz
    for t, name in ((T0, "T0"), (S1, "S1"), (T_out, "T_out"), (rhs, "rhs"), (S1_out, "S1_out"),
                    (accum, "accum"), (max_accum, "max_accum")):
        _chk_f32(t, nn, name)
    a = _lib.StepArgs()
    a.grid = grid
    a.T0, a.S1, a.rhs = T0.data_ptr(), S1.data_ptr(), (rhs.data_ptr() if rhs is not None else None)
    if src is not None:
        tx, ty, tz, coef = src
        a.src_x, a.src_y, a.src_z, a.src_coef = tx.data_ptr(), ty.data_ptr(), tz.data_ptr(), float(coef)
    a.topflux = topflux.data_ptr() if topflux is not None else None
    a.dt = float(dt)
    a.nz_active = int(grid.nz if nz_active is None else nz_active)
    a.n_substrate = int(n_substrate)
    a.flags = int(flags)
    if bc5 is not None:
        a.bc5 = (C.c_float * 5)(*[float(v) for v in bc5])
    a.T_out = T_out.data_ptr()
    a.S1_out = S1_out.data_ptr() if S1_out is not None else None
    a.S2_out = S2_out.data_ptr() if S2_out is not None else None
    a.S2_prev = S2_prev.data_ptr() if S2_prev is not None else None
    a.accum = accum.data_ptr() if accum is not None else None
    a.max_accum = max_accum.data_ptr() if max_accum is not None else None
    a.z_chunk = int(z_chunk)
    if z_range is not None:
        a.z_begin, a.z_end = int(z_range[0]), int(z_range[1])
    a.peer_lo, a.peer_hi = (int(peer_lo) if peer_lo else None), (int(peer_hi) if peer_hi else None)
    if halo is not None:
        sync, slo, shi, seq = halo
        a.halo_sync, a.halo_sync_lo, a.halo_sync_hi = int(sync), (int(slo) if slo else None), (int(shi) if shi else None)
        a.halo_seq = int(seq)
    if bk_queue is not None:  # int32 / uint32 scratch: the hot-plane queue of the melt-time bookkeeping (gomelt_abi.h)
        a.bk_queue, a.bk_queue_words = bk_queue.data_ptr(), int(bk_queue.numel())
        a.bk_queue_keep = 2 if bk_queue_force else 0
    _lib.check(lib.gomelt_level_step_f32(C.byref(props), C.byref(a), _lib.stream_ptr()), "gomelt_level_step_f32")
    _count()
    return T_out


def minmax(x, out=None):
    """gomelt_minmax_f32: device tensor [min, max, n_nonfinite] of a float32 field (one fused reduction)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    _chk_f32(x, x.numel(), "x")
    if out is None:
        out = torch.empty(3, device=x.device, dtype=torch.float32)
    _lib.check(lib.gomelt_minmax_f32(_lib.ptr(x), int(x.numel()), _lib.ptr(out), _lib.stream_ptr()), "gomelt_minmax_f32")
    _count()
    return out


def state_props(props, T, S1, n_substrate=0, *, S1_out=None, S2_out=None, k_out=None, rhocp_out=None):
    """computeStateProperties cF:2567-2614 (gomelt_state_props_f32)."""
    lib = _lib.load()
    nn = T.numel()
    _chk_f32(T, nn, "T")
    _chk_f32(S1, nn, "S1")
    _lib.check(lib.gomelt_state_props_f32(C.byref(props), _lib.ptr(T), _lib.ptr(S1), nn, int(n_substrate),
                                          _lib.ptr(S1_out), _lib.ptr(S2_out), _lib.ptr(k_out),
                                          _lib.ptr(rhocp_out), _lib.stream_ptr()), "gomelt_state_props_f32")
    _count()


def surface_flux(props, grid, T0, flux, nz_active=None, add=False):
    """computeConvRadBC cF:2207-2301 (gomelt_surface_flux_f32) -> flux[nx*ny]."""
    lib = _lib.load()
    _chk_f32(T0, grid.nx * grid.ny * grid.nz, "T0")
    _chk_f32(flux, grid.nx * grid.ny, "flux")
    nza = int(grid.nz if nz_active is None else nz_active)
    _lib.check(lib.gomelt_surface_flux_f32(C.byref(props), C.byref(grid), _lib.ptr(T0), nza, _lib.ptr(flux),
                                           int(bool(add)), _lib.stream_ptr()), "gomelt_surface_flux_f32")
    _count()
    return flux


def source_tables(props, grid, coords, laser_xyz, laserP, tx, ty, tz):
    """computeSourcesL3 cF:2960-3012 as rank-1 tables (gomelt_source_tables_f32) -> coef.  One launch when
    tx, ty, tz are consecutive slices of one buffer, else three."""
    lib = _lib.load()
    v = (C.c_float * 3)(float(laser_xyz[0]), float(laser_xyz[1]), float(laser_xyz[2]))
    coef = C.c_float(0.0)
    _lib.check(lib.gomelt_source_tables_f32(C.byref(props), C.byref(grid), _lib.ptr(coords[0]),
                                            _lib.ptr(coords[1]), _lib.ptr(coords[2]), C.byref(v),
                                            float(laserP), _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(tz),
                                            C.byref(coef), _lib.stream_ptr()), "gomelt_source_tables_f32")
    es = tx.element_size()
    _count(1 if (ty.data_ptr() == tx.data_ptr() + es * tx.numel() and tz.data_ptr() == ty.data_ptr() + es * ty.numel())
           else 3)
    return coef.value


def _axes(coords):
    ax = (_lib.Axis * 3)()
    for d in range(3):
        ax[d].coords = coords[d].data_ptr()
        ax[d].n = int(coords[d].numel())
    return ax


def interp(src_coords, u, tgt_coords, out, *, mode=_lib.INTERP_SET, u2=None, alpha=1.0, beta=0.0, faces_only=False,
           clamp_min=None, index_map=None, base=None, u_offset=0):
    """gomelt_interp_f32: trilinear interpolation of ``u`` (on the level with device coordinate arrays
    ``src_coords``) at the tensor-product target grid ``tgt_coords`` -> ``out`` (see gomelt_abi.h).
    ``index_map`` = (mx, my, mz, big_nx, big_ny) scatters the result through a tensor-product index set.
    ``u_offset`` = number of leading nodes of the source level that ``u`` does NOT hold (a z-slab's local array used
    with the global coordinate arrays: the caller guarantees that only nodes it holds are read)."""
    lib = _lib.load()
    a = _lib.InterpArgs()
    a.src = _axes(src_coords)
    a.u, a.u2 = u.data_ptr() - 4 * int(u_offset), (u2.data_ptr() if u2 is not None else None)
    a.alpha, a.beta = float(alpha), float(beta)
    a.tx, a.ty, a.tz = (t.data_ptr() for t in tgt_coords)
    a.ntx, a.nty, a.ntz = (int(t.numel()) for t in tgt_coords)
    a.mode, a.faces_only = int(mode), int(bool(faces_only))
    a.has_clamp, a.clamp_min = (0, 0.0) if clamp_min is None else (1, float(clamp_min))
    if index_map is not None:
        mx, my, mz, bnx, bny = index_map
        a.map_x, a.map_y, a.map_z, a.map_nx, a.map_ny = mx.data_ptr(), my.data_ptr(), mz.data_ptr(), int(bnx), int(bny)
    a.base = base.data_ptr() if base is not None else None
    a.out = out.data_ptr()
    _lib.check(lib.gomelt_interp_f32(C.byref(a), _lib.stream_ptr()), "gomelt_interp_f32")
    _count()
    return out


def _interp_args(src_coords, u, tgt_coords, out, *, mode=_lib.INTERP_SET, u2=None, alpha=1.0, beta=0.0,
                 faces_only=False, clamp_min=None):
    a = _lib.InterpArgs()
    a.src = _axes(src_coords)
    a.u, a.u2 = u.data_ptr(), (u2.data_ptr() if u2 is not None else None)
    a.alpha, a.beta = float(alpha), float(beta)
    a.tx, a.ty, a.tz = (t.data_ptr() for t in tgt_coords)
    a.ntx, a.nty, a.ntz = (int(t.numel()) for t in tgt_coords)
    a.mode, a.faces_only = int(mode), int(bool(faces_only))
    a.has_clamp, a.clamp_min = (0, 0.0) if clamp_min is None else (1, float(clamp_min))
    a.out = out.data_ptr() if out is not None else None
    return a


_FACE_SCRATCH = {}


def faces_scratch(grid, device):
    """Device scratch for the compact face interpolants of a subcycle block (2 x gomelt_faces_count floats),
    cached per grid shape."""
    import torch

    lib = _lib.load()
    n = int(lib.gomelt_faces_count(grid.nx, grid.ny, grid.nz))
    key = (n, str(device))
    if key not in _FACE_SCRATCH:
        _FACE_SCRATCH[key] = torch.empty(2 * n, device=device, dtype=torch.float32)
    return _FACE_SCRATCH[key]


def l3_substeps(props, grid, coords, rows, T_in, T_a, T_b, S1, tables, *, S1_in=None, n_substrate=0, flags=0,
                S2=None, accum=None, max_accum=None, faces=None, compact_faces=True, step_events=None):
    """gomelt_l3_substeps_f32: the inner scan of subcycleGOMELT (cF:3367-3412 / 3530-3590) as one call.
    ``rows`` = host float32 array [n, 7] of toolpath rows; ``faces`` = None or
    (parent_coords, parent_new, parent_old, fN3, clamp_min).  Returns the tensor (T_a or T_b) that holds the
    newest temperature.  Substep i writes T_a (i even) / T_b (i odd); substep 0 reads T_in (may be T_b) and
    S1_in (default S1); S1 / S2 / accum / max_accum are updated in place."""
    import numpy as np

    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n = int(rows.shape[0])
    nn = grid.nx * grid.ny * grid.nz
    for t, name in ((T_in, "T_in"), (T_a, "T_a"), (T_b, "T_b"), (S1, "S1"), (S1_in, "S1_in"), (accum, "accum"),
                    (max_accum, "max_accum")):
        _chk_f32(t, nn, name)
    _chk_f32(tables, n * (grid.nx + grid.ny + grid.nz), "tables")
    a = _lib.SubstepsArgs()
    a.grid = grid
    a.x, a.y, a.z = (c.data_ptr() for c in coords)
    a.n = n
    a.rows = rows.ctypes.data
    a.T_in, a.T_a, a.T_b, a.S1 = T_in.data_ptr(), T_a.data_ptr(), T_b.data_ptr(), S1.data_ptr()
    a.S1_in = S1_in.data_ptr() if S1_in is not None else None
    a.n_substrate, a.flags = int(n_substrate), int(flags)
    a.tables = tables.data_ptr()
    a.S2 = S2.data_ptr() if S2 is not None else None
    a.accum = accum.data_ptr() if accum is not None else None
    a.max_accum = max_accum.data_ptr() if max_accum is not None else None
    fa = None
    if faces is not None:
        pc, pnew, pold, fN, clamp_min = faces
        fa = _interp_args(pc, pnew, coords, None, u2=pold, faces_only=True, clamp_min=clamp_min)
        a.faces = C.pointer(fa)
        a.faces_n = float(fN)
        if compact_faces and pold is not None and n > 1:  # parents interpolated once per block, blended per substep
            a.faces_scratch = faces_scratch(grid, T_a.device).data_ptr()
    ev = None
    if step_events is not None:  # 2 n torch.cuda.Event(enable_timing=True): recorded around every fused level step
        torch = _lib.require_cuda()
        for e in step_events:
            e.record()           # (creates the handle; re-recorded inside the call)
        ev = (C.c_void_p * (2 * n))(*[C.c_void_p(int(e.cuda_event)) for e in step_events])
        a.step_events = C.cast(ev, C.c_void_p)
    last = C.c_void_p(0)
    a.T_last = C.pointer(last)
    _lib.check(lib.gomelt_l3_substeps_f32(C.byref(props), C.byref(a), _lib.stream_ptr()), "gomelt_l3_substeps_f32")
    _count(1 + n * (2 if faces is not None else 1) + (1 if a.faces_scratch else 0))
    return T_a if last.value == T_a.data_ptr() else T_b


def box_copy(src, dst, idx3, big_nx, big_ny, scatter):
    """gomelt_box_copy: window <-> big grid through the tensor-product index vectors ``idx3`` (int32)."""
    lib = _lib.load()
    es = src.element_size()
    if es != dst.element_size() or es not in (1, 4):
        raise _lib.GomeltError("box_copy: float32 or uint8/bool tensors of the same width")
    nx, ny, nz = (int(t.numel()) for t in idx3)
    _lib.check(lib.gomelt_box_copy(_lib.ptr(src), _lib.ptr(dst), es, _lib.ptr(idx3[0]), _lib.ptr(idx3[1]),
                                   _lib.ptr(idx3[2]), nx, ny, nz, int(big_nx), int(big_ny), int(bool(scatter)),
                                   _lib.stream_ptr()), "gomelt_box_copy")
    _count()
    return dst


def rank1(F, tx, ty, tz, coef, accumulate=True):
    lib = _lib.load()
    _lib.check(lib.gomelt_rank1_f32(_lib.ptr(F), _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(tz), int(tx.numel()),
                                    int(ty.numel()), int(tz.numel()), float(coef), int(bool(accumulate)),
                                    _lib.stream_ptr()), "gomelt_rank1_f32")
    _count()
    return F


def coarse_source_tables(props, fine_coords, parent_coords, laser_xyz, laserP, tx, ty, tz):
    """gomelt_coarse_source_tables_f32 -> 6 sqrt(3) P eta (multiply by the fine wq and use with rank1)."""
    lib = _lib.load()
    v = (C.c_float * 3)(float(laser_xyz[0]), float(laser_xyz[1]), float(laser_xyz[2]))
    coef = C.c_float(0.0)
    fa, pa = _axes(fine_coords), _axes(parent_coords)
    _lib.check(lib.gomelt_coarse_source_tables_f32(C.byref(props), C.byref(fa), C.byref(pa), C.byref(v), float(laserP),
                                                   _lib.ptr(tx), _lib.ptr(ty), _lib.ptr(tz), C.byref(coef),
                                                   _lib.stream_ptr()), "gomelt_coarse_source_tables_f32")
    _count(3)
    return coef.value


def project(fine_coords, parent_coords, A, coef, V, cells, *, mode, scale=1.0, A2=None, accumulate=True, tiled=True,
            coef_from=None):
    """gomelt_project_f32.  ``cells`` = dict(cell0, ncell, first (3 int32 device arrays), cellsum, hint, wtab, rmax, hf,
    hc).  ``tiled=False`` forces the general kernel, ``tiled="tile"`` the shared-memory tile kernel instead of the marching
    kernel that nested windows get (A/B and parity tests).  ``coef_from`` = (props, T, S1, n_substrate) evaluates the
    coefficient in the kernel (then ``coef`` is None)."""
    lib = _lib.load()
    a = _lib.ProjectArgs()
    a.fine, a.parent = _axes(fine_coords), _axes(parent_coords)
    a.A, a.A2 = A.data_ptr(), (A2.data_ptr() if A2 is not None else None)
    if coef is not None:
        a.coef = coef.data_ptr()
    else:
        cprops, cT, cS1, cnsub = coef_from
        a.coef_T, a.coef_S1, a.coef_n_substrate, a.coef_props = cT.data_ptr(), cS1.data_ptr(), int(cnsub), C.pointer(cprops)
    a.mode, a.scale = int(mode), float(scale)
    a.cell0 = (C.c_int32 * 3)(*[int(v) for v in cells["cell0"]])
    a.ncell = (C.c_int32 * 3)(*[int(v) for v in cells["ncell"]])
    a.first_x, a.first_y, a.first_z = (t.data_ptr() for t in cells["first"])
    a.elems_per_cell_hint = -int(cells["hint"]) if tiled == "tile" else int(cells["hint"])
    if tiled and cells.get("wtab") is not None:
        a.wtab_x, a.wtab_y, a.wtab_z = (t.data_ptr() for t in cells["wtab"])
        a.rmax = (C.c_int32 * 3)(*[int(v) for v in cells["rmax"]])
        a.hf = (C.c_float * 3)(*cells["hf"])
        a.hc = (C.c_float * 3)(*cells["hc"])
        a.uniform_off = (C.c_int32 * 3)(*[int(v) for v in cells.get("off", (0, 0, 0))])
    a.cellsum, a.V, a.accumulate = cells["cellsum"].data_ptr(), V.data_ptr(), int(bool(accumulate))
    _lib.check(lib.gomelt_project_f32(C.byref(a), _lib.stream_ptr()), "gomelt_project_f32")
    _count(2)
    return V


def projected_source(props, fine_coords, parent_coords, wq_fine, rows, tables, F, accumulate=False):
    """gomelt_projected_source_f32: computeLevelSource cF:2667-2730 / computeSources cF:928-988 for all laser rows at
    once (two launches).  ``rows`` = host float32 [n, 7] with the laser power in column 6."""
    import numpy as np

    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 7)
    fa, pa = _axes(fine_coords), _axes(parent_coords)
    _lib.check(lib.gomelt_projected_source_f32(C.byref(props), C.byref(fa), C.byref(pa), float(wq_fine), rows.ctypes.data,
                                               int(rows.shape[0]), _lib.ptr(tables), _lib.ptr(F), int(bool(accumulate)),
                                               _lib.stream_ptr()), "gomelt_projected_source_f32")
    _count()
    return F


def shift_window(L1_coords, T1, old_coords, Tp_old, tgt_coords, Tp_new, T_new, mid_coords=None, Tp_mid=None):
    """gomelt_shift_window_f32: the window shift of moveEverything cF:2400-2510 for one level in one launch."""
    lib = _lib.load()
    a = _lib.ShiftArgs()
    a.L1, a.T1 = _axes(L1_coords), T1.data_ptr()
    if Tp_mid is not None:
        a.mid, a.Tp_mid = _axes(mid_coords), Tp_mid.data_ptr()
    a.old, a.Tp_old = _axes(old_coords), Tp_old.data_ptr()
    a.tx, a.ty, a.tz = (t.data_ptr() for t in tgt_coords)
    a.ntx, a.nty, a.ntz = (int(t.numel()) for t in tgt_coords)
    a.Tp_new, a.T_new = Tp_new.data_ptr(), T_new.data_ptr()
    _lib.check(lib.gomelt_shift_window_f32(C.byref(a), _lib.stream_ptr()), "gomelt_shift_window_f32")
    _count()
    return Tp_new, T_new


def clamp_min(x, lo):
    lib = _lib.load()
    _lib.check(lib.gomelt_clamp_min_f32(_lib.ptr(x), int(x.numel()), float(lo), _lib.stream_ptr()), "gomelt_clamp_min_f32")
    _count()
    return x


def patch_copy(src, sdims, slo, dst, ddims, dlo, n):
    """gomelt_patch_copy_f32: box [slo, slo + n) of the (nx, ny, nz) array ``src`` -> box at ``dlo`` of ``dst``.  ``src`` /
    ``dst`` are float32 tensors or raw device addresses (a peer-mapped array of another rank)."""
    lib = _lib.load()
    v3 = lambda t: (C.c_int32 * 3)(*[int(q) for q in t])
    a = src if isinstance(src, int) else src.data_ptr()
    b = dst if isinstance(dst, int) else dst.data_ptr()
    _lib.check(lib.gomelt_patch_copy_f32(C.c_void_p(a), C.byref(v3(sdims)), C.byref(v3(slo)), C.c_void_p(b), C.byref(v3(ddims)),
                                         C.byref(v3(dlo)), C.byref(v3(n)), _lib.stream_ptr()), "gomelt_patch_copy_f32")
    _count()
    return dst


def hier_work_floats(hier, N2, N3):
    return int(_lib.load().gomelt_hier_work_floats(C.byref(hier), int(N2), int(N3)))


def subcycle(props, hier, rows, N2, N3, max_accum, accum):
    """gomelt_subcycle_f32 (subcycleGOMELT cF:3224-3632 as one native call).  Returns True when the new Level-1 field
    was left in ``hier.L1_spare``."""
    import numpy as np

    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 7)
    flag = C.c_int32(0)
    _lib.check(lib.gomelt_subcycle_f32(C.byref(props), C.byref(hier), rows.ctypes.data, int(N2), int(N3),
                                       _lib.ptr(max_accum), _lib.ptr(accum), C.byref(flag), _lib.stream_ptr()),
               "gomelt_subcycle_f32")
    _count()
    return bool(flag.value)


def step(props, hier, row, resetmask):
    """gomelt_step_f32 (stepGOMELT cF:2304-2397 as one native call)."""
    import numpy as np

    lib = _lib.load()
    row = np.ascontiguousarray(row, dtype=np.float32).reshape(7)
    flag = C.c_int32(0)
    _lib.check(lib.gomelt_step_f32(C.byref(props), C.byref(hier), row.ctypes.data, _lib.ptr(resetmask), C.byref(flag),
                                   _lib.stream_ptr()), "gomelt_step_f32")
    _count()
    return bool(flag.value)


def dwell_step(props, hier, dt):
    """gomelt_dwell_step_f32 (stepGOMELTDwellTime cF:2617-2664)."""
    lib = _lib.load()
    flag = C.c_int32(0)
    _lib.check(lib.gomelt_dwell_step_f32(C.byref(props), C.byref(hier), float(dt), C.byref(flag), _lib.stream_ptr()),
               "gomelt_dwell_step_f32")
    _count()
    return bool(flag.value)


def accum_single_step(T3, resetmask, dt, T_liquidus, accum0, max_accum0, idx3, big_nx, big_ny):
    """gomelt_accum_single_step_f32: gm:339-357 + melting_temp cF:3696-3712 on the Level-0 arrays, one launch."""
    lib = _lib.load()
    nx, ny, nz = (int(t.numel()) for t in idx3)
    _lib.check(lib.gomelt_accum_single_step_f32(_lib.ptr(T3), _lib.ptr(resetmask), float(dt), float(T_liquidus),
                                                _lib.ptr(accum0), _lib.ptr(max_accum0), _lib.ptr(idx3[0]), _lib.ptr(idx3[1]),
                                                _lib.ptr(idx3[2]), nx, ny, nz, int(big_nx), int(big_ny), _lib.stream_ptr()),
               "gomelt_accum_single_step_f32")
    _count()


def diag_fp32_rate(kind, iters=4096, blocks=148 * 8, threads=256):
    """FP32 issue-rate micro-benchmark; returns lane-ops per second (timed with CUDA events)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    sink = torch.zeros(4, device="cuda")
    ops = C.c_double(0.0)
    for _ in range(2):
        _lib.check(lib.gomelt_diag_fp32_rate(kind, iters, blocks, threads, _lib.ptr(sink), C.byref(ops),
                                             _lib.stream_ptr()), "diag")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    _lib.check(lib.gomelt_diag_fp32_rate(kind, iters, blocks, threads, _lib.ptr(sink), C.byref(ops),
                                         _lib.stream_ptr()), "diag")
    e1.record()
    torch.cuda.synchronize()
    return ops.value / (e0.elapsed_time(e1) * 1e-3)
