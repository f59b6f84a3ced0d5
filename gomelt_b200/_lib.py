"""ctypes binding of libgomelt_sm100.so (the C ABI declared in include/gomelt_abi.h).

There is no CPU fallback: importing works without a GPU (so that host-side logic and symbol
checks run anywhere), but every compute entry point raises if the library is missing or CUDA
is not available.
"""
import ctypes as C
import os

from . import build as _build

c_float_p = C.POINTER(C.c_float)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("hx", C.c_float), ("hy", C.c_float), ("hz", C.c_float)]


class Props(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "k_powder", "k_bulk_a0", "k_bulk_a1", "k_fluid",
        "cp_solid_a0", "cp_solid_a1", "cp_mushy", "cp_fluid", "rho",
        "T_amb", "T_solidus", "T_liquidus", "T_boiling",
        "h_conv", "sigma_sb", "vareps", "evc", "Lev",
        "CM_coeff", "CT_coeff", "CP_coeff",
        "laser_radius", "laser_depth", "laser_eta")]


class StepArgs(C.Structure):
    _fields_ = [
        ("grid", Grid),
        ("T0", C.c_void_p), ("S1", C.c_void_p), ("rhs", C.c_void_p),
        ("src_x", C.c_void_p), ("src_y", C.c_void_p), ("src_z", C.c_void_p),
        ("src_coef", C.c_float),
        ("topflux", C.c_void_p),
        ("dt", C.c_float),
        ("nz_active", C.c_int32),
        ("n_substrate", C.c_int64),
        ("flags", C.c_int32),
        ("bc5", C.c_float * 5),
        ("T_out", C.c_void_p), ("S1_out", C.c_void_p), ("S2_out", C.c_void_p),
        ("S2_prev", C.c_void_p), ("accum", C.c_void_p), ("max_accum", C.c_void_p),
        ("z_chunk", C.c_int32),
        ("z_begin", C.c_int32), ("z_end", C.c_int32),
        ("peer_lo", C.c_void_p), ("peer_hi", C.c_void_p),
        ("halo_sync", C.c_void_p), ("halo_sync_lo", C.c_void_p), ("halo_sync_hi", C.c_void_p),
        ("halo_seq", C.c_uint32),
        ("bk_queue", C.c_void_p), ("bk_queue_words", C.c_int64), ("bk_queue_keep", C.c_int32),
    ]


class Axis(C.Structure):
    _fields_ = [("coords", C.c_void_p), ("n", C.c_int32)]


class InterpArgs(C.Structure):
    _fields_ = [
        ("src", Axis * 3),
        ("u", C.c_void_p), ("u2", C.c_void_p), ("alpha", C.c_float), ("beta", C.c_float),
        ("tx", C.c_void_p), ("ty", C.c_void_p), ("tz", C.c_void_p),
        ("ntx", C.c_int32), ("nty", C.c_int32), ("ntz", C.c_int32),
        ("mode", C.c_int32), ("faces_only", C.c_int32), ("has_clamp", C.c_int32), ("clamp_min", C.c_float),
        ("map_x", C.c_void_p), ("map_y", C.c_void_p), ("map_z", C.c_void_p),
        ("map_nx", C.c_int32), ("map_ny", C.c_int32),
        ("base", C.c_void_p), ("out", C.c_void_p),
    ]


class ProjectArgs(C.Structure):
    _fields_ = [
        ("fine", Axis * 3), ("parent", Axis * 3),
        ("A", C.c_void_p), ("A2", C.c_void_p), ("coef", C.c_void_p),
        ("mode", C.c_int32), ("scale", C.c_float),
        ("cell0", C.c_int32 * 3), ("ncell", C.c_int32 * 3),
        ("first_x", C.c_void_p), ("first_y", C.c_void_p), ("first_z", C.c_void_p),
        ("elems_per_cell_hint", C.c_int32),
        ("cellsum", C.c_void_p), ("V", C.c_void_p), ("accumulate", C.c_int32),
        ("coef_T", C.c_void_p), ("coef_S1", C.c_void_p), ("coef_n_substrate", C.c_int64),
        ("coef_props", C.POINTER(Props)),
        ("wtab_x", C.c_void_p), ("wtab_y", C.c_void_p), ("wtab_z", C.c_void_p),
        ("rmax", C.c_int32 * 3), ("hf", C.c_float * 3), ("hc", C.c_float * 3),
        ("uniform_off", C.c_int32 * 3),
    ]


class ShiftArgs(C.Structure):
    _fields_ = [
        ("L1", Axis * 3), ("T1", C.c_void_p),
        ("mid", Axis * 3), ("Tp_mid", C.c_void_p),
        ("old", Axis * 3), ("Tp_old", C.c_void_p),
        ("tx", C.c_void_p), ("ty", C.c_void_p), ("tz", C.c_void_p),
        ("ntx", C.c_int32), ("nty", C.c_int32), ("ntz", C.c_int32),
        ("Tp_new", C.c_void_p), ("T_new", C.c_void_p),
    ]


class Level(C.Structure):
    _fields_ = [
        ("grid", Grid),
        ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
        ("T0", C.c_void_p), ("S1", C.c_void_p), ("Tprime0", C.c_void_p), ("S2", C.c_void_p),
        ("n_substrate", C.c_int64),
    ]


class Pair(C.Structure):
    _fields_ = [
        ("cell0", C.c_int32 * 3), ("ncell", C.c_int32 * 3),
        ("first_x", C.c_void_p), ("first_y", C.c_void_p), ("first_z", C.c_void_p),
        ("elems_per_cell_hint", C.c_int32),
        ("wtab_x", C.c_void_p), ("wtab_y", C.c_void_p), ("wtab_z", C.c_void_p),
        ("rmax", C.c_int32 * 3),
        ("uniform_off", C.c_int32 * 3),
    ]


class Overlap(C.Structure):
    _fields_ = [
        ("ix", C.c_void_p), ("iy", C.c_void_p), ("iz", C.c_void_p),
        ("cx", C.c_void_p), ("cy", C.c_void_p), ("cz", C.c_void_p),
        ("n", C.c_int32 * 3),
    ]


class Hier(C.Structure):
    _fields_ = [
        ("L1", Level), ("L2", Level), ("L3", Level),
        ("L2L1", Pair), ("L3L1", Pair), ("L3L2", Pair),
        ("ov2", Overlap), ("ov3", Overlap),
        ("L0_S1", C.c_void_p), ("L0_S2", C.c_void_p),
        ("L0_nx", C.c_int32), ("L0_ny", C.c_int32), ("L0_nz", C.c_int32),
        ("l0_ix", C.c_void_p), ("l0_iy", C.c_void_p), ("l0_iz", C.c_void_p),
        ("l0p_ix", C.c_void_p), ("l0p_iy", C.c_void_p), ("l0p_iz", C.c_void_p), ("l0p_n", C.c_int32 * 3),
        ("bc5", C.c_float * 5),
        ("nz_active_L1", C.c_int32),
        ("L1_spare", C.c_void_p),
        ("work", C.c_void_p), ("work_floats", C.c_int64),
        ("l1_solve", C.c_void_p), ("l1_user", C.c_void_p),
    ]


class SubstepsArgs(C.Structure):
    _fields_ = [
        ("grid", Grid),
        ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
        ("n", C.c_int32),
        ("rows", C.c_void_p),
        ("T_in", C.c_void_p), ("T_a", C.c_void_p), ("T_b", C.c_void_p),
        ("S1_in", C.c_void_p), ("S1", C.c_void_p),
        ("n_substrate", C.c_int64),
        ("flags", C.c_int32),
        ("tables", C.c_void_p),
        ("S2", C.c_void_p), ("accum", C.c_void_p), ("max_accum", C.c_void_p),
        ("faces", C.POINTER(InterpArgs)),
        ("faces_n", C.c_float),
        ("T_last", C.POINTER(C.c_void_p)),
        ("bk_queue", C.c_void_p), ("bk_queue_words", C.c_int64),
        ("step_events", C.c_void_p),
        ("faces_scratch", C.c_void_p),
    ]


MAX_SUBSTEPS = 64
INTERP_SET, INTERP_ADD, INTERP_RSUB = 0, 1, 2
STEP_CLAMP, STEP_WRITE_S1, STEP_WRITE_S2 = 0x01, 0x02, 0x04
STEP_BC_CONST, STEP_SKIP_FACES, STEP_ACCUM = 0x08, 0x10, 0x20
STEP_FUSED_FLUX = 0x40
STEP_GENERAL_KERNEL = 0x80
STEP_NO_COLD_PLANES = 0x100
L1_CLAMP_AFTER = 0x10000
L1_SOLVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32)

# name -> (restype, argtypes); kept in one table so tests can check every header symbol loads
SIGNATURES = {
    "gomelt_abi_version": (C.c_int, []),
    "gomelt_launch_count": (C.c_longlong, []),
    "gomelt_halo_sync_words": (C.c_longlong, []),
    "gomelt_minmax_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "gomelt_last_error": (C.c_char_p, []),
    "gomelt_xla_ffi_available": (C.c_int, []),
    "gomelt_level_step_f32": (C.c_int, [C.POINTER(Props), C.POINTER(StepArgs), C.c_void_p]),
    "gomelt_state_props_f32": (C.c_int, [C.POINTER(Props), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gomelt_surface_flux_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Grid), C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_int32, C.c_void_p]),
    "gomelt_source_tables_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Grid), C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.POINTER(C.c_float * 3), C.c_float, C.c_void_p,
                                           C.c_void_p, C.c_void_p, c_float_p, C.c_void_p]),
    "gomelt_source_tables_batch_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Grid), C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, c_float_p,
                                                 C.c_void_p]),
    "gomelt_l3_substeps_f32": (C.c_int, [C.POINTER(Props), C.POINTER(SubstepsArgs), C.c_void_p]),
    "gomelt_interp_f32": (C.c_int, [C.POINTER(InterpArgs), C.c_void_p]),
    "gomelt_faces_count": (C.c_longlong, [C.c_int32, C.c_int32, C.c_int32]),
    "gomelt_faces_gather_f32": (C.c_int, [C.POINTER(InterpArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    "gomelt_faces_blend_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                         C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "gomelt_box_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "gomelt_rank1_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_float, C.c_int32, C.c_void_p]),
    "gomelt_coarse_source_tables_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Axis * 3), C.POINTER(Axis * 3),
                                                  C.POINTER(C.c_float * 3), C.c_float, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, c_float_p, C.c_void_p]),
    "gomelt_project_f32": (C.c_int, [C.POINTER(ProjectArgs), C.c_void_p]),
    "gomelt_projected_source_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Axis * 3), C.POINTER(Axis * 3), C.c_float,
                                              C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "gomelt_shift_window_f32": (C.c_int, [C.POINTER(ShiftArgs), C.c_void_p]),
    "gomelt_clamp_min_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "gomelt_patch_copy_f32": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.c_void_p,
                                        C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3),
                                        C.c_void_p]),
    "gomelt_hier_work_floats": (C.c_longlong, [C.POINTER(Hier), C.c_int32, C.c_int32]),
    "gomelt_subcycle_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Hier), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "gomelt_step_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Hier), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
                                  C.c_void_p]),
    "gomelt_dwell_step_f32": (C.c_int, [C.POINTER(Props), C.POINTER(Hier), C.c_float, C.POINTER(C.c_int32), C.c_void_p]),
    "gomelt_accum_single_step_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_void_p]),
    "gomelt_diag_fp32_rate": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.POINTER(C.c_double), C.c_void_p]),
}

_LIB = None


class GomeltError(RuntimeError):
    pass


def lib_path():
    return _build.lib_path()


def load(build_if_missing=False):
    """dlopen the library and attach signatures.  Raises GomeltError when it is not there."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if build_if_missing:
            _build.build_library()
        else:
            raise GomeltError(
                f"{path} is missing: run `python gomelt_b200/build.py` (or __graft_entry__.build()); "
                "there is no CPU fallback for the GO-MELT step")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().gomelt_last_error()
        raise GomeltError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


_TORCH_CUDA = None


def require_cuda():
    """torch, after checking ONCE that a CUDA device exists (torch.cuda.is_available() costs microseconds per
    call and the steppers ask ~100 times per toolpath row)."""
    global _TORCH_CUDA
    if _TORCH_CUDA is None:
        import torch

        if not torch.cuda.is_available():
            raise GomeltError("CUDA device required: the GO-MELT step has no CPU path in this package")
        _TORCH_CUDA = torch
    return _TORCH_CUDA


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    """torch's current CUDA stream of the current device as a raw cudaStream_t."""
    torch = require_cuda()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_props(P):
    """Properties dict (SetupProperties names, cF:267-345) -> gomelt_props_t, float32 like the
    scalars the reference's jitted functions trace."""
    p = Props()
    p.k_powder = P["k_powder"]
    p.k_bulk_a0 = P["k_bulk_coeff_a0"]
    p.k_bulk_a1 = P["k_bulk_coeff_a1"]
    p.k_fluid = P["k_fluid_coeff_a0"]
    p.cp_solid_a0 = P["cp_solid_coeff_a0"]
    p.cp_solid_a1 = P["cp_solid_coeff_a1"]
    p.cp_mushy = P["cp_mushy"]
    p.cp_fluid = P["cp_fluid"]
    p.rho = P["rho"]
    p.T_amb = P["T_amb"]
    p.T_solidus = P["T_solidus"]
    p.T_liquidus = P["T_liquidus"]
    p.T_boiling = P["T_boiling"]
    p.h_conv = P["h_conv"]
    p.sigma_sb = P["sigma_sb"]
    p.vareps = P["vareps"]
    p.evc = P["evc"]
    p.Lev = P["Lev"]
    p.CM_coeff = P["CM_coeff"]
    p.CT_coeff = P["CT_coeff"]
    p.CP_coeff = P["CP_coeff"]
    p.laser_radius = P["laser_radius"]
    p.laser_depth = P["laser_depth"]
    p.laser_eta = P["laser_eta"]
    return p


def make_grid(nodes, h):
    g = Grid()
    g.nx, g.ny, g.nz = int(nodes[0]), int(nodes[1]), int(nodes[2])
    g.hx, g.hy, g.hz = float(h[0]), float(h[1]), float(h[2])
    return g
