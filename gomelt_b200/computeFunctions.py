"""Drop-in ``computeFunctions`` namespace of the B200 build: the names the reference driver reaches
through ``from computeFunctions import *`` (gm:10), with the same positional arguments,
running on ``cuda`` through the C ABI of ``libgomelt_sm100.so``.

State lives on the device: ``Levels[i]["T0" | "Tprime0" | "S1" | "S2"]`` are torch CUDA tensors
(float32, S2 bool); geometry (``node_coords``, overlap index sets, bounds) is host NumPy.  Any field
may also be handed in as a NumPy array (it is uploaded), which is how the parity tests drive it.

What differs from the reference *by construction* (results agree to float32 rounding):
  * ``Shapes`` / ``LInterp`` are small descriptors, not (n,8) / (nef,8,8) operator arrays: every
    transfer weight is recomputed in registers (csrc/k_transfer.cu);
  * k and rho*cp are evaluated inside the fused level step from (T0, S1) - the arrays the reference
    passes around as ``Lk`` / ``Lrhocp`` are only materialised where a correction term integrates them;
  * the corrector sweep of Level 3 in ``stepGOMELT`` re-uses the predictor interior (identical inputs,
    cF:2199-2201 vs 2355/2375) and only re-applies the Dirichlet faces;
  * the subcycle histories ``L2all, L3all, L3pall`` (returned and dropped by the driver, gm:437) are
    returned as ``None``.
There is no CPU path: every entry point raises ``GomeltError`` without the library or a GPU.
"""
import copy  # noqa: F401  (re-exported: gm:200 uses ``copy`` through the star import)
import math

import numpy as np

from . import _lib, levels, ops, schema
from .schema import SetupNonmesh, SetupProperties, getStaticSubcycle  # noqa: F401
from .toolpath import count_lines, parsingGcode  # noqa: F401
from .output import saveFinalResult, saveResult, saveResults, saveResultsFinal, saveState  # noqa: F401

F32 = np.float32


def _torch():
    return _lib.require_cuda()


def _dev(x, dtype=None):
    """Field -> contiguous CUDA tensor (uploads NumPy input)."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x))).cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _f(x):
    return _dev(x, _torch().float32)


def _host(x):
    torch = _torch()
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _CoordCache:
    """Device copies of 1-D coordinate / index arrays, keyed by content (they are tiny)."""

    def __init__(self):
        self.store = {}

    def get(self, arr, dtype):
        a = np.ascontiguousarray(np.asarray(_host(arr)).astype(dtype))
        key = (a.dtype.str, a.tobytes())
        t = self.store.get(key)
        if t is None:
            if len(self.store) > 4096:
                self.store.clear()
            t = _torch().as_tensor(a).cuda()
            self.store[key] = t
        return t


_CACHE = _CoordCache()


def _coords(c3):
    return [_CACHE.get(c, np.float32) for c in c3]


def _index3(i3):
    return [_CACHE.get(i, np.int32) for i in i3]


def _props(properties):
    p = properties.get("_gomelt_props") if isinstance(properties, dict) else None
    if p is None:
        p = _lib.make_props(properties)
        if isinstance(properties, dict):
            properties["_gomelt_props"] = p
    return p


def _grid(L):
    return _lib.make_grid(L["nodes"], L["h"])


# ----------------------------------------------------------------------------------------------
# setup (host geometry + device fields)
# ----------------------------------------------------------------------------------------------
def SetupLevels(solver_input, properties):
    """cF:115-264.  Unused reference fields (T, k, rhocp, Tprime: cF:150-157, 172) are not allocated."""
    torch = _torch()
    L = levels.build_levels(solver_input, properties)
    for i in (1, 2, 3):
        nn = L[i]["nn"]
        L[i]["T0"] = torch.full((nn,), float(F32(properties["T_amb"])), device="cuda", dtype=torch.float32)
        L[i]["S1"] = torch.zeros(nn, device="cuda", dtype=torch.float32)
        L[i]["S2"] = torch.zeros(nn, device="cuda", dtype=torch.bool)
    L[1]["S1_storage"] = torch.zeros((L[1]["n_S1_storage"], L[1]["nn"]), device="cuda", dtype=torch.float32)
    for i in (2, 3):
        L[i]["Tprime0"] = torch.zeros(L[i]["nn"], device="cuda", dtype=torch.float32)
    L[0]["S1"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.float32)
    L[0]["S2"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.bool)
    return L


getStaticNodesAndElements = levels.static_sizes
calcStaticTmpNodesAndElements = levels.active_sizes
getSubstrateNodes = levels.substrate_counts


def getOverlapRegion(node_coords, nx, ny):
    return levels.overlap_ids(node_coords, nx, ny)


# ----------------------------------------------------------------------------------------------
# interpolation entry points
# ----------------------------------------------------------------------------------------------
def interpolatePointsMatrix(Level, node_coords_new):
    """cF:1028-1107.  The reference returns (n,8) weights + indices; here a descriptor: the weights are
    recomputed in registers whenever the pair is used (``LInterp`` is opaque to the driver, gm:74-75)."""
    return {"src_coords": [np.array(_host(c), F32) for c in Level["node_coords"]],
            "tgt_coords": [np.array(_host(c), F32) for c in node_coords_new]}


def interpolatePoints(Level, u, node_coords_new):
    """cF:1131-1210 -> CUDA tensor of length nx'*ny'*nz'."""
    torch = _torch()
    tgt = _coords(node_coords_new)
    out = torch.empty(tgt[0].numel() * tgt[1].numel() * tgt[2].numel(), device="cuda", dtype=torch.float32)
    return ops.interp(_coords(Level["node_coords"]), _f(u), tgt, out)


def interpolate_w_matrix(C2F, T):
    """cF:1110-1128 on a descriptor from interpolatePointsMatrix."""
    torch = _torch()
    tgt = _coords(C2F["tgt_coords"])
    out = torch.empty(tgt[0].numel() * tgt[1].numel() * tgt[2].numel(), device="cuda", dtype=torch.float32)
    return ops.interp(_coords(C2F["src_coords"]), _f(T), tgt, out)


# ----------------------------------------------------------------------------------------------
# per-level pieces
# ----------------------------------------------------------------------------------------------
def computeStateProperties(T, S1, properties, n_substrate):
    """cF:2567-2614 -> (S1 f32, S2 bool, k, rhocp) as CUDA tensors."""
    torch = _torch()
    T, S1 = _f(T), _f(S1)
    S1o, k, rc = torch.empty_like(T), torch.empty_like(T), torch.empty_like(T)
    S2o = torch.empty(T.numel(), device="cuda", dtype=torch.bool)
    ops.state_props(_props(properties), T, S1, int(n_substrate), S1_out=S1o, S2_out=S2o, k_out=k, rhocp_out=rc)
    return S1o, S2o, k, rc


def _surface_flux(L, T, nz_active, properties):
    """computeConvRadBC cF:2207-2301 -> [nx*ny] load of the top active plane."""
    torch = _torch()
    flux = torch.empty(L["nodes"][0] * L["nodes"][1], device="cuda", dtype=torch.float32)
    return ops.surface_flux(_props(properties), _grid(L), T, flux, nz_active=nz_active)


def computeConvRadBC(Level, LevelT0, ne, nn, properties, F):
    """cF:2207-2301 with the reference's signature: convection + radiation + evaporation load of the top face of the
    elements [ne - ne_x*ne_y, ne) added to ``F`` (exact-libm stand-alone kernel; the steppers use K1's fused
    epilogue).  Returns a new [nn] CUDA tensor."""
    nx, ny = int(Level["nodes"][0]), int(Level["nodes"][1])
    nz_active = int(ne) // ((nx - 1) * (ny - 1)) + 1
    out = _f(F).clone()
    top = out[(nz_active - 1) * nx * ny: nz_active * nx * ny]
    ops.surface_flux(_props(properties), _grid(Level), _f(LevelT0), top, nz_active=nz_active, add=True)
    return out


def _nz_active(L, tmp_nn):
    return int(tmp_nn) // (L["nodes"][0] * L["nodes"][1])


def _l3_source(L3, v, properties, laserP):
    """computeSourcesL3 cF:2960-3012 as rank-1 tables (tx, ty, tz, coef)."""
    torch = _torch()
    nx, ny, nz = L3["nodes"]
    tx, ty, tz = (torch.empty(n, device="cuda", dtype=torch.float32) for n in (nx, ny, nz))
    coef = ops.source_tables(_props(properties), _grid(L3), _coords(L3["node_coords"]), _host(v)[:3], float(laserP),
                             tx, ty, tz)
    return (tx, ty, tz, coef)


def _projected_source(L3, parent, rows, powers, properties, F=None):
    """computeSources cF:928-988 (one row) / computeLevelSource cF:2667-2730 (mean over rows): the laser
    source integrated at Level-3 Gauss points, projected on ``parent`` -> [nn_parent] load vector."""
    torch = _torch()
    nx, ny, nz = parent["nodes"]
    if F is None:
        F = torch.zeros(nx * ny * nz, device="cuda", dtype=torch.float32)
    tx, ty, tz = (torch.empty(n, device="cuda", dtype=torch.float32) for n in (nx, ny, nz))
    h3 = L3["h"]
    wq = F32(F32(F32(h3[0]) * F32(h3[1])) * F32(h3[2])) * F32(0.125)
    n = len(powers)
    fine, par = _coords(L3["node_coords"]), _coords(parent["node_coords"])
    for r in range(n):
        pc = ops.coarse_source_tables(_props(properties), fine, par, rows[r][:3], float(powers[r]), tx, ty, tz)
        ops.rank1(F, tx, ty, tz, float(F32(pc) * wq) / n, accumulate=True)
    return F


_PAIR_CACHE = {}


def _pair_cells(fine, parent):
    """Group the fine elements by parent cell (the role of Shapes[.][2] / the node set of Gauss point 0,
    cF:1351): per axis, parent cell of each fine element by the reference's floor rule.  Memoised by the
    content of the six coordinate arrays: the windows revisit the same integer shifts row after row."""
    key = tuple(np.asarray(L["node_coords"][d], F32).tobytes() for L in (fine, parent) for d in range(3))
    hit = _PAIR_CACHE.get(key)
    if hit is None:
        if len(_PAIR_CACHE) > 512:
            _PAIR_CACHE.clear()
        hit = _PAIR_CACHE[key] = _pair_cells_build(fine, parent)
    return hit


def _pair_cells_build(fine, parent):
    torch = _torch()
    g = F32(0.57735026918962576)
    lo, hi = F32(0.5) * (F32(1) + g), F32(0.5) * (F32(1) - g)
    cell0, ncell, first, hint = [], [], [], 1
    for d in range(3):
        xf = np.asarray(fine["node_coords"][d], F32)
        xc = np.asarray(parent["node_coords"][d], F32)
        xq0 = lo * xf[:-1] + hi * xf[1:]
        hc = xc[1] - xc[0]
        ec = np.clip(np.floor((xq0 - xc[0]) / hc).astype(np.int64), 0, xc.size - 2)
        c0, c1 = int(ec[0]), int(ec[-1])
        cell0.append(c0)
        ncell.append(c1 - c0 + 1)
        fs = np.searchsorted(ec, np.arange(c0, c1 + 2), side="left").astype(np.int32)
        first.append(_CACHE.get(fs, np.int32))
        hint *= int(np.diff(fs).max())
    cellsum = torch.empty(ncell[0] * ncell[1] * ncell[2] * 8, device="cuda", dtype=torch.float32)
    return {"cell0": cell0, "ncell": ncell, "first": first, "hint": hint, "cellsum": cellsum,
            "fine": _coords(fine["node_coords"]), "parent": _coords(parent["node_coords"])}


def _project(cells, A, coef, V, mode, scale=1.0, A2=None):
    return ops.project(cells["fine"], cells["parent"], A, coef, V, cells, mode=mode, scale=scale, A2=A2,
                       accumulate=True)


def _zeros_like_level(L):
    return _torch().zeros(L["nn"], device="cuda", dtype=_torch().float32)


def _faces_from_parent(parent, Tparent, child, Tchild, T_amb=None, blend=None):
    """assignBCsFine cF:1598-1620 (+ the following max(T_amb, .)): the 5 Dirichlet faces of ``Tchild`` <-
    parent field interpolated at the child's nodes.  ``blend`` = (alpha, beta, Tparent_old)."""
    kw = {}
    if blend is not None:
        kw = dict(alpha=blend[0], beta=blend[1], u2=blend[2])
    return ops.interp(_coords(parent["node_coords"]), Tparent, _coords(child["node_coords"]), Tchild,
                      faces_only=True, clamp_min=T_amb, **kw)


def _bc5(L1):
    c = L1["conditions"]
    return [c["y"][0], c["y"][1], c["x"][0], c["x"][1], c["z"][0]]


def _solve_L1(Levels, T0, S1, rhs, tmp_ne_nn, dt, properties, n_sub, clamp=True):
    """computeConvRadBC + solveMatrixFreeFE + substitute_Tbar + assignBCs (+ clamp) on Level 1
    (cF:2207-2301, 2172-2185, 2813-2854); the surface load of the top active plane is evaluated inside K1."""
    L1 = Levels[1]
    out = _torch().empty_like(T0)
    flags = ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX | (ops.STEP_CLAMP if clamp else 0)
    return ops.level_step(_props(properties), _grid(L1), T0, S1, out, float(dt), rhs=rhs,
                          nz_active=_nz_active(L1, tmp_ne_nn[1]), n_substrate=int(n_sub), flags=flags,
                          bc5=_bc5(L1))


def _solve_child(L, T0, S1, rhs, src, dt, properties, n_sub, **kw):
    """computeConvRadBC + solveMatrixFreeFE on a window level (surface load fused into K1); the 5 Dirichlet
    faces are left for _faces_from_parent."""
    out = _torch().empty_like(T0)
    flags = ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_FUSED_FLUX | kw.pop("flags", 0)
    return ops.level_step(_props(properties), _grid(L), T0, S1, out, float(dt), rhs=rhs, src=src,
                          n_substrate=int(n_sub), flags=flags, **kw)


def getNewTprime(Fine, FineT0, CoarseT, Coarse, C2F=None):
    """cF:2060-2099: inject the fine solution into the parent's overlap nodes (in place on ``CoarseT``),
    then T' = T_fine - I(parent).  Returns (Tprime, CoarseT)."""
    torch = _torch()
    FineT0, CoarseT = _f(FineT0), _f(CoarseT)
    fc, cc = _coords(Fine["node_coords"]), _coords(Coarse["node_coords"])
    ops.interp(fc, FineT0, _coords(Fine["overlapCoords"]), CoarseT,
               index_map=(*_index3(Fine["overlapNodes"]), Coarse["nodes"][0], Coarse["nodes"][1]))
    Tprime = torch.empty_like(FineT0)
    ops.interp(cc, CoarseT, fc, Tprime, mode=_lib.INTERP_RSUB, base=FineT0)
    return Tprime, CoarseT


def getBothNewTprimes(Levels, FineT, MesoT, M2F, CoarseT, C2M):
    """cF:2102-2132."""
    lTp, mT0 = getNewTprime(Levels[3], FineT, MesoT, Levels[2])
    mTp, uT0 = getNewTprime(Levels[2], mT0, CoarseT, Levels[1])
    return lTp, mTp, mT0, uT0


def _push_S1_to_L1(Levels, substrate):
    """cF:2546-2556 / 3272-3278: Level-2 S1 -> Level-1 overlap nodes, substrate planes -> 1."""
    L1, L2 = Levels[1], Levels[2]
    L1["S1"] = _f(L1["S1"])
    ops.interp(_coords(L2["node_coords"]), _f(L2["S1"]), _coords(L2["overlapCoords"]), L1["S1"],
               index_map=(*_index3(L2["overlapNodes"]), L1["nodes"][0], L1["nodes"][1]))
    L1["S1"][: int(substrate[1])] = 1.0


def _scatter_L0(Levels):
    """cF:2390-2392 / 3628-3630."""
    L0, L3 = Levels[0], Levels[3]
    L0["S1"], L0["S2"] = _f(L0["S1"]), _dev(L0["S2"], _torch().bool)
    idx3 = _index3(L0["overlapNodes"])
    ops.box_copy(_f(L3["S1"]), L0["S1"], idx3, L0["nodes"][0], L0["nodes"][1], scatter=True)
    L0["S2"].zero_()
    ops.box_copy(_dev(L3["S2"], _torch().bool), L0["S2"], idx3, L0["nodes"][0], L0["nodes"][1], scatter=True)


def _ensure_fields(Levels):
    torch = _torch()
    for i in (1, 2, 3):
        Levels[i]["T0"], Levels[i]["S1"] = _f(Levels[i]["T0"]), _f(Levels[i]["S1"])
    for i in (2, 3):
        Levels[i]["Tprime0"] = _f(Levels[i]["Tprime0"])
    Levels[3]["S2"] = _dev(Levels[3]["S2"], torch.bool)


# ----------------------------------------------------------------------------------------------
# step orchestrators
# ----------------------------------------------------------------------------------------------
def stepGOMELT(Levels, ne_nn, tmp_ne_nn, Shapes, LInterp, v, properties, dt, laserP, substrate):
    """cF:2304-2397: one single-step predictor / corrector update of Levels 1-3."""
    torch = _torch()
    _ensure_fields(Levels)
    L1, L2, L3 = Levels[1], Levels[2], Levels[3]
    T_amb = float(F32(properties["T_amb"]))
    v = _host(v)
    dt, laserP = float(_host(dt)), float(_host(laserP))
    preS2 = L3["S2"]
    # updateStateProperties cF:2513-2564
    L3["S1"], L3["S2"], k3, rc3 = computeStateProperties(L3["T0"], L3["S1"], properties, substrate[3])
    L2["S1"], _, k2, rc2 = computeStateProperties(L2["T0"], L2["S1"], properties, substrate[2])
    _push_S1_to_L1(Levels, substrate)
    # loads: laser source on Level 3 (rank-1 tables) and its projections on Levels 1-2; the surface fluxes
    # (computeConvRadBC cF:2348-2350) are evaluated inside each level step from that level's T0
    src3 = _l3_source(L3, v, properties, laserP)
    F1 = _projected_source(L3, L1, [v], [laserP], properties)
    F2 = _projected_source(L3, L2, [v], [laserP], properties)
    # computeCoarseTprimeTerm_jax cF:1477-1565
    Vcu, Vmu = _zeros_like_level(L1), _zeros_like_level(L2)
    _project(Shapes["L3L1"], L3["Tprime0"], k3, Vcu, mode=0)
    _project(Shapes["L2L1"], L2["Tprime0"], k2, Vcu, mode=0)
    _project(Shapes["L3L2"], L3["Tprime0"], k3, Vmu, mode=0)

    def solutions(Vc, Vm, L3_interior=None):
        T1 = _solve_L1(Levels, L1["T0"], L1["S1"], F1 + Vc, tmp_ne_nn, dt, properties, substrate[1])
        T2 = _solve_child(L2, L2["T0"], L2["S1"], F2 + Vm, None, dt, properties, substrate[2])
        _faces_from_parent(L1, T1, L2, T2, T_amb)
        if L3_interior is None:
            T3 = _solve_child(L3, L3["T0"], L3["S1"], None, src3, dt, properties, substrate[3])
        else:
            T3 = L3_interior  # same T0, F, k, rho*cp, Corr = 0: only the faces change (cF:2199-2201)
        _faces_from_parent(L2, T2, L3, T3, T_amb)
        return T1, T2, T3

    T1, T2, T3 = solutions(Vcu, Vmu)
    L3Tp, L2Tp, T2, T1 = getBothNewTprimes(Levels, T3, T2, None, T1, None)
    # computeCoarseTprimeMassTerm_jax cF:1396-1474
    _project(Shapes["L3L1"], L3Tp, rc3, Vcu, mode=1, scale=1.0 / F32(dt), A2=L3["Tprime0"])
    _project(Shapes["L2L1"], L2Tp, rc2, Vcu, mode=1, scale=1.0 / F32(dt), A2=L2["Tprime0"])
    _project(Shapes["L3L2"], L3Tp, rc3, Vmu, mode=1, scale=1.0 / F32(dt), A2=L3["Tprime0"])
    T1, T2, T3 = solutions(Vcu, Vmu, L3_interior=T3)
    L3["T0"] = T3
    L3["Tprime0"], L2["Tprime0"], L2["T0"], L1["T0"] = getBothNewTprimes(Levels, L3["T0"], T2, None, T1, None)
    _scatter_L0(Levels)
    resetmask = torch.logical_and(torch.logical_not(_dev(preS2, torch.bool)), L3["S2"])
    return Levels, resetmask


def stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, properties, dt, substrate):
    """cF:2617-2664: Level 1 only, no clamp."""
    L1 = Levels[1]
    L1["T0"], L1["S1"] = _f(L1["T0"]), _f(L1["S1"])
    L1["T0"] = _solve_L1(Levels, L1["T0"], L1["S1"], None, tmp_ne_nn, float(_host(dt)), properties,
                         substrate[1], clamp=False)
    return Levels


def subcycleGOMELT(Levels, ne_nn, Shapes, substrate, LInterp, tmp_ne_nn, laser_position, properties, laserP,
                   subcycle, max_accum_L3, accum_L3):
    """cF:3224-3632: Level 1 once, Level 2 x N2, Level 3 x N2*N3, predictor pass then corrector pass."""
    torch = _torch()
    _ensure_fields(Levels)
    L1, L2, L3 = Levels[1], Levels[2], Levels[3]
    T_amb = float(F32(properties["T_amb"]))
    rows = np.asarray(_host(laser_position), F32)
    P = np.asarray(_host(laserP), F32)
    N2, N3 = int(subcycle[0]), int(subcycle[1])
    fN2, fN3 = F32(subcycle[3]), F32(subcycle[4])

    _, _, k3_L1, rc3_L1 = computeStateProperties(L3["T0"], L3["S1"], properties, substrate[3])
    _, _, k2_L1, rc2_L1 = computeStateProperties(L2["T0"], L2["S1"], properties, substrate[2])
    _push_S1_to_L1(Levels, substrate)
    dt_all = float(rows[:, 5].sum(dtype=F32))
    F1 = _projected_source(L3, L1, rows, P, properties)
    V1 = _zeros_like_level(L1)
    _project(Shapes["L3L1"], L3["Tprime0"], k3_L1, V1, mode=0)
    _project(Shapes["L2L1"], L2["Tprime0"], k2_L1, V1, mode=0)
    L1T = _solve_L1(Levels, L1["T0"], L1["S1"], F1 + V1, tmp_ne_nn, dt_all, properties, substrate[1])

    def L2_common(T2, S12, T3, Tp3, S13, isub):
        a2 = F32(isub + 1) / fN2
        b2 = F32(1) - a2
        sl = slice(isub * N3, (isub + 1) * N3)
        _, _, k3, rc3 = computeStateProperties(T3, S13, properties, substrate[3])
        S12n, _, k2, rc2 = computeStateProperties(T2, S12, properties, substrate[2])
        F2 = _projected_source(L3, L2, rows[sl], P[sl], properties)
        V2 = _zeros_like_level(L2)
        _project(Shapes["L3L2"], Tp3, k3, V2, mode=0)
        dt2 = float(rows[sl, 5].sum(dtype=F32))
        return float(a2), float(b2), rc3, S12n, F2, V2, dt2

    def solve_L2(T2, S12, F2, V2, dt2, a2, b2, L1new):
        T2n = _solve_child(L2, T2, S12, F2 + V2, None, dt2, properties, substrate[2])
        _faces_from_parent(L1, L1new, L2, T2n, T_amb, blend=(a2, b2, L1["T0"]))
        return T2n

    # the inner scan (subcycleL3_Part1 / _Part2, cF:3367-3412 / 3530-3590) is one C-ABI call per Level-2 substep:
    # all N3 source tables in one launch, then N3 x (fused level step + face prolongation from Level 2)
    rowsP = rows.copy()
    rowsP[:, 6] = P
    nx3, ny3, nz3 = L3["nodes"]
    scratch = {"T": [torch.empty_like(L3["T0"]) for _ in range(2)], "S1": torch.empty_like(L3["S1"]),
               "tables": torch.empty(N3 * (nx3 + ny3 + nz3), device="cuda", dtype=torch.float32)}
    c3, c2 = _coords(L3["node_coords"]), _coords(L2["node_coords"])

    def L3_block(T3, S13, i2, L2new, L2prev, accum=None):
        A, B = scratch["T"]
        Ta, Tb = (B, A) if T3 is A else (A, B)  # T_a != T_in; T_in may be T_b
        kw, flags = {}, ops.STEP_SKIP_FACES | ops.STEP_CLAMP
        if accum is not None:
            S2w, mx_, ac_ = accum
            kw = dict(S2=S2w, accum=ac_, max_accum=mx_)
            flags |= ops.STEP_WRITE_S2 | ops.STEP_ACCUM
        T3n = ops.l3_substeps(_props(properties), _grid(L3), c3, rowsP[i2 * N3:(i2 + 1) * N3], T3, Ta, Tb,
                              scratch["S1"], scratch["tables"], S1_in=S13, n_substrate=int(substrate[3]),
                              flags=flags, faces=(c2, L2new, L2prev, fN3, T_amb), **kw)
        return T3n, scratch["S1"]

    # ---- predictor pass cF:3308-3430 ----
    T2, S12, T3, Tp3, S13 = L2["T0"], L2["S1"], L3["T0"], L3["Tprime0"], L3["S1"]
    Tp3_hist = []
    for i2 in range(N2):
        a2, b2, _, S12n, F2, V2, dt2 = L2_common(T2, S12, T3, Tp3, S13, i2)
        T2n = solve_L2(T2, S12, F2, V2, dt2, a2, b2, L1T)
        T3, S13 = L3_block(T3, S13, i2, T2n, T2)
        Tp3, T2n = getNewTprime(L3, T3, T2n, L2)
        T2, S12 = T2n, S12n
        Tp3_hist.append(Tp3)
    # ---- Level-1 corrector cF:3432-3456 ----
    Tp2, L1T = getNewTprime(L2, T2, L1T, L1)
    _project(Shapes["L3L1"], Tp3, rc3_L1, V1, mode=1, scale=1.0 / F32(dt_all), A2=L3["Tprime0"])
    _project(Shapes["L2L1"], Tp2, rc2_L1, V1, mode=1, scale=1.0 / F32(dt_all), A2=L2["Tprime0"])
    L1T = _solve_L1(Levels, L1["T0"], L1["S1"], F1 + V1, tmp_ne_nn, dt_all, properties, substrate[1])
    # ---- corrector pass cF:3458-3622 ----
    T2, S12, T3, Tp3, S13 = L2["T0"], L2["S1"], L3["T0"], L3["Tprime0"], L3["S1"]
    S23 = L3["S2"].clone()  # updated in place by the corrector substeps
    mx = _f(max_accum_L3).clone()
    ac = _f(accum_L3).clone()
    for i2 in range(N2):
        a2, b2, rc3, S12n, F2, V2, dt2 = L2_common(T2, S12, T3, Tp3, S13, i2)
        _project(Shapes["L3L2"], Tp3_hist[i2], rc3, V2, mode=1, scale=1.0 / F32(dt2), A2=Tp3)
        T2n = solve_L2(T2, S12, F2, V2, dt2, a2, b2, L1T)
        T3, S13 = L3_block(T3, S13, i2, T2n, T2, accum=(S23, mx, ac))
        Tp3, T2n = getNewTprime(L3, T3, T2n, L2)
        T2, S12 = T2n, S12n
    L2["T0"], L2["S1"], L3["T0"], L3["Tprime0"], L3["S1"], L3["S2"] = T2, S12, T3, Tp3, S13, S23
    L2["Tprime0"], L1["T0"] = getNewTprime(L2, L2["T0"], L1T, L1)
    _scatter_L0(Levels)
    return Levels, None, None, None, mx, ac


# ----------------------------------------------------------------------------------------------
# window shift
# ----------------------------------------------------------------------------------------------
def _trunc_shift(v, h):
    """(v / h + 1e-2).astype(int): truncation toward zero of a float32 quotient (cF:1716-1718)."""
    return int(np.asarray(F32(v) / F32(h) + F32(1e-2)).astype(int))


def _constrain(vtot, L):
    b = L["bounds"]
    return [np.clip(F32(vtot[i]), F32(b[k][0]), F32(b[k][1])) for i, k in enumerate(("ix", "iy", "iz"))]


def moveEverything(v, vstart, Levels, move_v, LInterp, L1L2Eratio, L2L3Eratio, height):
    """cF:2400-2510: integer-cell shift of the Level-3 / Level-2 windows; T0 / T'0 re-interpolated at the new
    window nodes; overlap index sets updated; S1 / S2 regathered from Level 0.  ``Shapes`` = per-pair
    fine-element -> parent-cell grouping (three small int arrays each)."""
    torch = _torch()
    _ensure_fields(Levels)
    L0, L1, L2, L3 = Levels
    vtot = np.asarray(_host(v), F32) - np.asarray(_host(vstart), F32)
    # ---- Level 3 (shifts in Level-2 cells) ----
    v3 = _constrain(vtot, L3)
    h2 = L2["h"]
    s3 = [_trunc_shift(v3[i], h2[i]) for i in range(3)]
    new3 = [(np.asarray(L3["init_node_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32) for i in range(3)]
    L3["overlapNodes"] = [np.asarray(L3["orig_overlap_nodes"][i]) + s3[i] for i in range(3)]
    L3["overlapCoords"] = [(np.asarray(L3["orig_overlap_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32)
                           for i in range(3)]
    tgt = _coords(new3)
    n3 = L3["nn"]
    Tp3 = ops.interp(_coords(L3["node_coords"]), L3["Tprime0"], tgt, torch.empty(n3, device="cuda"))
    Tp2on3 = ops.interp(_coords(L2["node_coords"]), L2["Tprime0"], tgt, torch.empty(n3, device="cuda"))
    T1on3 = ops.interp(_coords(L1["node_coords"]), L1["T0"], tgt, torch.empty(n3, device="cuda"))
    L3["T0"] = T1on3 + (Tp2on3 + Tp3)
    L3["Tprime0"] = Tp3
    L3["node_coords"] = new3
    # ---- Level 2 (shifts in Level-1 cells in x, y; whole layers in z) ----
    v2 = _constrain(vtot, L2)
    h1 = [L1["h"][0], L1["h"][1], F32(height)]
    s2 = [_trunc_shift(v2[i], h1[i]) for i in range(3)]
    new2 = [(np.asarray(L2["init_node_coors"][i], F32) + F32(h1[i]) * s2[i]).astype(F32) for i in range(3)]
    move_v = [s2[i] * int(L1L2Eratio[i]) for i in range(3)]
    sz1 = _trunc_shift(v2[2], L1["h"][2])
    o, c = L2["orig_overlap_nodes"], L2["orig_overlap_coors"]
    L2["overlapNodes"] = [np.asarray(o[0]) + s2[0], np.asarray(o[1]) + s2[1], np.asarray(o[2]) + sz1]
    L2["overlapCoords"] = [(np.asarray(c[0], F32) + F32(L1["h"][0]) * s2[0]).astype(F32),
                           (np.asarray(c[1], F32) + F32(L1["h"][1]) * s2[1]).astype(F32),
                           (np.asarray(c[2], F32) + F32(height) * _trunc_shift(v2[2], height)).astype(F32)]
    tgt = _coords(new2)
    n2 = L2["nn"]
    Tp2 = ops.interp(_coords(L2["node_coords"]), L2["Tprime0"], tgt, torch.empty(n2, device="cuda"))
    T1on2 = ops.interp(_coords(L1["node_coords"]), L1["T0"], tgt, torch.empty(n2, device="cuda"))
    L2["Tprime0"] = Tp2
    L2["T0"] = T1on2 + Tp2
    L2["node_coords"] = new2
    LInterp = [interpolatePointsMatrix(L1, new2), None]
    # Level-3 overlap indices are relative to Level 2, which has itself moved
    L3["overlapNodes"] = [L3["overlapNodes"][i] - move_v[i] for i in range(3)]
    # ---- Level 0 index sets of the two windows ----
    r3 = [int(x) for x in L2L3Eratio]
    L0["overlapNodes"] = [np.asarray(L0["orig_overlap_nodes"][i]) + r3[i] * s3[i] for i in range(3)]
    L0["overlapCoords"] = [(np.asarray(L0["orig_overlap_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32)
                           for i in range(3)]
    hz = [L1["h"][0], L1["h"][1], L2["h"][2]]
    rz = [int(L1L2Eratio[0]) * r3[0], int(L1L2Eratio[1]) * r3[1], r3[2]]
    s0 = [_trunc_shift(v2[i], hz[i]) for i in range(3)]
    L0["overlapNodes_L2"] = [np.asarray(L0["orig_overlap_nodes_L2"][i]) + rz[i] * s0[i] for i in range(3)]
    L0["overlapCoords_L2"] = [(np.asarray(L0["orig_overlap_coors_L2"][i], F32) + F32(hz[i]) * s0[i]).astype(F32)
                              for i in range(3)]
    L0["overlapNodes"][2] = L0["overlapNodes"][2] - move_v[2] * r3[2]
    L0["overlapNodes_L2"][2] = L0["overlapNodes_L2"][2] - move_v[2] * r3[2]
    L0["idx"] = levels.overlap_ids(L0["overlapNodes"], L0["nodes"][0], L0["nodes"][1])
    L0["idx_L2"] = levels.overlap_ids(L0["overlapNodes_L2"], L0["nodes"][0], L0["nodes"][1])
    # ---- state regather from Level 0 (cF:2500-2502) ----
    L0["S1"], L0["S2"] = _f(L0["S1"]), _dev(L0["S2"], torch.bool)
    i3, i2 = _index3(L0["overlapNodes"]), _index3(L0["overlapNodes_L2"])
    L2["S1"] = ops.box_copy(L0["S1"], torch.empty(n2, device="cuda"), i2, L0["nodes"][0], L0["nodes"][1], scatter=False)
    L3["S1"] = ops.box_copy(L0["S1"], torch.empty(n3, device="cuda"), i3, L0["nodes"][0], L0["nodes"][1], scatter=False)
    L3["S2"] = ops.box_copy(L0["S2"], torch.empty(n3, device="cuda", dtype=torch.bool), i3, L0["nodes"][0],
                            L0["nodes"][1], scatter=False)
    LInterp[1] = interpolatePointsMatrix(L2, new3)
    Shapes = {"L2L1": _pair_cells(L2, L1), "L3L1": _pair_cells(L3, L1), "L3L2": _pair_cells(L3, L2)}
    return Levels, Shapes, LInterp, move_v


# ----------------------------------------------------------------------------------------------
# melt-time bookkeeping and monitors
# ----------------------------------------------------------------------------------------------
def melting_temp(temps, delt_T, T_melt, accum_time, idx):
    """cF:3696-3712: accum_time[idx] += (temps > T_melt) * dt."""
    torch = _torch()
    acc = _f(accum_time).clone()
    idx_t = _dev(np.asarray(_host(idx)).astype(np.int64)) if not isinstance(idx, torch.Tensor) else idx.long()
    above = (_f(temps) > float(F32(T_melt))).to(torch.float32) * float(F32(_host(delt_T)))
    acc.index_add_(0, idx_t, above)
    return acc


def printLevelMaxMin(Ls, Lnames):
    """cF:3635-3665: print min / max of every level's T0 and stop on invalid physics (non-finite, <= 0 or
    > 1e5 K) like the reference (``sys.exit(1)``); one fused reduction per level instead of two host syncs."""
    import sys

    from . import output

    for name, (lo, hi, bad) in zip(Lnames, output.level_minmax([None] + list(Ls))):
        print(f"{name}: min T = {lo:.2f} K, max T = {hi:.2f} K")
        if bad or not (0 < lo <= 1e5) or not (0 < hi <= 1e5):
            print("Terminating program: temperature out of range")
            sys.exit(1)


def save_object(obj, filename):
    """gm:398 / cF save_object: ``obj`` = [Levels, accum_time, max_accum_time, time_inc, record_inc]; written as a
    raw-dump checkpoint directory next to where the reference writes its ``.pkl`` (output.save_checkpoint)."""
    from . import output

    path = str(filename)
    return output.save_checkpoint(path[:-4] if path.endswith(".pkl") else path, *obj)


def levelMaxMin(Ls):
    """printLevelMaxMin cF:3635-3665 without the prints / exit: [(min, max, ok)] for levels 1.."""
    res = []
    for i in range(1, len(Ls)):
        T = _f(Ls[i]["T0"])
        lo, hi = float(T.min()), float(T.max())
        res.append((lo, hi, all(math.isfinite(x) and 0 < x <= 1e5 for x in (lo, hi))))
    return res
