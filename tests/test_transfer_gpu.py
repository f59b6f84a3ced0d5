"""Direct CUDA-vs-oracle parity of the inter-level transfer kernels (csrc/k_transfer.cu) through the C ABI:

* ``interp_kernel`` (gomelt_interp_f32) in every mode - set / add / RSUB, two-field time blend, faces-only, clamp,
  scatter through an index map - against interpolatePoints cF:1131-1210, assignBCsFine cF:1598-1620 and
  getNewTprime cF:2060-2099;
* ``faces_gather_kernel`` + ``faces_blend_kernel`` against the same face prolongation;
* ``project_cells_kernel`` + ``project_nodes_kernel`` (gomelt_project_f32), mode 0 and mode 1, against
  computeCoarseTprimeTerm_jax cF:1477-1565 / computeCoarseTprimeMassTerm_jax cF:1396-1474 on the materialised
  operators of computeCoarseFineShapeFunctions cF:1213-1358;
* ``coarse_source_table_kernel`` + ``rank1_kernel`` against computeSources cF:928-988 / computeLevelSource
  cF:2667-2730;
* ``box_copy_kernel`` (gather / scatter, float32 and uint8) against NumPy fancy indexing (getOverlapRegion
  cF:1642-1669).

Sizes: examples/example.json (Level 3 100x100x10 el at h = 0.02 in Level 2 100x100x10 at h = 0.04 in Level 1
50x20x30 at h = 0.2: ratios 2, 5 and 10), shifted windows, windows clipped at / hanging over the parent's edge,
a non-integer ratio with unequal fine-element counts per parent cell, and the committed edge-case inputs of
tests/golden/edge_cases.py (reference outputs: edge_cases_reference.npz).

Tolerances: interpolation weights follow the reference's float32 formulas operation for operation, only the order
of the 8-term sum may differ: <= 2e-6 of the field's magnitude; exact zeros (outside the +-1e-2 window) bit-exact.
Projected vectors are sums over up to 1000 fine elements x 8 Gauss points per parent node in a different order than
the reference's scatter-add: <= 3e-5 of the vector's max norm (float32).
"""
import os
import sys

import numpy as np
import pytest

from oracle import computeFunctions as cF
from oracle import transfer as otr
from oracle.util import make_level

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.gpu
F32 = np.float32
TOL_I = 2e-6
TOL_P = 3e-5


def _lv(elements, bounds):
    return make_level(elements, bounds)


def _L1():
    return _lv((50, 20, 30), ((0.0, 10.0), (0.0, 4.0), (-4.0, 2.0)))


def _L2(dx=0.0, dy=0.0, dz=0.0):
    return _lv((100, 100, 10), ((0.0 + dx, 4.0 + dx), (0.0 + dy, 4.0 + dy), (-0.4 + dz, 0.0 + dz)))


def _L3(dx=0.0, dy=0.0, dz=0.0):
    return _lv((100, 100, 10), ((1.0 + dx, 3.0 + dx), (1.0 + dy, 3.0 + dy), (-0.2 + dz, 0.0 + dz)))


def _field(lv, seed, lo=300.0, hi=1900.0):
    rng = np.random.default_rng(seed)
    x, y, z = lv["node_coords"]
    X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]
    f = lo + (hi - lo) * np.exp(-0.6 * ((X - x.mean()) ** 2 + (Y - y.mean()) ** 2)) * np.exp(1.5 * (Z - z[-1]))
    return (f + 5.0 * rng.standard_normal(f.shape)).astype(F32).reshape(-1)


def _relmax(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(float(np.max(np.abs(b))), 1e-30))


@pytest.fixture(scope="module")
def T(gm):
    import torch

    class Dev:
        torch_ = torch

        @staticmethod
        def f(a):
            return torch.as_tensor(np.ascontiguousarray(a, dtype=F32)).cuda()

        @staticmethod
        def i(a):
            return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).cuda()

        @staticmethod
        def coords(lv):
            return [Dev.f(c) for c in lv["node_coords"]]

        @staticmethod
        def h(t):
            torch.cuda.synchronize()
            return t.cpu().numpy()

    return Dev


# ------------------------------------------------------------------------------------------------------------
# interpolation
# ------------------------------------------------------------------------------------------------------------
PAIRS = {
    "L1->L2 (ratio 5)": (_L1, lambda: _L2()),
    "L2->L3 (ratio 2)": (lambda: _L2(), lambda: _L3()),
    "L1->L3 (ratio 10)": (_L1, lambda: _L3()),
    "L1->L2 shifted 7 cells x, 2 y, 1 layer z": (_L1, lambda: _L2(1.4, 0.4, 0.04)),
    "L2->L3 shifted (3, -5, 0) parent cells": (lambda: _L2(), lambda: _L3(0.12, -0.2, 0.0)),
    # a Level-2 window pushed against the +x / +y faces of Level 1 (jit_constrain_v cF:1672-1695 clips it there)
    "L1->L2 clipped at the +x,+y edge": (_L1, lambda: _L2(6.0, 0.0, 0.0)),
    # targets hanging over the parent on -x and +z: zeros outside the +-1e-2 window
    "L1->L2 hanging over -x and +z": (_L1, lambda: _L2(-0.52, 0.0, 2.2)),
    "parent coarser by 2.5 (non-integer ratio)": (lambda: _lv((16, 12, 6), ((0.0, 4.0), (0.0, 3.0), (-1.5, 0.0))),
                                                  lambda: _lv((25, 20, 9), ((0.5, 3.0), (0.5, 2.5), (-0.9, 0.0)))),
}


@pytest.mark.parametrize("pair", list(PAIRS))
def test_interp_set_add_rsub_clamp(gm, T, pair):
    src, tgt = (f() for f in PAIRS[pair])
    u = _field(src, 1)
    want = otr.interpolatePoints(src, u, tgt["node_coords"])
    sc, tc, du = T.coords(src), T.coords(tgt), T.f(u)
    out = T.torch_.full((tgt["nn"],), -3.0, device="cuda")
    got = T.h(gm.ops.interp(sc, du, tc, out))
    assert np.array_equal(got == 0, want == 0), "support of the interpolant (the +-1e-2 window) differs"
    assert _relmax(got, want) <= TOL_I
    # ADD and RSUB on a preset output / base
    base = _field(tgt, 2)
    got = T.h(gm.ops.interp(sc, du, tc, T.f(base), mode=gm._lib.INTERP_ADD))
    assert _relmax(got, base + want) <= TOL_I
    out = T.torch_.empty(tgt["nn"], device="cuda")
    got = T.h(gm.ops.interp(sc, du, tc, out, mode=gm._lib.INTERP_RSUB, base=T.f(base)))
    # (base - I cancels: the error is measured against the magnitude of the operands)
    assert float(np.max(np.abs(got - (base - want)))) <= TOL_I * float(np.max(np.abs(base)))
    # clamp (gm:215-218: max(I(T0), T_amb))
    cl = float(np.median(want))
    got = T.h(gm.ops.interp(sc, du, tc, T.torch_.empty(tgt["nn"], device="cuda"), clamp_min=cl))
    assert _relmax(got, np.maximum(want, F32(cl))) <= TOL_I


@pytest.mark.parametrize("pair", list(PAIRS))
def test_interp_marching_kernel_equals_the_per_target_kernel_bit_for_bit(gm, T, pair):
    """A full-grid interpolation runs the marching kernel (parent values kept in registers along y, the validity window
    from one product); with an index map - here the identity - the same call runs the per-target kernel.  Same bits in
    every mode, incl. the pairs whose targets hang over the parent (those take the plain evaluation inside the marching
    kernel) and the non-nested pair."""
    import torch

    src, tgt = (f() for f in PAIRS[pair])
    u, u2, base = T.f(_field(src, 11)), T.f(_field(src, 12, hi=1400.0)), T.f(_field(tgt, 13))
    sc, tc = T.coords(src), T.coords(tgt)
    ident = [torch.arange(n, dtype=torch.int32, device="cuda") for n in tgt["nodes"]]
    e = lambda: torch.empty(tgt["nn"], device="cuda")
    for kw in ({}, {"u2": u2, "alpha": 0.4, "beta": 0.6}, {"mode": gm._lib.INTERP_RSUB, "base": base}, {"clamp_min": 500.0}):
        a = gm.ops.interp(sc, u, tc, e(), **kw)
        b = gm.ops.interp(sc, u, tc, e(), index_map=(ident[0], ident[1], ident[2], tgt["nodes"][0], tgt["nodes"][1]), **kw)
        assert torch.equal(a, b), kw
    a = gm.ops.interp(sc, u, tc, base.clone(), mode=gm._lib.INTERP_ADD)
    b = gm.ops.interp(sc, u, tc, base.clone(), mode=gm._lib.INTERP_ADD, index_map=(ident[0], ident[1], ident[2], tgt["nodes"][0], tgt["nodes"][1]))
    assert torch.equal(a, b)


@pytest.mark.parametrize("pair", ["L1->L2 (ratio 5)", "L2->L3 shifted (3, -5, 0) parent cells",
                                  "parent coarser by 2.5 (non-integer ratio)"])
def test_interp_faces_only_and_blend(gm, T, pair):
    """assignBCsFine cF:1598-1620 with the time blend of cF:3386-3389 and the following max(T_amb, .): only the five
    Dirichlet faces are written; the compact gather + blend pair gives the same faces."""
    src, tgt = (f() for f in PAIRS[pair])
    u, u2 = _field(src, 3), _field(src, 4, hi=1500.0)
    a, b = F32(0.6), F32(1) - F32(0.6)
    TfAll = otr.interpolatePoints(src, a * u + b * u2, tgt["node_coords"])
    inside = _field(tgt, 5)
    want = np.maximum(cF.assignBCsFine(np.full(tgt["nn"], F32(1e9)), TfAll, tgt["BC"]), F32(400.0))  # faces: max(400, I)
    sc, tc = T.coords(src), T.coords(tgt)
    got = T.h(gm.ops.interp(sc, T.f(u), tc, T.f(inside), u2=T.f(u2), alpha=float(a), beta=float(b), faces_only=True,
                            clamp_min=400.0))
    face = want < F32(1e8)  # the nodes assignBCsFine wrote
    assert face.sum() == tgt["nodes"][0] * tgt["nodes"][1] + 2 * (tgt["nodes"][0] + tgt["nodes"][1] - 2) * (tgt["nodes"][2] - 1)
    top_only = np.setdiff1d(np.asarray(tgt["BC"][5]), np.nonzero(face)[0])
    assert np.array_equal(got[~face], inside[~face]), "faces_only touched a non-face node"
    assert top_only.size and np.array_equal(got[top_only], inside[top_only]), "the free top face was written"
    assert _relmax(got[face], want[face]) <= 2 * TOL_I
    # compact form: both parents gathered once, blended per substep
    lib = gm._lib.load()
    g = gm._lib.make_grid(tgt["nodes"], tgt["h"])
    nface = int(lib.gomelt_faces_count(g.nx, g.ny, g.nz))
    assert nface == int(face.sum())
    scratch = T.torch_.empty(2 * nface, device="cuda")
    ia = gm.ops._interp_args(sc, T.f(u), tc, None, u2=T.f(u2), faces_only=True)
    keep = (ia, sc, tc)  # noqa: F841  (device arrays stay alive through the launches)
    du, du2 = T.f(u), T.f(u2)
    ia.u, ia.u2 = du.data_ptr(), du2.data_ptr()
    import ctypes as C

    gm._lib.check(lib.gomelt_faces_gather_f32(C.byref(ia), gm._lib.ptr(scratch[:nface]), gm._lib.ptr(scratch[nface:]),
                                              gm._lib.stream_ptr()))
    out = T.f(inside)
    gm._lib.check(lib.gomelt_faces_blend_f32(gm._lib.ptr(scratch[:nface]), gm._lib.ptr(scratch[nface:]), g.nx, g.ny,
                                             g.nz, float(a), float(b), 1, 400.0, gm._lib.ptr(out), gm._lib.stream_ptr()))
    got2 = T.h(out)
    assert np.array_equal(got2[~face], inside[~face])
    assert _relmax(got2[face], want[face]) <= 4 * TOL_I


@pytest.mark.parametrize("shift", [(0, 0, 0), (3, -2, 0)])
def test_inject_and_tprime(gm, T, shift):
    """getNewTprime cF:2060-2099 through the drop-in (index-map scatter + RSUB) on the example's Level 3 in Level 2,
    window at its initial place and shifted by whole parent cells."""
    L2 = _L2()
    L3 = _L3(shift[0] * 0.04, shift[1] * 0.04, shift[2] * 0.04)
    # overlap set: parent nodes inside the window (cF:185-237), every second fine node
    ov = [np.nonzero((L2["node_coords"][d] >= L3["node_coords"][d][0] - 1e-4) &
                     (L2["node_coords"][d] <= L3["node_coords"][d][-1] + 1e-4))[0] for d in range(3)]
    L3["overlapNodes"] = ov
    L3["overlapCoords"] = [L2["node_coords"][d][ov[d]] for d in range(3)]
    assert [len(o) for o in ov] == [51, 51, 6]
    Tf, Tc = _field(L3, 6), _field(L2, 7, hi=1200.0)
    C2F = otr.interpolatePointsMatrix(L2, L3["node_coords"])
    wantTp, wantTc = otr.getNewTprime(L3, Tf, Tc, L2, C2F)
    cf = gm.computeFunctions
    Tp, Tc2 = cf.getNewTprime(L3, Tf, Tc, L2)
    gotTp, gotTc = T.h(Tp), T.h(Tc2)
    idx = otr.getOverlapRegion(ov, L2["nodes"][0], L2["nodes"][1])
    rest = np.ones(L2["nn"], bool)
    rest[idx] = False
    assert np.array_equal(gotTc[rest], Tc[rest]), "injection wrote outside the overlap set"
    assert _relmax(gotTc[idx], wantTc[idx]) <= TOL_I
    assert float(np.max(np.abs(gotTp - wantTp))) <= TOL_I * float(np.max(np.abs(Tf)))


def test_interp_edge_cases_against_the_reference_outputs(gm, T):
    """tests/golden/edge_cases.py through the CUDA drop-in against edge_cases_reference.npz (written by the reference's
    own functions): targets outside / on / a hair around the parent's faces, state thresholds to the ulp (bit-exact),
    the capped evaporation flux."""
    import edge_cases

    ref = np.load(os.path.join(HERE, "golden", "edge_cases_reference.npz"))
    prod = gm.computeFunctions

    class HostView:  # the drop-in namespace with its CUDA results brought to the host (edge_cases.py speaks NumPy)
        SetupProperties, SetupLevels = staticmethod(prod.SetupProperties), staticmethod(prod.SetupLevels)

        @staticmethod
        def computeStateProperties(*a):
            return tuple(T.h(v) for v in prod.computeStateProperties(*a))

        @staticmethod
        def computeConvRadBC(*a):
            return T.h(prod.computeConvRadBC(*a))

        @staticmethod
        def interpolatePoints(*a):
            return T.h(prod.interpolatePoints(*a))

    got = {k: np.asarray(v) for k, v in edge_cases.run(HostView).items()}
    assert set(got) == set(ref.files)
    for k in ref.files:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        assert a.shape == b.shape, k
        if k.startswith("state_properties_at_thresholds/"):
            if b.dtype == bool or k.split("/")[1].startswith("S"):
                assert np.array_equal(a.astype(b.dtype), b), k
            else:  # k (/1000 folded on the host) and rho*cp (rho folded): one rounding of a constant
                assert _relmax(a, b) <= 4e-7, k
        elif k.startswith("interpolation_outside_the_parent/"):
            assert np.array_equal(a == 0, b == 0), k
            assert _relmax(a, b) <= TOL_I, k
        else:
            assert _relmax(a, b) <= 1e-5, k


# ------------------------------------------------------------------------------------------------------------
# correction-vector projection
# ------------------------------------------------------------------------------------------------------------
PROJ = {
    "L3->L2 (ratio 2, 8 fine elements per cell)": (lambda: _L3(), lambda: _L2()),
    "L2->L1 (ratio 5, 125 per cell: one warp per cell)": (lambda: _L2(), _L1),
    "L3->L1 (ratio 10, 1000 per cell)": (lambda: _L3(), _L1),
    "L3->L2 shifted (3, -5, 0)": (lambda: _L3(0.12, -0.2, 0.0), lambda: _L2()),
    "L2->L1 shifted 7, 2 cells and one 0.04 layer in z (unequal counts per cell in z)": (lambda: _L2(1.4, 0.4, 0.04), _L1),
    "L2->L1 clipped at the +x,+y edge": (lambda: _L2(6.0, 0.0, 0.0), _L1),
    # Level 3 moves in Level-2 cells: relative to Level 1 it starts inside a parent cell (partial first and last cells)
    "L3->L1 shifted by (3, -5) Level-2 cells": (lambda: _L3(0.12, -0.2, 0.0), _L1),
    "L3->L1 shifted by (1, 2) Level-2 cells and one layer": (lambda: _L3(0.04, 0.08, 0.04), _L1),
    "non-integer ratio 2.5 (2 or 3 fine elements per cell)": (
        lambda: _lv((25, 20, 9), ((0.5, 3.0), (0.5, 2.5), (-0.9, 0.0))),
        lambda: _lv((16, 12, 6), ((0.0, 4.0), (0.0, 3.0), (-1.5, 0.0)))),
}


@pytest.mark.parametrize("pair", list(PROJ))
def test_project_grad_and_mass_terms(gm, T, pair, example_props):
    fine, parent = (f() for f in PROJ[pair])
    P = cF.SetupProperties(example_props)
    Tp0 = (_field(fine, 8, lo=-40.0, hi=60.0)).astype(F32)
    Tp1 = (Tp0 + _field(fine, 9, lo=-3.0, hi=4.0)).astype(F32)
    Tf = _field(fine, 10, lo=300.0, hi=2400.0)
    S1 = (np.random.default_rng(11).random(fine["nn"]) > 0.4).astype(F32)
    _, _, k, rc = cF.computeStateProperties(Tf, S1, P, 0)
    Sh = otr.computeCoarseFineShapeFunctions(parent, fine)
    dt = F32(2.5e-5)
    want0 = otr.project(Sh[2], otr._grad_term(fine, Tp0, k, Sh[1]))
    want1 = otr.project(Sh[2], otr._mass_term(fine, Tp1 - Tp0, rc, Sh[0], dt))
    cf = gm.computeFunctions
    cells = cf._pair_cells(fine, parent)
    V = T.torch_.zeros(parent["nn"], device="cuda")
    cf._project(cells, T.f(Tp0), T.f(k), V, mode=0)
    got0 = T.h(V).copy()
    assert np.array_equal(got0 == 0, want0 == 0) or _relmax(got0, want0) <= TOL_P  # same footprint on the parent
    assert _relmax(got0, want0) <= TOL_P, pair
    # mode 1 accumulates on top of mode 0 (cF:1467-1474: Vcu + project(...))
    cf._project(cells, T.f(Tp1), T.f(rc), V, mode=1, scale=1.0 / dt, A2=T.f(Tp0))
    got01 = T.h(V)
    assert _relmax(got01, want0 + want1) <= TOL_P, pair
    # deterministic: a second evaluation is bit-identical
    V2 = T.torch_.zeros(parent["nn"], device="cuda")
    cf._project(cells, T.f(Tp0), T.f(k), V2, mode=0)
    cf._project(cells, T.f(Tp1), T.f(rc), V2, mode=1, scale=1.0 / dt, A2=T.f(Tp0))
    assert np.array_equal(T.h(V2), got01)
    # the general kernel (no tables, everything derived from the coordinate arrays) gives the same vectors
    V3 = T.torch_.zeros(parent["nn"], device="cuda")
    cf._project(cells, T.f(Tp0), T.f(k), V3, mode=0, tiled=False)
    cf._project(cells, T.f(Tp1), T.f(rc), V3, mode=1, scale=1.0 / dt, A2=T.f(Tp0), tiled=False)
    assert _relmax(T.h(V3), want0 + want1) <= TOL_P, pair
    # ... the shared-memory tile kernel (the fast path of windows that do not nest; nested ones take the marching kernel above)
    V5 = T.torch_.zeros(parent["nn"], device="cuda")
    cf._project(cells, T.f(Tp0), T.f(k), V5, mode=0, tiled="tile")
    cf._project(cells, T.f(Tp1), T.f(rc), V5, mode=1, scale=1.0 / dt, A2=T.f(Tp0), tiled="tile")
    assert _relmax(T.h(V5), want0 + want1) <= TOL_P, pair
    # ... and so does the form the steppers use: k / rho*cp evaluated inside the kernel from (T, S1) - no coefficient array
    props = gm._lib.make_props(P)
    V4 = T.torch_.zeros(parent["nn"], device="cuda")
    dT, dS = T.f(Tf), T.f(S1)
    cf._project(cells, T.f(Tp0), None, V4, mode=0, coef_from=(props, dT, dS, 0))
    cf._project(cells, T.f(Tp1), None, V4, mode=1, scale=1.0 / dt, A2=T.f(Tp0), coef_from=(props, dT, dS, 0))
    assert _relmax(T.h(V4), want0 + want1) <= TOL_P, pair
    # (ratio 10 runs half a parent cell row per thread, the two halves added onto a zeroed cell sum: order-independent,
    # hence the bit-identical V2 above)


# ------------------------------------------------------------------------------------------------------------
# projected laser source
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["centre", "shifted, laser near the window edge"])
def test_projected_source_tables(gm, T, case, example_props):
    """computeSources cF:928-988 (Fc on Level 1, Fm on Level 2, one row) and computeLevelSource cF:2667-2730 (mean
    over the rows of a subcycle block) against coarse_source_table_kernel + rank1_kernel."""
    P = cF.SetupProperties(example_props)
    L1, L2 = _L1(), _L2()
    L3 = _L3() if case == "centre" else _L3(0.12, -0.2, 0.0)
    v = np.array([2.0, 2.0, 0.0], F32) if case == "centre" else np.array([3.05, 0.85, 0.0], F32)
    Sh = [None, otr.computeCoarseFineShapeFunctions(L1, L3), otr.computeCoarseFineShapeFunctions(L2, L3)]
    ne_nn = (0, L3["ne"], 0, 0, L3["nn"])
    Fc, Fm, _ = otr.computeSources(L3, v, Sh, ne_nn, P, F32(285.0))
    cf = gm.computeFunctions
    gFc = T.h(cf._projected_source(L3, L1, [v], [285.0], P))
    gFm = T.h(cf._projected_source(L3, L2, [v], [285.0], P))
    assert Fc.max() > 0 and _relmax(gFc, Fc) <= TOL_P
    assert _relmax(gFm, Fm) <= TOL_P
    rows = np.stack([v + np.array([0.01 * i, 0.004 * i, 0.0], F32) for i in range(5)]).astype(F32)
    powers = np.array([285.0, 285.0, 0.0, 150.0, 285.0], F32)
    want = otr.computeLevelSource([None, L1, L2, L3], ne_nn, rows, Sh[2], P, powers)
    got = T.h(cf._projected_source(L3, L2, rows, powers, P))
    assert _relmax(got, want) <= TOL_P


# ------------------------------------------------------------------------------------------------------------
# box gather / scatter
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_box_copy_gather_scatter(gm, T, dtype):
    rng = np.random.default_rng(12)
    big = (501, 201, 21)
    ix, iy, iz = np.arange(120, 221), np.arange(50, 151), np.arange(10, 21)
    idx = otr.getOverlapRegion([ix, iy, iz], big[0], big[1])
    nbig = big[0] * big[1] * big[2]
    if dtype == "f32":
        src = rng.random(nbig).astype(F32)
        win = rng.random(idx.size).astype(F32)
        dsrc, dwin = T.f(src), T.f(win)
        empty = lambda n: T.torch_.empty(n, device="cuda")
    else:
        src = (rng.random(nbig) > 0.5)
        win = (rng.random(idx.size) > 0.5)
        dsrc, dwin = T.torch_.as_tensor(src).cuda(), T.torch_.as_tensor(win).cuda()
        empty = lambda n: T.torch_.empty(n, device="cuda", dtype=T.torch_.bool)
    i3 = [T.i(ix), T.i(iy), T.i(iz)]
    got = T.h(gm.ops.box_copy(dsrc, empty(idx.size), i3, big[0], big[1], scatter=False))
    assert np.array_equal(got, src[idx])
    dst = dsrc.clone()
    got = T.h(gm.ops.box_copy(dwin, dst, i3, big[0], big[1], scatter=True))
    want = src.copy()
    want[idx] = win
    assert np.array_equal(got, want)
