"""GPU parity of K1 (gomelt_level_step_f32) and its helper kernels against the oracle.

Tolerance: BASELINE.json north_star — temperatures within 1e-5 relative (float32); the set of
nodes with T >= T_liquidus identical (nodes whose oracle T lies within 8 ulp of the threshold are
counted and excluded).
"""
import numpy as np
import pytest

from oracle import computeFunctions as cF
from oracle.util import make_level, smooth_field

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _dev(a, dtype=None):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def _setup(gm, example_props, elements, bounds, seed=0, nsub_planes=3):
    rng = np.random.default_rng(seed)
    P = cF.SetupProperties(example_props)
    lv = make_level(elements, bounds)
    T0 = smooth_field(lv, rng)
    S1 = (rng.random(lv["nn"]) > 0.5).astype(np.float32)
    nsub = nsub_planes * lv["nodes"][0] * lv["nodes"][1]
    return P, lv, T0, S1, nsub


def _oracle_step(P, lv, T0, S1, nsub, dt, v, laserP, rhs=None, ne=None, flux=True):
    ne = lv["ne"] if ne is None else ne
    ne_nn = (0, lv["ne"], 0, 0, lv["nn"])
    S1n, S2, k, rc = cF.computeStateProperties(T0, S1, P, nsub)
    F = cF.computeSourcesL3(lv, v, ne_nn, P, laserP) if laserP else np.zeros(lv["nn"], np.float32)
    if flux:
        F = cF.computeConvRadBC(lv, T0, ne, lv["nn"], P, F)
    T = cF.solveMatrixFreeFE(lv, lv["nn"], ne, k, rc, dt, T0, F, 0 if rhs is None else rhs)
    return T, S1n, S2


def _gpu_step(gm, P, lv, T0, S1, nsub, dt, v, laserP, rhs=None, nz_active=None, flags=0, bc5=None,
              z_chunk=0, flux=True, extra=None, fused=False):
    import torch

    ops = gm.ops
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    dT0, dS1 = _dev(T0), _dev(S1)
    nn, nx, ny, nz = lv["nn"], *lv["nodes"]
    src = None
    if laserP:
        coords = [_dev(c) for c in lv["node_coords"]]
        tx, ty, tz = (torch.empty(n, device="cuda") for n in (nx, ny, nz))
        coef = ops.source_tables(props, grid, coords, v, laserP, tx, ty, tz)
        src = (tx, ty, tz, coef)
    top = None
    if flux and fused:  # computeConvRadBC evaluated inside K1 (GOMELT_STEP_FUSED_FLUX)
        flags |= ops.STEP_FUSED_FLUX
    elif flux:
        top = torch.empty(nx * ny, device="cuda")
        ops.surface_flux(props, grid, dT0, top, nz_active=nz_active)
    Tout = torch.full((nn,), -7.0, device="cuda")
    S1o = torch.empty(nn, device="cuda")
    S2o = torch.empty(nn, device="cuda", dtype=torch.uint8)
    kw = dict(extra or {})
    ops.level_step(props, grid, dT0, dS1, Tout, dt, rhs=None if rhs is None else _dev(rhs), src=src,
                   topflux=top, nz_active=nz_active, n_substrate=nsub,
                   flags=flags | ops.STEP_WRITE_S1 | ops.STEP_WRITE_S2, bc5=bc5, S1_out=S1o, S2_out=S2o,
                   z_chunk=z_chunk, **kw)
    torch.cuda.synchronize()
    return Tout.cpu().numpy(), S1o.cpu().numpy(), S2o.cpu().numpy().astype(bool)


@pytest.mark.parametrize("elements,seed", [((36, 28, 12), 0), ((61, 17, 5), 1), ((7, 9, 3), 2), ((100, 100, 10), 3)])
def test_level3_substep_matches_oracle(gm, example_props, elements, seed):
    bounds = ((1.0, 1.0 + 0.02 * elements[0]), (1.0, 1.0 + 0.02 * elements[1]), (-0.02 * elements[2], 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, seed)
    v = np.array([0.5 * (bounds[0][0] + bounds[0][1]), 0.5 * (bounds[1][0] + bounds[1][1]), 0.0], np.float32)
    dt = 1e-5
    Tref, S1ref, S2ref = _oracle_step(P, lv, T0, S1, nsub, dt, v, 285.0)
    T, S1g, S2g = _gpu_step(gm, P, lv, T0, S1, nsub, dt, v, 285.0)
    assert np.isfinite(T).all()
    assert _rel(T, Tref) <= RTOL, _rel(T, Tref)
    assert np.array_equal(S1g, S1ref)
    assert np.array_equal(S2g, S2ref)  # S2 is a pure function of the (identical) input T0: bit-exact


@pytest.mark.parametrize("elements,seed,hot", [((36, 28, 12), 0, 1.0), ((61, 17, 5), 1, 1.3), ((7, 9, 3), 2, 1.0),
                                               ((1, 1, 1), 6, 1.0), ((59, 4, 2), 7, 1.3), ((60, 5, 2), 8, 1.0),
                                               ((123, 41, 3), 9, 1.0)])
def test_fused_surface_flux_matches_oracle_and_separate_kernel(gm, example_props, elements, seed, hot):
    """GOMELT_STEP_FUSED_FLUX: computeConvRadBC cF:2207-2301 inside K1's top plane (ragged tiles, tile
    seams at 60 columns / 4 rows, temperatures past the T_boiling + 1000 cap)."""
    bounds = ((1.0, 1.0 + 0.02 * elements[0]), (1.0, 1.0 + 0.02 * elements[1]), (-0.02 * elements[2], 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, seed, nsub_planes=1)
    T0 = (T0 * hot).astype(np.float32)
    v = np.array([0.5 * (bounds[0][0] + bounds[0][1]), 0.5 * (bounds[1][0] + bounds[1][1]), 0.0], np.float32)
    Tref, _, _ = _oracle_step(P, lv, T0, S1, nsub, 1e-5, v, 285.0)
    Tsep, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0)
    Tfus, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, fused=True)
    Tnone, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, flux=False)
    assert np.isfinite(Tfus).all()
    assert _rel(Tfus, Tref) <= RTOL, _rel(Tfus, Tref)
    assert _rel(Tfus, Tsep) <= 2e-6
    nxy = lv["nodes"][0] * lv["nodes"][1]
    assert np.array_equal(Tfus[:-nxy], Tnone[:-nxy])       # only the top plane carries the surface load
    assert np.abs(Tfus[-nxy:] - Tnone[-nxy:]).max() > 0    # ... and it does carry it
    for zc in (1, 2):
        Tz, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, fused=True, z_chunk=zc)
        assert np.array_equal(Tz, Tfus), zc


def test_fused_surface_flux_on_the_active_layer(gm, example_props):
    """Level 1: the surface load sits on plane nz_active-1, inactive planes above are T_amb."""
    elements = (25, 10, 12)
    bounds = ((0.0, 10.0), (0.0, 4.0), (-2.0, 0.4))
    P, lv, T0, S1, _ = _setup(gm, example_props, elements, bounds, 9)
    nx, ny, nz = lv["nodes"]
    cond = {"x": [301.0, 302.0], "y": [303.0, 304.0], "z": [305.0, 306.0]}
    bc5 = [cond["y"][0], cond["y"][1], cond["x"][0], cond["x"][1], cond["z"][0]]
    for nz_active in (2, 9, nz):
        tmp_ne, tmp_nn = elements[0] * elements[1] * (nz_active - 1), nx * ny * nz_active
        Levels = [None, dict(lv, T0=T0, S1=S1, conditions=cond)]
        Tref = cF.stepGOMELTDwellTime(Levels, (tmp_ne, tmp_nn), (0, 0, lv["nn"]), P, 2e-3, (0, 4 * nx * ny))[1]["T0"]
        T, _, _ = _gpu_step(gm, P, lv, T0, S1, 4 * nx * ny, 2e-3, None, 0.0, nz_active=nz_active,
                            flags=gm.ops.STEP_BC_CONST, bc5=bc5, fused=True)
        assert _rel(T, Tref) <= RTOL, (nz_active, _rel(T, Tref))


def test_z_chunking_is_bit_identical(gm, example_props):
    elements = (40, 33, 17)
    bounds = ((0.0, 0.8), (0.0, 0.66), (-0.34, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 5)
    v = np.array([0.4, 0.3, 0.0], np.float32)
    base = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0)[0]
    for zc in (1, 2, 5, 7, 18):
        T = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, z_chunk=zc)[0]
        assert np.array_equal(T, base), zc


def test_rhs_and_clamp(gm, example_props):
    elements = (30, 22, 9)
    bounds = ((0.0, 1.2), (0.0, 0.88), (-0.36, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 7)
    rng = np.random.default_rng(11)
    rhs = (1e-3 * rng.standard_normal(lv["nn"])).astype(np.float32)
    v = np.array([0.6, 0.4, 0.0], np.float32)
    Tref, _, _ = _oracle_step(P, lv, T0, S1, nsub, 2e-5, v, 0.0, rhs=rhs)
    Tref = np.maximum(np.float32(P["T_amb"]), Tref)
    T, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 2e-5, v, 0.0, rhs=rhs, flags=gm.ops.STEP_CLAMP)
    assert _rel(T, Tref) <= RTOL


def test_level1_active_layer_and_dirichlet_constants(gm, example_props):
    """stepGOMELTDwellTime cF:2617-2664: active elements only, inactive planes <- T_amb, 5 faces
    <- conditions (assignBCs order y-, y+, x-, x+, z-), no clamp."""
    elements = (25, 10, 12)
    bounds = ((0.0, 10.0), (0.0, 4.0), (-2.0, 0.4))
    P, lv, T0, S1, _ = _setup(gm, example_props, elements, bounds, 9)
    nx, ny, nz = lv["nodes"]
    nz_active = 9
    tmp_ne, tmp_nn = elements[0] * elements[1] * (nz_active - 1), nx * ny * nz_active
    nsub = 4 * nx * ny
    cond = {"x": [301.0, 302.0], "y": [303.0, 304.0], "z": [305.0, 306.0]}
    Levels = [None, dict(lv, T0=T0, S1=S1, conditions=cond)]
    out = cF.stepGOMELTDwellTime(Levels, (tmp_ne, tmp_nn), (0, 0, lv["nn"]), P, 2e-3, (0, nsub))
    Tref = out[1]["T0"]
    bc5 = [cond["y"][0], cond["y"][1], cond["x"][0], cond["x"][1], cond["z"][0]]
    T, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 2e-3, None, 0.0, nz_active=nz_active,
                        flags=gm.ops.STEP_BC_CONST, bc5=bc5)
    assert np.isfinite(Tref).all()
    assert _rel(T, Tref) <= RTOL, _rel(T, Tref)


def test_skip_faces_leaves_dirichlet_faces_untouched(gm, example_props):
    elements = (12, 11, 6)
    bounds = ((0.0, 0.24), (0.0, 0.22), (-0.12, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 4)
    v = np.array([0.1, 0.1, 0.0], np.float32)
    T, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, flags=gm.ops.STEP_SKIP_FACES)
    full, _, _ = _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0)
    nx, ny, nz = lv["nodes"]
    T3, F3 = T.reshape(nz, ny, nx), full.reshape(nz, ny, nx)
    face = np.zeros((nz, ny, nx), bool)
    face[0] = True
    face[:, 0] = face[:, -1] = True
    face[:, :, 0] = face[:, :, -1] = True
    assert (T3[face] == -7.0).all()          # sentinel untouched on the 5 Dirichlet faces
    assert np.array_equal(T3[~face], F3[~face])  # top plane interior is computed (free surface)


def test_melt_time_bookkeeping(gm, example_props):
    """cF:3568-3578 fused into the step."""
    import torch

    elements = (20, 20, 6)
    bounds = ((0.0, 0.4), (0.0, 0.4), (-0.12, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 13)
    rng = np.random.default_rng(17)
    nn = lv["nn"]
    prev = rng.random(nn) > 0.5
    acc = rng.random(nn).astype(np.float32) * 1e-3
    mx = rng.random(nn).astype(np.float32) * 1e-3
    S2 = T0 >= np.float32(P["T_liquidus"])
    reset = acc * ((~prev) & S2)
    mx_ref = np.maximum(reset, mx)
    acc_ref = (acc + np.float32(1e-5) * S2 - reset).astype(np.float32)
    dS2 = _dev(prev.astype(np.uint8))
    dacc, dmx = _dev(acc), _dev(mx)
    v = np.array([0.2, 0.2, 0.0], np.float32)
    _gpu_step(gm, P, lv, T0, S1, nsub, 1e-5, v, 285.0, flags=gm.ops.STEP_ACCUM,
              extra=dict(S2_prev=dS2, accum=dacc, max_accum=dmx))
    torch.cuda.synchronize()
    assert np.array_equal(dmx.cpu().numpy(), mx_ref)
    assert np.allclose(dacc.cpu().numpy(), acc_ref, rtol=1e-6, atol=0)


def test_state_props_kernel(gm, example_props):
    import torch

    rng = np.random.default_rng(3)
    P = cF.SetupProperties(example_props)
    n = 100003
    T = rng.uniform(250, 3500, n).astype(np.float32)
    T[:7] = np.float32([P["T_liquidus"], P["T_solidus"], np.nextafter(np.float32(P["T_liquidus"]), np.float32(0)),
                        np.nextafter(np.float32(P["T_solidus"]), np.float32(1e9)), 298.15, 1609.0, 1533.0])
    S1 = rng.choice(np.float32([0.0, 1.0, 0.3, 0.499, 0.5, 0.7]), n)
    S1r, S2r, kr, rr = cF.computeStateProperties(T, S1, P, 1000)
    S1o, ko, ro = (torch.empty(n, device="cuda") for _ in range(3))
    S2o = torch.empty(n, device="cuda", dtype=torch.uint8)
    gm.ops.state_props(gm._lib.make_props(P), _dev(T), _dev(S1), 1000, S1_out=S1o, S2_out=S2o, k_out=ko, rhocp_out=ro)
    assert np.array_equal(S1o.cpu().numpy(), S1r)
    assert np.array_equal(S2o.cpu().numpy().astype(bool), S2r)
    assert np.allclose(ko.cpu().numpy(), kr, rtol=3e-7, atol=0)
    assert np.allclose(ro.cpu().numpy(), rr, rtol=3e-7, atol=0)


def test_surface_flux_and_source_tables(gm, example_props):
    import torch

    elements = (31, 23, 4)
    bounds = ((0.0, 0.62), (0.0, 0.46), (-0.08, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 21)
    T0 = (T0 * 1.3).astype(np.float32)  # push some nodes past T_boiling + 1000 (the min() cap)
    nx, ny, nz = lv["nodes"]
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    ref = cF.computeConvRadBC(lv, T0, lv["ne"], lv["nn"], P, 0)
    flux = torch.empty(nx * ny, device="cuda")
    gm.ops.surface_flux(props, grid, _dev(T0), flux)
    got = flux.cpu().numpy()
    top = ref[-nx * ny:]
    assert np.abs(ref[:-nx * ny]).max() == 0
    assert np.max(np.abs(got - top)) <= 2e-6 * np.max(np.abs(top))
    v = np.array([0.3, 0.2, 0.0], np.float32)
    Fref = cF.computeSourcesL3(lv, v, (0, lv["ne"], 0, 0, lv["nn"]), P, 285.0)
    coords = [_dev(c) for c in lv["node_coords"]]
    tx, ty, tz = (torch.empty(n, device="cuda") for n in (nx, ny, nz))
    coef = gm.ops.source_tables(props, grid, coords, v, 285.0, tx, ty, tz)
    F = coef * np.einsum("k,j,i->kji", tz.cpu().numpy(), ty.cpu().numpy(), tx.cpu().numpy()).reshape(-1)
    assert np.max(np.abs(F - Fref)) <= 3e-6 * np.max(np.abs(Fref))


def test_l3_substeps_one_call_equals_single_launches(gm, example_props):
    """gomelt_l3_substeps_f32 (the inner scan of subcycleGOMELT as one call) against the same substeps
    launched one by one, and against the oracle's substep sequence."""
    import torch

    ops = gm.ops
    elements = (47, 21, 6)
    bounds = ((0.0, 0.94), (0.0, 0.42), (-0.12, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 31, nsub_planes=2)
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    nx, ny, nz = lv["nodes"]
    n = 5
    rows = np.zeros((n, 7), np.float32)
    for i in range(n):
        rows[i] = (0.40 + 0.013 * i, 0.2 + 0.004 * i, 0.0, 1, 1, 1e-5 * (1 + 0.1 * i), 285.0 - 10 * i)
    coords = [_dev(c) for c in lv["node_coords"]]
    # one call
    Tin, S1w = _dev(T0), _dev(S1)
    A, B = torch.empty_like(Tin), torch.empty_like(Tin)
    tables = torch.empty(n * (nx + ny + nz), device="cuda")
    S2 = torch.zeros(lv["nn"], device="cuda", dtype=torch.uint8)
    last = ops.l3_substeps(props, grid, coords, rows, Tin, A, B, S1w.clone(), tables, n_substrate=nsub,
                           flags=ops.STEP_CLAMP | ops.STEP_WRITE_S2, S2=S2)
    assert last is A  # n odd: the last substep wrote T_a
    assert torch.equal(Tin, _dev(T0))  # T_in is preserved
    # one by one
    cur, S1c = _dev(T0), _dev(S1)
    S2c = torch.zeros(lv["nn"], device="cuda", dtype=torch.uint8)
    tx, ty, tz = (torch.empty(m, device="cuda") for m in (nx, ny, nz))
    Tref, S1ref = T0.copy(), S1.copy()
    ne_nn = (0, lv["ne"], 0, 0, lv["nn"])
    for i in range(n):
        coef = ops.source_tables(props, grid, coords, rows[i, :3], float(rows[i, 6]), tx, ty, tz)
        nxt = torch.empty_like(cur)
        ops.level_step(props, grid, cur, S1c, nxt, float(rows[i, 5]), src=(tx, ty, tz, coef), n_substrate=nsub,
                       flags=ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_WRITE_S2 | ops.STEP_FUSED_FLUX,
                       S1_out=S1c, S2_out=S2c)
        cur = nxt
        S1ref, S2ref, k, rc = cF.computeStateProperties(Tref, S1ref, P, nsub)
        F = cF.computeSourcesL3(lv, rows[i, :3], ne_nn, P, rows[i, 6])
        F = cF.computeConvRadBC(lv, Tref, lv["ne"], lv["nn"], P, F)
        Tref = np.maximum(np.float32(P["T_amb"]),
                          cF.solveMatrixFreeFE(lv, lv["nn"], lv["ne"], k, rc, np.float32(rows[i, 5]), Tref, F, 0))
    torch.cuda.synchronize()
    assert torch.equal(last, cur)
    assert torch.equal(S2, S2c)
    assert _rel(last.cpu().numpy(), Tref) <= RTOL
    assert np.array_equal(S2.cpu().numpy().astype(bool), S2ref)


def _faces_mask(nx, ny, nz):
    face = np.zeros((nz, ny, nx), bool)
    face[0] = True
    face[:, 0] = face[:, -1] = True
    face[:, :, 0] = face[:, :, -1] = True
    return face


@pytest.mark.parametrize("elements,nsub_planes,shape", [
    ((100, 100, 10), 0, "l3_sub"), ((129, 37, 9), 3, "l3_sub"), ((61, 5, 4), 2, "l3_step"),
    ((140, 66, 12), 4, "l2"), ((75, 23, 7), 0, "l2")])
def test_dirichlet_side_face_kernel_matches_general_kernel_and_oracle(gm, example_props, elements, nsub_planes, shape):
    """The fast kernel for the steppers' Dirichlet-side-face call shapes (k_level_step_v3.cuh: inward-shifted
    overlapping tiles, folded y stage) against the general kernel on the same call and against the oracle:
    interior T within 1e-5, S1' bit-exact on EVERY node (faces included), the five Dirichlet faces untouched."""
    import torch

    ops = gm.ops
    ex, ey, ez = elements
    bounds = ((0.0, 0.02 * ex), (0.0, 0.02 * ey), (-0.02 * ez, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 11, nsub_planes=nsub_planes)
    T0 = (T0 * 1.2).astype(np.float32)
    nx, ny, nz = lv["nodes"]
    v = np.array([0.01 * ex, 0.01 * ey, 0.0], np.float32)
    dt = 1e-5
    rng = np.random.default_rng(5)
    rhs = (rng.standard_normal(lv["nn"]) * 1e-4).astype(np.float32) if shape == "l2" else None
    laserP = 0.0 if shape == "l2" else 285.0
    flags = ops.STEP_SKIP_FACES | ops.STEP_CLAMP
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    dT0, dS1 = _dev(T0), _dev(S1)
    src = None
    if laserP:
        coords = [_dev(c) for c in lv["node_coords"]]
        tx, ty, tz = (torch.empty(n, device="cuda") for n in (nx, ny, nz))
        src = (tx, ty, tz, ops.source_tables(props, grid, coords, v, laserP, tx, ty, tz))
    outs = []
    for extra in (0, ops.STEP_GENERAL_KERNEL):
        Tout = torch.full((lv["nn"],), -7.0, device="cuda")
        S1o = torch.full((lv["nn"],), -3.0, device="cuda")
        kw = dict(S1_out=S1o) if shape == "l3_sub" else {}
        f = flags | ops.STEP_FUSED_FLUX | extra | (ops.STEP_WRITE_S1 if shape == "l3_sub" else 0)
        ops.level_step(props, grid, dT0, dS1, Tout, dt, rhs=None if rhs is None else _dev(rhs), src=src,
                       n_substrate=nsub, flags=f, **kw)
        torch.cuda.synchronize()
        outs.append((Tout.cpu().numpy(), S1o.cpu().numpy()))
    (Tf, S1f), (Tg, S1g) = outs
    face = _faces_mask(nx, ny, nz).ravel()
    assert (Tf[face] == -7.0).all() and (Tg[face] == -7.0).all()
    assert _rel(Tf[~face], Tg[~face]) <= 2e-6, _rel(Tf[~face], Tg[~face])
    if shape == "l3_sub":
        assert np.array_equal(S1f, S1g)
    Tref, S1ref, _ = _oracle_step(P, lv, T0, S1, nsub, dt, v, laserP, rhs=rhs)
    Tref = np.maximum(np.float32(P["T_amb"]), Tref)
    assert _rel(Tf[~face], Tref[~face]) <= RTOL, _rel(Tf[~face], Tref[~face])
    if shape == "l3_sub":
        assert np.array_equal(S1f, S1ref)


def test_dirichlet_side_face_kernel_z_chunks_and_inactive_planes(gm, example_props):
    """z-chunked launches of the fast kernel are bit-identical to the single-chunk launch, and planes above
    nz_active are filled with T_amb on the owned nodes only."""
    import torch

    ops = gm.ops
    elements = (90, 30, 11)
    bounds = ((0.0, 1.8), (0.0, 0.6), (-0.22, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 3, nsub_planes=2)
    nx, ny, nz = lv["nodes"]
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    dT0, dS1 = _dev(T0), _dev(S1)
    rhs = _dev((np.random.default_rng(1).standard_normal(lv["nn"]) * 1e-4).astype(np.float32))
    res = []
    for zc, extra in ((0, 0), (3, 0), (2, 0), (0, ops.STEP_GENERAL_KERNEL)):
        Tout = torch.full((lv["nn"],), -7.0, device="cuda")
        ops.level_step(props, grid, dT0, dS1, Tout, 1e-5, rhs=rhs, n_substrate=nsub, nz_active=nz - 3,
                       flags=ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_FUSED_FLUX | extra, z_chunk=zc)
        torch.cuda.synchronize()
        res.append(Tout.cpu().numpy())
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
    face = _faces_mask(nx, ny, nz).ravel()
    assert (res[0][face] == -7.0).all()
    assert _rel(res[0][~face], res[3][~face]) <= 2e-6
    top = res[0].reshape(nz, ny, nx)[nz - 3:, 1:-1, 1:-1]
    assert (top == np.float32(P["T_amb"])).all()


@pytest.mark.parametrize("z_range", [None, (0, 5), (3, 8), (4, 5)])
def test_dirichlet_side_face_kernel_level1_constants(gm, example_props, z_range):
    """GOMELT_STEP_BC_CONST on the fast kernel (owned interior by the step, the five constant faces by
    face_const_kernel) against the general kernel: interior within f32 rounding, faces exactly the constants in
    assignBCs order (cF:1568-1595), planes outside z_range untouched, inactive planes T_amb + faces."""
    import torch

    ops = gm.ops
    elements = (70, 21, 9)
    bounds = ((0.0, 14.0), (0.0, 4.2), (-1.8, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 8, nsub_planes=3)
    nx, ny, nz = lv["nodes"]
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    dT0, dS1 = _dev(T0), _dev(S1)
    bc5 = [301.0, 302.0, 303.0, 304.0, 305.0]
    res = []
    for extra in (0, ops.STEP_GENERAL_KERNEL):
        Tout = torch.full((lv["nn"],), -7.0, device="cuda")
        ops.level_step(props, grid, dT0, dS1, Tout, 2e-3, n_substrate=nsub, nz_active=nz - 2, bc5=bc5,
                       flags=ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX | extra, z_range=z_range)
        torch.cuda.synchronize()
        res.append(Tout.cpu().numpy().reshape(nz, ny, nx))
    fast, gen = res
    face = _faces_mask(nx, ny, nz)
    assert np.array_equal(fast[face], gen[face])
    assert _rel(fast[~face], gen[~face]) <= 2e-6, _rel(fast[~face], gen[~face])
    z0, z1 = z_range or (0, nz)
    assert (fast[:z0] == -7.0).all() and (fast[z1:] == -7.0).all()
    assert fast[max(z0, 1), 0, 0] == np.float32(303.0) and fast[max(z0, 1), -1, -1] == np.float32(304.0)


def test_l3_substeps_compact_faces_equal_per_substep_prolongation(gm, example_props):
    """The block-level split of the face prolongation (parents interpolated once at the face nodes, blended per
    substep: gomelt_faces_gather_f32 / gomelt_faces_blend_f32) against the per-substep interpolation."""
    import torch

    ops = gm.ops
    elements = (64, 22, 6)
    bounds = ((0.2, 0.2 + 64 * 0.02), (0.1, 0.1 + 22 * 0.02), (-0.12, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 13, nsub_planes=2)
    par = make_level((40, 20, 5), ((0.0, 1.6), (0.0, 0.8), (-0.2, 0.0)))
    rng = np.random.default_rng(2)
    Pnew = smooth_field(par, rng)
    Pold = (Pnew - 3.0 * rng.random(par["nn"])).astype(np.float32)
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    nx, ny, nz = lv["nodes"]
    n = 5
    rows = np.zeros((n, 7), np.float32)
    for i in range(n):
        rows[i] = (0.7 + 0.013 * i, 0.3, 0.0, 1, 1, 1e-5, 285.0)
    coords = [_dev(c) for c in lv["node_coords"]]
    pc = [_dev(c) for c in par["node_coords"]]
    res = []
    for compact in (True, False):
        Tin, S1w = _dev(T0), _dev(S1)
        A, B = torch.empty_like(Tin), torch.empty_like(Tin)
        tables = torch.empty(n * (nx + ny + nz), device="cuda")
        last = ops.l3_substeps(props, grid, coords, rows, Tin, A, B, S1w, tables, n_substrate=nsub,
                               flags=ops.STEP_CLAMP | ops.STEP_SKIP_FACES,
                               faces=(pc, _dev(Pnew), _dev(Pold), float(n), float(P["T_amb"])), compact_faces=compact)
        torch.cuda.synchronize()
        res.append((last.cpu().numpy(), S1w.cpu().numpy()))
    assert _rel(res[0][0], res[1][0]) <= 1e-6, _rel(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    face = _faces_mask(nx, ny, nz).ravel()
    assert np.isfinite(res[0][0]).all() and (res[0][0][face] >= np.float32(P["T_amb"])).all()


@pytest.mark.parametrize("elements,z_chunk", [((129, 37, 9), 0), ((100, 100, 10), 4), ((61, 5, 4), 0), ((150, 13, 6), 2)])
def test_fast_kernel_melt_time_bookkeeping_has_one_owner_per_node(gm, example_props, elements, z_chunk):
    """subcycleL3_Part2's call shape (SKIP_FACES + WRITE_S2 + ACCUM, S2 updated in place) on the fast kernel: S2,
    accum and max_accum (cF:3568-3578) exact on EVERY node - faces, the overlap of the shifted last tile / strip
    and chunk boundaries included - against the closed form and the general kernel."""
    import torch

    ops = gm.ops
    ex, ey, ez = elements
    bounds = ((0.0, 0.02 * ex), (0.0, 0.02 * ey), (-0.02 * ez, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 21, nsub_planes=2)
    T0 = (T0 * 1.1).astype(np.float32)
    nn, (nx, ny, nz) = lv["nn"], lv["nodes"]
    rng = np.random.default_rng(23)
    prev = rng.random(nn) > 0.5
    acc = (rng.random(nn) * 1e-3).astype(np.float32)
    mx = (rng.random(nn) * 1e-3).astype(np.float32)
    S2 = T0 >= np.float32(P["T_liquidus"])
    reset = acc * ((~prev) & S2)
    mx_ref = np.maximum(reset, mx)
    acc_ref = (acc + np.float32(1e-5) * S2 - reset).astype(np.float32)
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    coords = [_dev(c) for c in lv["node_coords"]]
    tx, ty, tz = (torch.empty(n, device="cuda") for n in (nx, ny, nz))
    v = np.array([0.01 * ex, 0.01 * ey, 0.0], np.float32)
    coef = ops.source_tables(props, grid, coords, v, 285.0, tx, ty, tz)
    outs = []
    # the fast kernel on its own, the general kernel, and the fast kernel with the hot-plane work queue
    # (gomelt_step_args_t.bk_queue: ample, and with room for 3 entries only so that most warps keep their planes)
    for extra, qwords in ((0, 0), (ops.STEP_GENERAL_KERNEL, 0), (0, 2 + 2 * (nn // 120 + 1024)), (0, 2 + 2 * 3)):
        dS2, dacc, dmx = _dev(prev.astype(np.uint8)), _dev(acc), _dev(mx)
        Tout = torch.full((nn,), -7.0, device="cuda")
        S1o = torch.empty(nn, device="cuda")
        queue = torch.full((qwords,), 12345, device="cuda", dtype=torch.int32) if qwords else None
        l0 = ops.LAUNCHES
        ops.level_step(props, grid, _dev(T0), _dev(S1), Tout, 1e-5, src=(tx, ty, tz, coef), n_substrate=nsub,
                       flags=ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_WRITE_S2 |
                       ops.STEP_ACCUM | ops.STEP_FUSED_FLUX | extra,
                       S1_out=S1o, S2_out=dS2, S2_prev=dS2, accum=dacc, max_accum=dmx, z_chunk=z_chunk, bk_queue=queue, bk_queue_force=True)
        torch.cuda.synchronize()
        if qwords and nx >= 62:  # the fast kernel took the call: step + queue kernel, which leaves the header zeroed
            assert ops.LAUNCHES - l0 == 2 and int(queue[0]) == 0 and int(queue[1]) == 0
        outs.append([t.cpu().numpy() for t in (Tout, S1o, dS2, dacc, dmx)])
    for k in range(5):  # the queue changes who does the bookkeeping, not a single bit of any output
        assert np.array_equal(outs[2][k], outs[0][k]) and np.array_equal(outs[3][k], outs[0][k])
    for T, S1o, s2, a, m in outs:
        assert np.array_equal(s2.astype(bool), S2)
        assert np.array_equal(m, mx_ref)
        assert np.array_equal(a, acc_ref)
    face = _faces_mask(nx, ny, nz).ravel()
    assert np.array_equal(outs[0][1], outs[1][1])
    assert _rel(outs[0][0][~face], outs[1][0][~face]) <= 2e-6
    assert (outs[0][0][face] == -7.0).all()


@pytest.mark.parametrize("elements,nsub_planes,z_chunk", [((129, 37, 12), 3, 0), ((100, 100, 10), 0, 4), ((61, 5, 9), 9, 0)])
def test_fast_kernel_in_place_state_on_a_mostly_cold_field(gm, example_props, elements, nsub_planes, z_chunk):
    """The fast kernel's in-place state update (S1_out == S1, what the steppers pass: a node's state is stored only
    when it changed) on the kind of field a run has - mostly below the solidus, one melt pool - with fractional,
    negative-zero, at-threshold and powder-in-substrate states.  In place vs separate S1_out: T and S1' identical;
    against the general kernel: T to rounding, S1' exact; against the oracle: 1e-5, S1' exact."""
    import torch

    ops = gm.ops
    ex, ey, ez = elements
    bounds = ((0.0, 0.02 * ex), (0.0, 0.02 * ey), (-0.02 * ez, 0.0))
    P = cF.SetupProperties(example_props)
    lv = make_level(elements, bounds)
    nx, ny, nz = lv["nodes"]
    rng = np.random.default_rng(21)
    x, y, z = lv["node_coords"]
    r2 = ((x[None, None, :] - 0.3 * x[-1]) / 0.2) ** 2 + ((y[None, :, None] - 0.5 * y[-1]) / 0.15) ** 2 + (z[:, None, None] / 0.08) ** 2
    T0 = (320.0 + 600.0 * rng.random((nz, ny, nx)) + 2400.0 * np.exp(-r2)).astype(np.float32).ravel()
    assert (T0 >= P["T_liquidus"]).sum() > 20 and (T0 < P["T_solidus"]).mean() > 0.8
    S1 = (rng.random(lv["nn"]) > 0.5).astype(np.float32)
    odd = rng.choice(lv["nn"], 200, replace=False)
    S1[odd[:80]] = rng.random(80).astype(np.float32)        # fractional states (Level 1 after interpolation)
    S1[odd[80:120]] = np.float32(-0.0)
    S1[odd[120:160]] = np.float32(0.499)
    S1[odd[160:]] = np.float32(0.4990001)
    nsub = nsub_planes * nx * ny
    v = np.array([0.3 * x[-1], 0.5 * y[-1], 0.0], np.float32)
    dt = 1e-5
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    dT0 = _dev(T0)
    coords = [_dev(c) for c in lv["node_coords"]]
    tx, ty, tz = (torch.empty(n, device="cuda") for n in (nx, ny, nz))
    src = (tx, ty, tz, ops.source_tables(props, grid, coords, v, 285.0, tx, ty, tz))
    base = ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX

    def run(extra, in_place):
        dS1 = _dev(S1)
        Tout = torch.full((lv["nn"],), -7.0, device="cuda")
        S1o = dS1 if in_place else torch.full((lv["nn"],), -3.0, device="cuda")
        ops.level_step(props, grid, dT0, dS1, Tout, dt, src=src, n_substrate=nsub, S1_out=S1o, flags=base | extra,
                       z_chunk=z_chunk)
        torch.cuda.synchronize()
        return Tout.cpu().numpy(), S1o.cpu().numpy()

    Tc, Sc = run(0, False)
    Ti, Si = run(0, True)
    Tih, Sih = run(ops.STEP_NO_COLD_PLANES, True)  # general property selects on every plane
    Tg, Sg = run(ops.STEP_GENERAL_KERNEL, False)
    assert np.array_equal(Tc, Ti) and np.array_equal(Tc, Tih) and np.array_equal(Sih, Sc)
    assert np.array_equal(Si, Sc)  # -0.0 == 0.0: an unchanged node may keep its bits in place
    assert np.array_equal(Sc, Sg)
    face = _faces_mask(nx, ny, nz).ravel()
    assert (Tc[face] == -7.0).all()
    assert _rel(Tc[~face], Tg[~face]) <= 2e-6
    Tref, S1ref, _ = _oracle_step(P, lv, T0, S1, nsub, dt, v, 285.0)
    Tref = np.maximum(np.float32(P["T_amb"]), Tref)
    assert _rel(Tc[~face], Tref[~face]) <= RTOL
    assert np.array_equal(Sc, S1ref)


@pytest.mark.parametrize("state", ["float32", "uint8"])
def test_host_block_pipeline_equals_device_blocks(gm, example_props, state):
    """hostpipe.HostBlockPipeline (host buffers, two blocks in flight on three streams) returns, for every block,
    exactly what the same gomelt_l3_substeps_f32 call returns on device-resident fields - with the state as float32 or
    as bytes on the host side and on the wire."""
    import torch

    sdt = torch.float32 if state == "float32" else torch.uint8

    ops = gm.ops
    elements = (75, 23, 7)
    bounds = ((0.0, 1.5), (0.0, 0.46), (-0.14, 0.0))
    P, lv, T0, S1, nsub = _setup(gm, example_props, elements, bounds, 41, nsub_planes=2)
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(lv["nodes"], lv["h"])
    nx, ny, nz = lv["nodes"]
    coords = [_dev(c) for c in lv["node_coords"]]
    n = 5
    flags = ops.STEP_CLAMP  # natural boundaries: every node of every substep is defined by the block's own input
    pipe = gm.hostpipe.HostBlockPipeline(ops, props, grid, coords, depth=2, n_rows=n, n_substrate=nsub, flags=flags)
    rng = np.random.default_rng(3)
    jobs = []
    for j in range(5):
        rows = np.zeros((n, 7), np.float32)
        for i in range(n):
            rows[i] = (0.5 + 0.05 * j + 0.01 * i, 0.23, 0.0, 1, 1, 1e-5, 285.0)
        Tj = (T0 * (0.6 + 0.1 * j)).astype(np.float32)
        Sj = (rng.random(lv["nn"]) > 0.4).astype(np.float32)
        hT, hS = torch.as_tensor(Tj).pin_memory(), torch.as_tensor(Sj).to(sdt).pin_memory()
        oT, oS = torch.empty(lv["nn"]).pin_memory(), torch.empty(lv["nn"], dtype=sdt).pin_memory()
        pipe.submit(hT, hS, rows, oT, oS)
        jobs.append((rows, Tj, Sj, oT, oS))
    pipe.drain()
    tables = torch.empty(n * (nx + ny + nz), device="cuda")
    for rows, Tj, Sj, oT, oS in jobs:
        A, S = _dev(Tj), _dev(Sj)
        B = torch.empty_like(A)
        last = ops.l3_substeps(props, grid, coords, rows, A, B, A, S, tables, n_substrate=nsub, flags=flags)
        torch.cuda.synchronize()
        assert np.array_equal(oT.numpy(), last.cpu().numpy())
        assert np.array_equal(oS.numpy().astype(np.float32), S.cpu().numpy())
