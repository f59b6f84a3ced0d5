"""K1 at BASELINE.json's full sizes (10.26 M-node Level-3 window, 25 M-node Level-1 slab), where the oracle
cannot run the whole field: size-independent properties of the explicit step plus an oracle check on cut-outs.

* a constant field is a fixed point of the conduction operator, bit for bit (every Haar difference is exactly 0),
  whatever the state / conductivity distribution;
* the update is linear in the load (properties depend on T0 only): T(2P) - T(0) = 2 (T(P) - T(0));
* z-chunked launches are bit-identical to the single-chunk launch;
* the fast kernel agrees with the general kernel on the same call;
* the stencil has radius 1, so the oracle run on a cut-out box reproduces the big run on the cut-out's interior
  (nodes at least one cell away from the cut): checked under the laser, at a corner tile and on the top plane.
"""
import numpy as np
import pytest

from oracle import computeFunctions as cF
from oracle.util import make_level

pytestmark = pytest.mark.gpu
RTOL = 1e-5
L3_NODES = (513, 513, 39)   # BASELINE.json configs[1]
L1_NODES = (1001, 1001, 25)  # one GPU's slab of configs[4]
H3 = 0.02


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def _fields(torch, nodes, seed, hot=2300.0):
    nx, ny, nz = nodes
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.linspace(-1, 1, nx, device="cuda")[None, None, :]
    y = torch.linspace(-1, 1, ny, device="cuda")[None, :, None]
    z = torch.linspace(-1, 0, nz, device="cuda")[:, None, None]
    T = 300.0 + hot * torch.exp(-14.0 * (x * x + y * y) + 3.0 * z) + 40.0 * torch.rand(nz, ny, nx, device="cuda", generator=g)
    S1 = (torch.rand(nz, ny, nx, device="cuda", generator=g) > 0.4).float()
    return T.reshape(-1).contiguous(), S1.reshape(-1).contiguous()


def _l3_call(gm, props, grid, T0, S1, Tout, S1o, src, extra=0, z_chunk=0, nsub=0):
    ops = gm.ops
    ops.level_step(props, grid, T0, S1, Tout, 1e-5, src=src, n_substrate=nsub, S1_out=S1o, z_chunk=z_chunk,
                   flags=ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX | extra)


@pytest.fixture(scope="module")
def l3(gm, example_props):
    import torch

    P = cF.SetupProperties(example_props)
    nx, ny, nz = L3_NODES
    props = gm._lib.make_props(P)
    grid = gm._lib.make_grid(L3_NODES, (H3, H3, H3))
    coords_h = [np.linspace(0, (nx - 1) * H3, nx, dtype=np.float32), np.linspace(0, (ny - 1) * H3, ny, dtype=np.float32),
                np.linspace(-(nz - 1) * H3, 0, nz, dtype=np.float32)]
    coords = [torch.as_tensor(c).cuda() for c in coords_h]
    T0, S1 = _fields(torch, L3_NODES, 0)
    v = np.array([0.5 * (nx - 1) * H3, 0.5 * (ny - 1) * H3, 0.0], np.float32)
    tx, ty, tz = (torch.empty(n, device="cuda") for n in L3_NODES)
    coef = gm.ops.source_tables(props, grid, coords, v, 285.0, tx, ty, tz)
    nsub = 2 * nx * ny
    Tout = torch.full_like(T0, -7.0)
    S1o = torch.empty_like(S1)
    _l3_call(gm, props, grid, T0, S1, Tout, S1o, (tx, ty, tz, coef), nsub=nsub)
    torch.cuda.synchronize()
    return dict(P=P, props=props, grid=grid, coords_h=coords_h, T0=T0, S1=S1, src=(tx, ty, tz, coef), v=v, nsub=nsub,
                Tout=Tout, S1o=S1o)


def test_constant_field_is_a_fixed_point_bitwise(gm, example_props):
    import torch

    P = cF.SetupProperties(example_props)
    props = gm._lib.make_props(P)
    for nodes, flags, bc in ((L3_NODES, gm.ops.STEP_SKIP_FACES | gm.ops.STEP_CLAMP, None),
                             (L1_NODES, gm.ops.STEP_BC_CONST, [1234.5] * 5)):
        grid = gm._lib.make_grid(nodes, (H3, H3, H3))
        nn = nodes[0] * nodes[1] * nodes[2]
        _, S1 = _fields(torch, nodes, 3)
        T0 = torch.full((nn,), 1234.5, device="cuda")
        Tout = torch.full((nn,), 1234.5, device="cuda")  # faces are not written under SKIP_FACES
        gm.ops.level_step(props, grid, T0, S1, Tout, 1e-5, flags=flags, bc5=bc)   # no source, no surface flux
        torch.cuda.synchronize()
        assert bool((Tout == 1234.5).all()), nodes


def test_update_is_linear_in_the_load(gm, l3):
    import torch

    tx, ty, tz, coef = l3["src"]
    outs = []
    for mult in (0.0, 1.0, 2.0):
        Tout, S1o = torch.empty_like(l3["T0"]), torch.empty_like(l3["S1"])
        gm.ops.level_step(l3["props"], l3["grid"], l3["T0"], l3["S1"], Tout, 1e-5, src=(tx, ty, tz, coef * mult),
                          n_substrate=l3["nsub"],
                          flags=gm.ops.STEP_SKIP_FACES | gm.ops.STEP_CLAMP | gm.ops.STEP_FUSED_FLUX)  # stepGOMELT's Level-3 shape
        outs.append(Tout)
    torch.cuda.synchronize()
    d1, d2 = (outs[1] - outs[0]).double(), (outs[2] - outs[0]).double()
    nx, ny, nz = L3_NODES
    inner = torch.zeros(nz, ny, nx, dtype=torch.bool, device="cuda")
    inner[1:, 1:-1, 1:-1] = True
    inner = inner.reshape(-1)
    scale = float(d1[inner].abs().max())
    assert scale > 1.0  # the laser does heat the window
    assert float(outs[0][inner].min()) > 298.15 + 1e-3  # the clamp is inactive, so the step is affine in the load
    # both differences carry the f32 rounding of T itself (~1e-7 * 2600 K)
    assert float((d2[inner] - 2.0 * d1[inner]).abs().max()) <= 2e-3 + 1e-5 * scale


def test_z_chunks_and_general_kernel_agree_at_full_size(gm, l3):
    import torch

    nx, ny, nz = L3_NODES
    face = np.zeros((nz, ny, nx), bool)
    face[0] = True
    face[:, 0] = face[:, -1] = True
    face[:, :, 0] = face[:, :, -1] = True
    face = torch.as_tensor(face.reshape(-1)).cuda()
    ref_T, ref_S = l3["Tout"], l3["S1o"]
    assert bool((ref_T[face] == -7.0).all())
    for zc in (13, 7):
        Tout, S1o = torch.full_like(ref_T, -7.0), torch.empty_like(ref_S)
        _l3_call(gm, l3["props"], l3["grid"], l3["T0"], l3["S1"], Tout, S1o, l3["src"], z_chunk=zc, nsub=l3["nsub"])
        assert torch.equal(Tout, ref_T) and torch.equal(S1o, ref_S), zc
    Tout, S1o = torch.full_like(ref_T, -7.0), torch.empty_like(ref_S)
    _l3_call(gm, l3["props"], l3["grid"], l3["T0"], l3["S1"], Tout, S1o, l3["src"], extra=gm.ops.STEP_GENERAL_KERNEL,
             nsub=l3["nsub"])
    torch.cuda.synchronize()
    assert torch.equal(S1o, ref_S)
    rel = ((Tout - ref_T).abs() / ref_T.abs().clamp_min(1.0))[~face].max().item()
    assert rel <= 2e-6, rel


@pytest.mark.parametrize("box", ["laser", "corner", "edge_tile"])
def test_oracle_on_cut_outs_of_the_full_window(gm, l3, box):
    """Oracle (whole substep: state, source, surface flux, solve, clamp) on a cut-out with all 39 planes."""
    nx, ny, nz = L3_NODES
    i0, j0, w = {"laser": (236, 240, 40), "corner": (0, 0, 36), "edge_tile": (nx - 45, 300, 44)}[box]
    i1, j1 = min(i0 + w, nx - 1), min(j0 + w, ny - 1)
    cx, cy, cz = l3["coords_h"]
    sub = make_level((i1 - i0, j1 - j0, nz - 1), ((float(cx[i0]), float(cx[i1])), (float(cy[j0]), float(cy[j1])),
                                                 (float(cz[0]), float(cz[-1]))))
    sub["node_coords"] = [cx[i0:i1 + 1].copy(), cy[j0:j1 + 1].copy(), cz.copy()]  # the big grid's own float32 coordinates
    cut = lambda t: t.reshape(nz, ny, nx)[:, j0:j1 + 1, i0:i1 + 1].cpu().numpy().reshape(-1)
    T0, S1 = cut(l3["T0"]), cut(l3["S1"])
    P = l3["P"]
    nsub_cut = 2 * (i1 - i0 + 1) * (j1 - j0 + 1)
    S1n, _, k, rc = cF.computeStateProperties(T0, S1, P, nsub_cut)
    F = cF.computeSourcesL3(sub, l3["v"], (0, sub["ne"], 0, 0, sub["nn"]), P, 285.0)
    F = cF.computeConvRadBC(sub, T0, sub["ne"], sub["nn"], P, F)
    Tref = np.maximum(np.float32(P["T_amb"]), cF.solveMatrixFreeFE(sub, sub["nn"], sub["ne"], k, rc, 1e-5, T0, F, 0))
    got, gotS = cut(l3["Tout"]), cut(l3["S1o"])
    sx, sy = i1 - i0 + 1, j1 - j0 + 1
    keep = np.zeros((nz, sy, sx), bool)
    keep[1:, 1:-1, 1:-1] = True   # one cell away from the cut; plane 0 is a Dirichlet face of the big window
    keep = keep.reshape(-1)
    assert _rel(got[keep], Tref[keep]) <= RTOL, _rel(got[keep], Tref[keep])
    assert np.array_equal(gotS, S1n)  # S1' is node-local: exact on the whole cut-out, faces included
    above = np.float32(P["T_liquidus"])
    near = np.abs(Tref - above) <= 8 * np.spacing(above)
    assert np.array_equal((got >= above)[keep & ~near], (Tref >= above)[keep & ~near])
    if box == "laser":
        assert Tref[keep].max() > 2000.0


def test_level1_slab_size_cut_out_and_general_kernel(gm, example_props):
    """The 25-plane Level-1 slab (one GPU's share of configs[4]): dwell-step shape (BC_CONST, fused flux, no clamp)."""
    import torch

    P = cF.SetupProperties(example_props)
    props = gm._lib.make_props(P)
    nx, ny, nz = L1_NODES
    h = 0.2
    grid = gm._lib.make_grid(L1_NODES, (h, h, h))
    T0, S1 = _fields(torch, L1_NODES, 5, hot=1500.0)
    bc5 = [298.15] * 5
    outs = []
    for extra in (0, gm.ops.STEP_GENERAL_KERNEL):
        Tout = torch.full_like(T0, -7.0)
        gm.ops.level_step(props, grid, T0, S1, Tout, 2e-3, n_substrate=3 * nx * ny, bc5=bc5,
                          flags=gm.ops.STEP_BC_CONST | gm.ops.STEP_FUSED_FLUX | extra)
        outs.append(Tout)
    torch.cuda.synchronize()
    rel = ((outs[0] - outs[1]).abs() / outs[1].abs().clamp_min(1.0)).max().item()
    assert rel <= 2e-6, rel
    # oracle on a cut-out around the hot spot
    i0, j0, w = 480, 478, 40
    cx = np.linspace(0, (nx - 1) * h, nx, dtype=np.float32)
    cz = np.linspace(-(nz - 1) * h, 0, nz, dtype=np.float32)
    sub = make_level((w, w, nz - 1), ((float(cx[i0]), float(cx[i0 + w])), (float(cx[j0]), float(cx[j0 + w])),
                                      (float(cz[0]), float(cz[-1]))))
    cut = lambda t: t.reshape(nz, ny, nx)[:, j0:j0 + w + 1, i0:i0 + w + 1].cpu().numpy().reshape(-1)
    T0c, S1c = cut(T0), cut(S1)
    _, _, k, rc = cF.computeStateProperties(T0c, S1c, P, 3 * (w + 1) * (w + 1))
    F = cF.computeConvRadBC(sub, T0c, sub["ne"], sub["nn"], P, np.zeros(sub["nn"], np.float32))
    Tref = cF.solveMatrixFreeFE(sub, sub["nn"], sub["ne"], k, rc, 2e-3, T0c, F, 0)
    keep = np.zeros((nz, w + 1, w + 1), bool)
    keep[1:, 1:-1, 1:-1] = True
    keep = keep.reshape(-1)
    got = cut(outs[0])
    assert _rel(got[keep], Tref[keep]) <= RTOL, _rel(got[keep], Tref[keep])


# ------------------------------------------------------------------------------------------------------------
# K3 at C2 size: the marching projection kernel (nested windows) against the tile kernel, plus a property the
# domain offers at any size - the parent's shape functions are a partition of unity, so the projected MASS vector
# sums to -scale * integral(rho*cp A) and the projected GRAD vector sums to 0
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ratio", [2, 5, 10])
def test_projection_march_vs_tile_at_full_size(gm, example_props, ratio):
    import torch

    P = cF.SetupProperties(example_props)
    cf = gm.computeFunctions
    ne = (500, 500, 40)
    hf = 0.02
    hc = hf * ratio
    fine = make_level(ne, ((4.0, 4.0 + ne[0] * hf), (4.0, 4.0 + ne[1] * hf), (-ne[2] * hf, 0.0)))
    npar = (int(np.ceil(ne[0] / ratio)) + 8 * 2, int(np.ceil(ne[1] / ratio)) + 8 * 2, int(np.ceil(ne[2] / ratio)) + 2)
    parent = make_level(npar, ((4.0 - 8 * hc, 4.0 - 8 * hc + npar[0] * hc), (4.0 - 8 * hc, 4.0 - 8 * hc + npar[1] * hc),
                               (-npar[2] * hc, 0.0)))
    cells = cf._pair_cells(fine, parent)
    assert [int(v) for v in cells["rmax"]] == [ratio] * 3
    Tf, S1 = _fields(torch, fine["nodes"], 5)
    g = torch.Generator(device="cuda").manual_seed(6)
    Tp0 = 60.0 * torch.rand(fine["nn"], device="cuda", generator=g) - 30.0
    Tp1 = Tp0 + 4.0 * torch.rand(fine["nn"], device="cuda", generator=g)
    props = gm._lib.make_props(P)
    dt = 1e-5
    out = {}
    for how in (True, "tile"):
        V = torch.zeros(parent["nn"], device="cuda")
        cf._project(cells, Tp0, None, V, mode=0, coef_from=(props, Tf, S1, 0), tiled=how)
        g0 = V.clone()
        cf._project(cells, Tp1, None, V, mode=1, scale=1.0 / dt, A2=Tp0, coef_from=(props, Tf, S1, 0), tiled=how)
        out[how] = (g0, V - g0)
    for q in (0, 1):
        a, b = out[True][q].double(), out["tile"][q].double()
        assert float((a - b).abs().max() / b.abs().max()) <= 3e-5, (ratio, q)
    # partition of unity: sum_c Nc = 1, sum_c grad Nc = 0
    g0, m1 = out[True]
    assert abs(float(g0.double().sum())) <= 1e-4 * float(g0.double().abs().sum())
    _, _, _, rc = cf.computeStateProperties(Tf, S1, P, 0)
    # integral of (element-mean rho*cp) * (trilinear dA) with the 2x2x2 rule = wq * cbar * sum_q N[q,:] . dA = V_e * cbar * mean_8(dA)
    nx, ny, nz = fine["nodes"]
    dA = (Tp1 - Tp0).double().reshape(nz, ny, nx)
    rc3 = rc.double().reshape(nz, ny, nx)

    def mean8(f):
        return (f[:-1, :-1, :-1] + f[:-1, :-1, 1:] + f[:-1, 1:, :-1] + f[:-1, 1:, 1:] + f[1:, :-1, :-1] + f[1:, :-1, 1:]
                + f[1:, 1:, :-1] + f[1:, 1:, 1:]) / 8.0

    want = -(1.0 / dt) * float((mean8(rc3) * mean8(dA)).sum()) * hf ** 3
    assert abs(float(m1.double().sum()) - want) <= 2e-5 * abs(want), (float(m1.double().sum()), want)


def test_melt_time_work_queue_at_full_size(gm, example_props):
    """The corrector-substep shape (SKIP_FACES + WRITE_S2 + ACCUM, in place) on the 10.26 M-node window with one melt pool:
    with the hot-plane work queue (what the steppers pass; used by default at this size) every output equals the
    in-kernel bookkeeping bit for bit, and S2 / accum / max_accum equal the closed form."""
    import torch

    ops = gm.ops
    P = cF.SetupProperties(example_props)
    props = gm._lib.make_props(P)
    nx, ny, nz = L3_NODES
    nn = nx * ny * nz
    grid = gm._lib.make_grid(L3_NODES, (H3, H3, H3))
    T0, S1 = _fields(torch, L3_NODES, 3, hot=2600.0)
    g = torch.Generator(device="cuda").manual_seed(4)
    prev = (torch.rand(nn, device="cuda", generator=g) > 0.5).to(torch.uint8)
    acc = torch.rand(nn, device="cuda", generator=g) * 1e-3
    mx = torch.rand(nn, device="cuda", generator=g) * 1e-3
    tx, ty, tz = (torch.rand(n, device="cuda", generator=g) for n in L3_NODES)
    outs = []
    for queue in (None, torch.zeros(2 + 2 * (nn // 120 + 1024), device="cuda", dtype=torch.int32)):
        dS2, dacc, dmx, dS1 = prev.clone(), acc.clone(), mx.clone(), S1.clone()
        Tout = torch.full((nn,), -7.0, device="cuda")   # (the step leaves the Dirichlet faces alone)
        l0 = ops.LAUNCHES
        ops.level_step(props, grid, T0, dS1, Tout, 1e-5, src=(tx, ty, tz, 1e-3), n_substrate=2 * nx * ny,
                       flags=ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_WRITE_S2 | ops.STEP_ACCUM |
                       ops.STEP_FUSED_FLUX, S1_out=dS1, S2_out=dS2, S2_prev=dS2, accum=dacc, max_accum=dmx, bk_queue=queue)
        torch.cuda.synchronize()
        assert ops.LAUNCHES - l0 == (1 if queue is None else 2)
        outs.append((Tout, dS1, dS2, dacc, dmx))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    S2 = T0 >= float(np.float32(P["T_liquidus"]))
    assert 100 < int(S2.sum()) < nn // 50   # one pool
    reset = acc * ((prev == 0) & S2)
    assert torch.equal(outs[1][2].bool(), S2)
    assert torch.equal(outs[1][4], torch.maximum(reset, mx))
    assert torch.equal(outs[1][3], (acc + torch.where(S2, torch.full_like(acc, 1e-5), torch.zeros_like(acc))) - reset)
