"""Oracle (test infrastructure): per-level FE primitives of the reference, NumPy float32.

Restates, in the reference's own operation order:
  createMesh3D cF:32-64 - convert2XYZ cF:645-689 - computeQuad3dFemShapeFunctions_jax cF:692-775
  computeQuad2dFemShapeFunctions_jax cF:778-855 - getSampleCoords cF:3135-3148
  getQuadratureCoords cF:3151-3166 - solveMatrixFreeFE cF:582-642
  computeStateProperties cF:2567-2614 - computeConvRadBC cF:2207-2301
  computeSourceFunction_jax cF:991-1025 - computeSourcesL3 cF:2960-3012 - bincount cF:1623-1639
"""
import numpy as np

from . import config


def _f(x):
    return config.f(x)


def linspace(start, stop, num):
    """jnp.linspace in the computation dtype (jax 0.4.16 `_linspace`): iota/div, then
    start*(1-step) + stop*step, endpoint concatenated.  Used by createMesh3D cF:47."""
    FDT = config.FDT
    num = int(num)
    div = num - 1
    start = FDT(start)
    stop = FDT(stop)
    if num == 1:
        return np.array([start], dtype=FDT)
    step = np.arange(div, dtype=FDT) / FDT(div)
    out = start * (FDT(1) - step) + stop * step
    return np.concatenate([out, np.array([stop], dtype=FDT)]).astype(FDT)


def createMesh3D(x, y, z):
    """cF:32-64: node coordinates (3 x 1-D) and per-axis hex8 connectivity (ne_d, 8)."""
    nx, ny, nz = [linspace(*axis) for axis in (x, y, z)]
    cx0 = np.arange(0, x[2] - 1).reshape(-1, 1)
    cx1 = np.arange(1, x[2]).reshape(-1, 1)
    nconn_x = np.concatenate([cx0, cx1, cx1, cx0, cx0, cx1, cx1, cx0], axis=1)
    cy0 = np.arange(0, y[2] - 1).reshape(-1, 1)
    cy1 = np.arange(1, y[2]).reshape(-1, 1)
    nconn_y = np.concatenate([cy0, cy0, cy1, cy1, cy0, cy0, cy1, cy1], axis=1)
    cz0 = np.arange(0, z[2] - 1).reshape(-1, 1)
    cz1 = np.arange(1, z[2]).reshape(-1, 1)
    nconn_z = np.concatenate([cz0, cz0, cz0, cz0, cz1, cz1, cz1, cz1], axis=1)
    return [nx, ny, nz], [nconn_x, nconn_y, nconn_z]


def convert2XYZ(i, ne_x, ne_y, nn_x, nn_y):
    """cF:645-689: element id(s) -> (ix, iy, iz, idx[8(,ne)]); x-fastest numbering."""
    i = np.asarray(i)
    ne_x, ne_y, nn_x, nn_y = int(ne_x), int(ne_y), int(nn_x), int(nn_y)
    ne_xy = ne_x * ne_y
    nn_xy = nn_x * nn_y
    iz = i // ne_xy
    iy = (i // ne_x) - iz * ne_y
    ix = i % ne_x
    base = ix + iy * nn_x + iz * nn_xy
    dx, dy, dz = 1, nn_x, nn_xy
    idx = np.array(
        [
            base,
            base + dx,
            base + dx + dy,
            base + dy,
            base + dz,
            base + dx + dz,
            base + dx + dy + dz,
            base + dy + dz,
        ]
    )
    return ix, iy, iz, idx


_KSI = np.array([-1, 1, 1, -1, -1, 1, 1, -1])
_ETA = np.array([-1, -1, 1, 1, -1, -1, 1, 1])
_ZETA = np.array([-1, -1, -1, -1, 1, 1, 1, 1])


def computeQuad3dFemShapeFunctions(coords):
    """cF:692-775: N (8q,8a), dNdx (8q,8a,3), wq (8,1) for one hex8, 2x2x2 Gauss."""
    FDT = config.FDT
    coords = np.asarray(coords, dtype=FDT)
    ksi_i = _KSI.astype(FDT)
    eta_i = _ETA.astype(FDT)
    zeta_i = _ZETA.astype(FDT)
    inv_s3 = FDT(1) / np.sqrt(FDT(3))
    ksi_q = inv_s3 * ksi_i
    eta_q = inv_s3 * eta_i
    zeta_q = inv_s3 * zeta_i
    _ksi = FDT(1) + ksi_q[:, None] @ ksi_i[None, :]
    _eta = FDT(1) + eta_q[:, None] @ eta_i[None, :]
    _zeta = FDT(1) + zeta_q[:, None] @ zeta_i[None, :]
    N = FDT(1 / 8) * _ksi * _eta * _zeta
    dNdksi = FDT(1 / 8) * ksi_i[None, :] * _eta * _zeta
    dNdeta = FDT(1 / 8) * eta_i[None, :] * _ksi * _zeta
    dNdzeta = FDT(1 / 8) * zeta_i[None, :] * _ksi * _eta
    dxdksi = dNdksi @ coords[:, 0]
    dydeta = dNdeta @ coords[:, 1]
    dzdzeta = dNdzeta @ coords[:, 2]
    Jinv = np.array(
        [
            [FDT(1) / dxdksi[0], 0, 0],
            [0, FDT(1) / dydeta[0], 0],
            [0, 0, FDT(1) / dzdzeta[0]],
        ],
        dtype=FDT,
    )
    detJ = FDT(dxdksi[0] * dydeta[0] * dzdzeta[0])  # det of the diagonal J (cF:773)
    dNdx = np.zeros((8, 8, 3), dtype=FDT)
    wq = np.zeros((8, 1), dtype=FDT)
    for q in range(8):
        dN_dxi = np.stack([dNdksi[q], dNdeta[q], dNdzeta[q]], axis=1)
        dNdx[q] = dN_dxi @ Jinv
        wq[q] = detJ * FDT(1)
    return N.astype(FDT), dNdx, wq


def computeQuad2dFemShapeFunctions(coords):
    """cF:778-855: N (4q,4a), dNdx (4,4,2), wq (4,1) of the top quad face (coords[4:])."""
    FDT = config.FDT
    coords = np.asarray(coords, dtype=FDT)
    ksi_i = np.array([-1, 1, 1, -1], dtype=FDT)
    eta_i = np.array([-1, -1, 1, 1], dtype=FDT)
    inv_s3 = FDT(1) / np.sqrt(FDT(3))
    ksi_q = inv_s3 * ksi_i
    eta_q = inv_s3 * eta_i
    _ksi = FDT(1) + ksi_q[:, None] @ ksi_i[None, :]
    _eta = FDT(1) + eta_q[:, None] @ eta_i[None, :]
    N = FDT(1 / 4) * _ksi * _eta
    dNdksi = FDT(1 / 4) * ksi_i[None, :] * _eta
    dNdeta = FDT(1 / 4) * eta_i[None, :] * _ksi
    dxdksi = dNdksi @ coords[4:, 0]
    dydeta = dNdeta @ coords[4:, 1]
    Jinv = np.array([[FDT(1) / dxdksi[0], 0], [0, FDT(1) / dydeta[0]]], dtype=FDT)
    detJ = FDT(dxdksi[0] * dydeta[0])
    dNdx = np.zeros((4, 4, 2), dtype=FDT)
    wq = np.zeros((4, 1), dtype=FDT)
    for q in range(4):
        dN_dxi = np.stack([dNdksi[q], dNdeta[q]], axis=1)
        dNdx[q] = dN_dxi @ Jinv
        wq[q] = detJ
    return N.astype(FDT), dNdx, wq


def getSampleCoords(Level):
    """cF:3135-3148: corner coordinates (8,3) of element 0."""
    x = Level["node_coords"][0][Level["connect"][0][0, :]].reshape(-1, 1)
    y = Level["node_coords"][1][Level["connect"][1][0, :]].reshape(-1, 1)
    z = Level["node_coords"][2][Level["connect"][2][0, :]].reshape(-1, 1)
    return np.concatenate([x, y, z], axis=1)


def getQuadratureCoords(Level, ix, iy, iz, Nf):
    """cF:3151-3166, vectorised over elements: returns x, y, z of shape (ne, 8q)."""
    cx = Level["node_coords"][0][Level["connect"][0][ix, :]]  # (ne, 8)
    cy = Level["node_coords"][1][Level["connect"][1][iy, :]]
    cz = Level["node_coords"][2][Level["connect"][2][iz, :]]
    return cx @ Nf.T, cy @ Nf.T, cz @ Nf.T


def bincount(N, D, nn):
    """cF:1623-1639 / jnp.bincount: float32 scatter-add applied in (e, a) row-major order
    (np.bincount would accumulate in float64, which the reference dtype never does)."""
    out = np.zeros(int(nn), dtype=config.FDT)
    np.add.at(out, np.asarray(N).reshape(-1), np.asarray(D, dtype=config.FDT).reshape(-1))
    return out


def _level_dims(Level):
    return (
        int(Level["elements"][0]),
        int(Level["elements"][1]),
        int(Level["nodes"][0]),
        int(Level["nodes"][1]),
    )


def solveMatrixFreeFE(Level, nn, ne, k, rhocp, dt, T, Fc, Corr, chunk=1 << 18):
    """cF:582-642: explicit update (sum_e LHSe T_e + F + Corr) / sum_e Me with element-mean
    k and rhocp; LHSe = diag(Me) - Ke, Me = sum_axis0(NTN*wq*m), Ke = BTB*k*wq.

    Nodes not touched by any of the first ``ne`` elements come out 0/0 = NaN, as in the
    reference (they are overwritten by substitute_Tbar, cF:2183)."""
    FDT = config.FDT
    nn, ne = int(nn), int(ne)
    coords = getSampleCoords(Level)
    N, dNdx, wq = computeQuad3dFemShapeFunctions(coords)
    wq = wq[0][0]
    NTN = N.T @ N
    BTB = np.zeros((8, 8), dtype=FDT)
    for idim in range(3):
        BTB += dNdx[:, :, idim].T @ dNdx[:, :, idim]
    dt = FDT(dt)
    k = np.asarray(k, dtype=FDT)
    rhocp = np.asarray(rhocp, dtype=FDT)
    T = np.asarray(T, dtype=FDT)
    ne_x, ne_y, nn_x, nn_y = _level_dims(Level)
    newT = np.zeros(nn, dtype=FDT)
    newM = np.zeros(nn, dtype=FDT)
    NTNw = NTN * wq
    for e0 in range(0, ne, chunk):
        e = np.arange(e0, min(ne, e0 + chunk))
        _, _, _, idx = convert2XYZ(e, ne_x, ne_y, nn_x, nn_y)
        idx = idx.T  # (ne_c, 8)
        kvec = k[idx].mean(axis=1, dtype=FDT)
        mvec = rhocp[idx].mean(axis=1, dtype=FDT) / dt
        Me = (NTNw[None, :, :] * mvec[:, None, None]).sum(axis=1, dtype=FDT)  # (ne_c, 8)
        Ke = BTB[None, :, :] * kvec[:, None, None] * wq
        LHSe = -Ke
        ar = np.arange(8)
        LHSe[:, ar, ar] += Me
        aT = np.matmul(LHSe, T[idx][:, :, None])[:, :, 0]
        np.add.at(newT, idx.reshape(-1), aT.reshape(-1))
        np.add.at(newM, idx.reshape(-1), Me.reshape(-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        return ((newT + Fc + Corr) / newM).astype(FDT)


def f32props(properties):
    """Inside jax.jit every scalar leaf of `properties` is a float32 tracer."""
    out = {}
    for key, v in properties.items():
        if isinstance(v, (int, float, np.floating, np.integer)) and not isinstance(v, bool):
            out[key] = config.FDT(v)
        else:
            out[key] = v
    return out


def computeStateProperties(T, S1, properties, Level_nodes_substrate):
    """cF:2567-2614: S1 (f32 0/1), S2 (bool), k [W/mm K], rhocp [J/mm^3 K]."""
    FDT = config.FDT
    p = f32props(properties)
    T = np.asarray(T, dtype=FDT)
    S2 = T >= p["T_liquidus"]
    S3 = (T > p["T_solidus"]) & (T < p["T_liquidus"])
    S1 = (FDT(1.0) * ((np.asarray(S1) > 0.499) | S2)).astype(FDT)
    S1[: int(Level_nodes_substrate)] = 1
    S2f = S2.astype(FDT)
    S3f = S3.astype(FDT)
    k_powder = (1 - S1) * (1 - S2f) * p["k_powder"]
    k_bulk = S1 * (1 - S2f) * (p["k_bulk_coeff_a1"] * T + p["k_bulk_coeff_a0"])
    k_fluid = S2f * p["k_fluid_coeff_a0"]
    k = (k_powder + k_bulk + k_fluid) / FDT(1000)
    cp_solid = (1 - S2f) * (1 - S3f) * (p["cp_solid_coeff_a1"] * T + p["cp_solid_coeff_a0"])
    cp_mushy = S3f * p["cp_mushy"]
    cp_fluid = S2f * p["cp_fluid"]
    rhocp = p["rho"] * (cp_solid + cp_mushy + cp_fluid)
    return S1, S2, k.astype(FDT), rhocp.astype(FDT)


def computeConvRadBC(Level, LevelT0, ne, nn, properties, F):
    """cF:2207-2301: convection + radiation + evaporation flux on the top face of the
    elements [ne - ne_x*ne_y, ne), 2x2 Gauss, assembled on their 4 top nodes."""
    FDT = config.FDT
    p = f32props(properties)
    ne, nn = int(ne), int(nn)
    T_amb = p["T_amb"]
    T_boiling = p["T_boiling"]
    invT_b = FDT(1.0) / T_boiling
    x, y = Level["node_coords"][0], Level["node_coords"][1]
    cx, cy = Level["connect"][0], Level["connect"][1]
    ne_x, ne_y = cx.shape[0], cy.shape[0]
    top_ne = ne - ne_x * ne_y
    nn_x, nn_y = ne_x + 1, ne_y + 1
    coords = np.stack([x[cx[0, :]], y[cy[0, :]]], axis=1)
    N, _, wq = computeQuad2dFemShapeFunctions(coords)
    _, _, _, idx = convert2XYZ(np.arange(top_ne, ne), ne_x, ne_y, nn_x, nn_y)
    idx4 = idx[4:].T  # (n_top, 4)
    LevelT0 = np.asarray(LevelT0, dtype=FDT)
    Tq = LevelT0[idx4] @ N.T  # (n_top, 4q)
    Tq = np.minimum(Tq, T_boiling + FDT(1000))
    invT = FDT(1.0) / Tq
    E_pv = p["Lev"] + p["cp_fluid"] * (Tq - T_amb)
    MolMot = np.sqrt(p["CM_coeff"] * invT)
    S = p["evc"] * p["CP_coeff"] * np.exp(-p["CT_coeff"] * (invT - invT_b)) * MolMot * E_pv
    q_flux = p["h_conv"] * (T_amb - Tq) + p["sigma_sb"] * p["vareps"] * (T_amb**4 - Tq**4) - S
    q_flux = q_flux * FDT(1e-6)
    aT = (q_flux * wq.reshape(-1)[None, :]) @ N  # N.T @ (q*wq) per element
    NeumannBC = bincount(idx4.reshape(-1), aT.reshape(-1), nn)
    return (F + NeumannBC).astype(FDT)


def computeSourceFunction(x, y, z, v, properties, P):
    """cF:991-1025: 6*sqrt(3)*P*eta * prod_d exp(-3 (x_d - v_d)^2 / s_d^2) / (s_d sqrt(pi))."""
    FDT = config.FDT
    p = f32props(properties)
    v = np.asarray(v, dtype=FDT)
    P = FDT(P) if np.ndim(P) == 0 else np.asarray(P, dtype=FDT)
    _pcoeff = 6 * np.sqrt(FDT(3)) * P * p["laser_eta"]
    _rcoeff = 1 / (p["laser_radius"] * np.sqrt(FDT(np.pi)))
    _dcoeff = 1 / (p["laser_depth"] * np.sqrt(FDT(np.pi)))
    _rsq = p["laser_radius"] ** 2
    _dsq = p["laser_depth"] ** 2
    Qx = _rcoeff * np.exp(-3 * (x - v[0]) ** 2 / _rsq)
    Qy = _rcoeff * np.exp(-3 * (y - v[1]) ** 2 / _rsq)
    Qz = _dcoeff * np.exp(-3 * (z - v[2]) ** 2 / _dsq)
    return (_pcoeff * Qx * Qy * Qz).astype(FDT)


def elementSourceAtGauss(Level, v, ne, properties, laserP):
    """Shared body of cF:960-974 / 2986-3003 / 2695-2712: Q at the 8 Gauss points of the
    first ``ne`` elements -> (Q (ne,8q), Nf, wq scalar, idx (ne,8))."""
    coords = getSampleCoords(Level)
    Nf, _, wqf = computeQuad3dFemShapeFunctions(coords)
    ne_x, ne_y, nn_x, nn_y = _level_dims(Level)
    ix, iy, iz, idx = convert2XYZ(np.arange(int(ne)), ne_x, ne_y, nn_x, nn_y)
    x, y, z = getQuadratureCoords(Level, ix, iy, iz, Nf)
    Q = computeSourceFunction(x, y, z, v, properties, laserP)
    return Q, Nf, wqf[0, 0], idx.T


def computeSourcesL3(Level, v, ne_nn, properties, laserP):
    """cF:2960-3012: Level-3 load vector, assemble(Nf @ Q * wq)."""
    Q, Nf, wq, idx = elementSourceAtGauss(Level, v, ne_nn[1], properties, laserP)
    _data3 = (Q @ Nf.T) * wq  # "Nf @ Q * w" == (Nf @ Q) * w, cF:3003
    return bincount(idx.reshape(-1), _data3.reshape(-1), ne_nn[4])
