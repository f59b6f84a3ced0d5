"""Oracle (test infrastructure): inter-level transfer operators of the reference, NumPy f32.

Restates compute3DN cF:1361-1393 - interpolatePointsMatrix cF:1028-1107 -
interpolate_w_matrix cF:1110-1128 - interpolatePoints cF:1131-1210 -
computeCoarseFineShapeFunctions cF:1213-1358 (materialised (nef,8,8) operators + scatter
indices in place of the BCOO matrix) - getOverlapRegion cF:1642-1669 - getNewTprime cF:2060-2099 -
getBothNewTprimes cF:2102-2132 - computeCoarseTprimeTerm_jax cF:1477-1565 -
computeCoarseTprimeMassTerm_jax cF:1396-1474 - computeL1TprimeTerms_Part1/2 cF:2733-2810,
3057-3132 - computeL2TprimeTerms_Part1/2 cF:2857-2914, 3169-3221 - computeSources cF:928-988 -
computeLevelSource cF:2667-2730.
"""
import numpy as np

from . import config
from .fem import (
    bincount,
    computeQuad3dFemShapeFunctions,
    convert2XYZ,
    elementSourceAtGauss,
    getSampleCoords,
    _level_dims,
)


def compute3DN(q, x, y, z, h):
    """cF:1361-1393: the 8 trilinear weights of point q in the box [x0,x1]x[y0,y1]x[z0,z1].
    Vectorised: every argument may be an array; result has a trailing axis of 8."""
    FDT = config.FDT
    inv_vol = FDT(1.0) / (h[0] * h[1] * h[2])
    N = np.stack(
        [
            (x[1] - q[0]) * (y[1] - q[1]) * (z[1] - q[2]),
            (q[0] - x[0]) * (y[1] - q[1]) * (z[1] - q[2]),
            (q[0] - x[0]) * (q[1] - y[0]) * (z[1] - q[2]),
            (x[1] - q[0]) * (q[1] - y[0]) * (z[1] - q[2]),
            (x[1] - q[0]) * (y[1] - q[1]) * (q[2] - z[0]),
            (q[0] - x[0]) * (y[1] - q[1]) * (q[2] - z[0]),
            (q[0] - x[0]) * (q[1] - y[0]) * (q[2] - z[0]),
            (x[1] - q[0]) * (q[1] - y[0]) * (q[2] - z[0]),
        ],
        axis=-1,
    )
    return (N * inv_vol).astype(FDT)


def _locate(Level, node_coords_new):
    """Common body of cF:1046-1098 / 1150-1202 for the tensor-product target grid."""
    FDT = config.FDT
    nc = Level["node_coords"]
    cn = Level["connect"]
    ne_x, ne_y, ne_z = cn[0].shape[0], cn[1].shape[0], cn[2].shape[0]
    nn_x, nn_y = ne_x + 1, ne_y + 1
    xn, yn, zn = [np.asarray(c, dtype=FDT) for c in node_coords_new]
    nn_xn, nn_yn, nn_zn = len(xn), len(yn), len(zn)
    total = nn_xn * nn_yn * nn_zn
    h_x = nc[0][1] - nc[0][0]
    h_y = nc[1][1] - nc[1][0]
    h_z = nc[2][1] - nc[2][0]
    ielt = np.arange(total)
    izn, rem = np.divmod(ielt, nn_xn * nn_yn)
    iyn, ixn = np.divmod(rem, nn_xn)
    _x, _y, _z = xn[ixn], yn[iyn], zn[izn]
    ielt_x = np.clip(np.floor((_x - nc[0][0]) / h_x).astype(int), 0, ne_x - 1)
    ielt_y = np.clip(np.floor((_y - nc[1][0]) / h_y).astype(int), 0, ne_y - 1)
    ielt_z = np.clip(np.floor((_z - nc[2][0]) / h_z).astype(int), 0, ne_z - 1)
    nodex = cn[0][ielt_x, :]
    nodey = cn[1][ielt_y, :]
    nodez = cn[2][ielt_z, :]
    node = nodex + nodey * nn_x + nodez * (nn_x * nn_y)
    xx = nc[0][nodex]
    yy = nc[1][nodey]
    zz = nc[2][nodez]
    Nc = compute3DN(
        [_x, _y, _z],
        [xx[:, 0], xx[:, 1]],
        [yy[:, 0], yy[:, 3]],
        [zz[:, 0], zz[:, 5]],
        [h_x, h_y, h_z],
    )
    valid = np.logical_and((Nc >= -1e-2).all(axis=1), (Nc <= 1 + 1e-2).all(axis=1))
    Nc = np.where(valid[:, None], np.clip(Nc, 0.0, 1.0), np.zeros_like(Nc)).astype(FDT)
    return Nc, node


def interpolatePointsMatrix(Level, node_coords_new):
    """cF:1028-1107: [Nc (n,8) f32, node (n,8) int]; all-zero weights outside the +-1e-2 window."""
    Nc, node = _locate(Level, node_coords_new)
    return [Nc, node]


def interpolate_w_matrix(C2F, T):
    """cF:1110-1128."""
    T = np.asarray(T, dtype=config.FDT)
    return (C2F[0] * T[C2F[1]]).sum(axis=1, dtype=config.FDT)


def interpolatePoints(Level, u, node_coords_new):
    """cF:1131-1210: Nc @ u[node] per target point."""
    Nc, node = _locate(Level, node_coords_new)
    u = np.asarray(u).astype(config.FDT)
    return (Nc * u[node]).sum(axis=1, dtype=config.FDT)


def computeCoarseFineShapeFunctions(Coarse, Fine):
    """cF:1213-1358: coarse N and grad N at every fine Gauss point, (nef, 8q, 8c) each, and
    the scatter targets (coarse node ids of Gauss point 0, cF:1351).  Returns
    [Nc, [dNcdx, dNcdy, dNcdz], nodes (nef, 8c)]; ``project(nodes, data, nnc)`` plays the
    role of the BCOO product ``Shapes[.][2] @ data.reshape(-1)``."""
    FDT = config.FDT
    cc, cn = Coarse["node_coords"], Coarse["connect"]
    fc, fn = Fine["node_coords"], Fine["connect"]
    nec_x, nec_y, nec_z = [cn[i].shape[0] for i in range(3)]
    nnc_x, nnc_y, nnc_z = [cc[i].shape[0] for i in range(3)]
    nef_x, nef_y, nef_z = [fn[i].shape[0] for i in range(3)]
    nef = nef_x * nef_y * nef_z
    nnf_x, nnf_y = [fc[i].shape[0] for i in range(2)]
    hc_x = cc[0][1] - cc[0][0]
    hc_y = cc[1][1] - cc[1][0]
    hc_z = cc[2][1] - cc[2][0]
    hc_xyz = hc_x * hc_y * hc_z
    xminc_x, xminc_y, xminc_z = [cc[i][0] for i in range(3)]
    coords = np.stack([fc[0][fn[0][0, :]], fc[1][fn[1][0, :]], fc[2][fn[2][0, :]]], axis=1)
    Nf, _, _ = computeQuad3dFemShapeFunctions(coords)

    ix, iy, iz, _ = convert2XYZ(np.arange(nef), nef_x, nef_y, nnf_x, nnf_y)
    x = fc[0][fn[0][ix, :]] @ Nf.T  # (nef, 8q)
    y = fc[1][fn[1][iy, :]] @ Nf.T
    z = fc[2][fn[2][iz, :]] @ Nf.T
    ieltc_x = np.clip(np.floor((x - xminc_x) / hc_x).astype(int), 0, nec_x - 1)
    ieltc_y = np.clip(np.floor((y - xminc_y) / hc_y).astype(int), 0, nec_y - 1)
    ieltc_z = np.clip(np.floor((z - xminc_z) / hc_z).astype(int), 0, nec_z - 1)
    nodec_x = cn[0][ieltc_x, :]  # (nef, 8q, 8c)
    nodec_y = cn[1][ieltc_y, :]
    nodec_z = cn[2][ieltc_z, :]
    nodes = nodec_x + nodec_y * nnc_x + nodec_z * nnc_x * nnc_y
    xc0 = cc[0][nodec_x[..., 0]]
    xc1 = cc[0][nodec_x[..., 1]]
    yc0 = cc[1][nodec_y[..., 0]]
    yc3 = cc[1][nodec_y[..., 3]]
    zc0 = cc[2][nodec_z[..., 0]]
    zc5 = cc[2][nodec_z[..., 5]]
    Nc = compute3DN([x, y, z], [xc0, xc1], [yc0, yc3], [zc0, zc5], [hc_x, hc_y, hc_z])
    _x, _y, _z = x, y, z
    dNcdx = (
        np.stack(
            [
                (-1 * (yc3 - _y) * (zc5 - _z)),
                (1 * (yc3 - _y) * (zc5 - _z)),
                (1 * (_y - yc0) * (zc5 - _z)),
                (-1 * (_y - yc0) * (zc5 - _z)),
                (-1 * (yc3 - _y) * (_z - zc0)),
                (1 * (yc3 - _y) * (_z - zc0)),
                (1 * (_y - yc0) * (_z - zc0)),
                (-1 * (_y - yc0) * (_z - zc0)),
            ],
            axis=-1,
        )
        / hc_xyz
    ).astype(FDT)
    dNcdy = (
        np.stack(
            [
                ((xc1 - _x) * -1 * (zc5 - _z)),
                ((_x - xc0) * -1 * (zc5 - _z)),
                ((_x - xc0) * 1 * (zc5 - _z)),
                ((xc1 - _x) * 1 * (zc5 - _z)),
                ((xc1 - _x) * -1 * (_z - zc0)),
                ((_x - xc0) * -1 * (_z - zc0)),
                ((_x - xc0) * 1 * (_z - zc0)),
                ((xc1 - _x) * 1 * (_z - zc0)),
            ],
            axis=-1,
        )
        / hc_xyz
    ).astype(FDT)
    dNcdz = (
        np.stack(
            [
                ((xc1 - _x) * (yc3 - _y) * -1),
                ((_x - xc0) * (yc3 - _y) * -1),
                ((_x - xc0) * (_y - yc0) * -1),
                ((xc1 - _x) * (_y - yc0) * -1),
                ((xc1 - _x) * (yc3 - _y) * 1),
                ((_x - xc0) * (yc3 - _y) * 1),
                ((_x - xc0) * (_y - yc0) * 1),
                ((xc1 - _x) * (_y - yc0) * 1),
            ],
            axis=-1,
        )
        / hc_xyz
    ).astype(FDT)
    _nodes = nodes[:, 0, :]
    nnc = nnc_x * nnc_y * nnc_z
    return [Nc, [dNcdx, dNcdy, dNcdz], (_nodes, nnc)]


def project(test, data):
    """``Shapes[.][2] @ data.reshape(-1)``: BCOO matvec with one 1 per column (cF:1354-1357)
    == float32 scatter-add of data[(e, c)] into coarse node _nodes[e, c]."""
    _nodes, nnc = test
    return bincount(_nodes.reshape(-1), np.asarray(data).reshape(-1), nnc)


def getOverlapRegion(node_coords, nx, ny):
    """cF:1642-1669: flat ids of the tensor-product index set."""
    nx, ny = int(nx), int(ny)
    a, b, c = [np.asarray(v) for v in node_coords]
    _x = np.tile(a, b.shape[0] * c.shape[0]).reshape(-1)
    _y = np.repeat(np.tile(b, c.shape[0]), a.shape[0]).reshape(-1)
    _z = np.repeat(c, a.shape[0] * b.shape[0])
    return _x + _y * nx + _z * nx * ny


def getNewTprime(Fine, FineT0, CoarseT, Coarse, C2F):
    """cF:2060-2099: inject fine T into the parent's overlap nodes; T' = T_f - I(parent)."""
    _val = interpolatePoints(Fine, FineT0, Fine["overlapCoords"])
    _idx = getOverlapRegion(Fine["overlapNodes"], Coarse["nodes"][0], Coarse["nodes"][1])
    CoarseT = np.array(CoarseT, dtype=config.FDT, copy=True)
    CoarseT[_idx] = _val
    Tprime = np.asarray(FineT0, dtype=config.FDT) - interpolate_w_matrix(C2F, CoarseT)
    return Tprime, CoarseT


def getBothNewTprimes(Levels, FineT, MesoT, M2F, CoarseT, C2M):
    """cF:2102-2132."""
    lTprime, mT0 = getNewTprime(Levels[3], FineT, MesoT, Levels[2], M2F)
    mTprime, uT0 = getNewTprime(Levels[2], mT0, CoarseT, Levels[1], C2M)
    return lTprime, mTprime, mT0, uT0


def _elem_setup(Level):
    ne_x, ne_y, nn_x, nn_y = _level_dims(Level)
    ne = ne_x * ne_y * int(Level["elements"][2])
    N, dNdx, wq = computeQuad3dFemShapeFunctions(getSampleCoords(Level))
    _, _, _, idx = convert2XYZ(np.arange(ne), ne_x, ne_y, nn_x, nn_y)
    return N, dNdx, wq, idx  # idx (8, ne)


def _grad_term(Level, Tprime0, kfield, dShape):
    """-(sum_d dNc_d * kbar dT'/dx_d) * wq summed over Gauss points -> (nef, 8c).
    cF:1509-1525 (and 2765-2783, 2885-2911): wq[None,None,:] is (1,1,8,1) so the product is
    (1,nef,8q,8c) and .sum(axis=2) runs over q."""
    FDT = config.FDT
    N, dNdx, wq, idx = _elem_setup(Level)
    Tp = np.asarray(Tprime0, dtype=FDT)[idx]  # (8a, nef)
    kMean = (N @ np.asarray(kfield, dtype=FDT)[idx]).mean(axis=0, dtype=FDT)  # (nef,)
    out = None
    for i in range(3):
        d = kMean * (dNdx[:, :, i] @ Tp)  # (8q, nef)
        term = ((-dShape[i] * d.T[:, :, None]) * wq[None, None, :]).sum(axis=2, dtype=FDT)
        out = term if out is None else out + term
    return out[0]  # drop the broadcast leading 1


def _mass_term(Level, dTprime, rhocp, Shape0, dt):
    """-(Nc * (N dT') * mean_q(N rhocp)) * (1/dt) wq summed over q -> (nef, 8c).
    cF:1439-1446 (and 3105-3109, 3123-3127, 3210-3216)."""
    FDT = config.FDT
    N, _, wq, idx = _elem_setup(Level)
    _Tp = (N @ np.asarray(dTprime, dtype=FDT)[idx]) * (
        N @ np.asarray(rhocp, dtype=FDT)[idx]
    ).mean(axis=0, dtype=FDT)  # (8q, nef)
    w = (FDT(1) / FDT(dt)) * wq[None, None, :]
    out = ((-Shape0 * _Tp.T[:, :, None]) * w).sum(axis=2, dtype=FDT)
    return out[0]


def computeCoarseTprimeTerm(Levels, L3k, L2k, Shapes):
    """cF:1477-1565: Vcu (on L1, from L3 and L2 T'0) and Vmu (on L2, from L3 T'0)."""
    _data1 = _grad_term(Levels[3], Levels[3]["Tprime0"], L3k, Shapes[1][1])
    _data2 = _grad_term(Levels[3], Levels[3]["Tprime0"], L3k, Shapes[2][1])
    _data3 = _grad_term(Levels[2], Levels[2]["Tprime0"], L2k, Shapes[0][1])
    Vcu = project(Shapes[1][2], _data1) + project(Shapes[0][2], _data3)
    Vmu = project(Shapes[2][2], _data2)
    return Vcu, Vmu


def computeCoarseTprimeMassTerm(Levels, Tprimef, Tprimem, L3rhocp, L2rhocp, dt, Shapes, Vcu, Vmu):
    """cF:1396-1474."""
    Tprimef_new = Tprimef - Levels[3]["Tprime0"]
    Tprimem_new = Tprimem - Levels[2]["Tprime0"]
    _data1 = _mass_term(Levels[3], Tprimef_new, L3rhocp, Shapes[1][0], dt)
    _data2 = _mass_term(Levels[3], Tprimef_new, L3rhocp, Shapes[2][0], dt)
    _data3 = _mass_term(Levels[2], Tprimem_new, L2rhocp, Shapes[0][0], dt)
    Vcu = Vcu + (project(Shapes[1][2], _data1) + project(Shapes[0][2], _data3))
    Vmu = Vmu + project(Shapes[2][2], _data2)
    return Vcu, Vmu


def computeL1TprimeTerms_Part1(Levels, ne_nn, L3k, Shapes, L2k):
    """cF:2733-2810."""
    _1 = _grad_term(Levels[3], Levels[3]["Tprime0"], L3k, Shapes[1][1])
    _2 = _grad_term(Levels[2], Levels[2]["Tprime0"], L2k, Shapes[0][1])
    return project(Shapes[1][2], _1) + project(Shapes[0][2], _2)


def computeL2TprimeTerms_Part1(Levels, ne_nn, L3Tprime0, L3k, Shapes):
    """cF:2857-2914."""
    _1 = _grad_term(Levels[3], L3Tprime0, L3k, Shapes[2][1])
    return project(Shapes[2][2], _1)


def computeL1TprimeTerms_Part2(Levels, ne_nn, L3Tp, L2Tp, L3rhocp, L2rhocp, dt, Shapes, Vcu):
    """cF:3057-3132."""
    L3Tp_new = L3Tp - Levels[3]["Tprime0"]
    L2Tp_new = L2Tp - Levels[2]["Tprime0"]
    _1 = _mass_term(Levels[3], L3Tp_new, L3rhocp, Shapes[1][0], dt)
    _2 = _mass_term(Levels[2], L2Tp_new, L2rhocp, Shapes[0][0], dt)
    return Vcu + (project(Shapes[1][2], _1) + project(Shapes[0][2], _2))


def computeL2TprimeTerms_Part2(Levels, ne_nn, L3Tp, L3Tp0, L3rhocp, dt, Shapes, L2V):
    """cF:3169-3221."""
    L3Tp_new = L3Tp - L3Tp0
    _data2 = _mass_term(Levels[3], L3Tp_new, L3rhocp, Shapes[2][0], dt)
    return L2V + project(Shapes[2][2], _data2)


def computeSources(Level, v, Shapes, ne_nn, properties, laserP):
    """cF:928-988: Ff assembled on L3; Fc, Fm = project(sum_q Nc * Q wq)."""
    FDT = config.FDT
    Q, Nf, wq, idx = elementSourceAtGauss(Level, v, ne_nn[1], properties, laserP)
    _data = Q * wq  # (nef, 8q)
    _data3 = (Q @ Nf.T) * wq
    _data1 = (Shapes[1][0] * _data[:, :, None]).sum(axis=1, dtype=FDT)
    _data2 = (Shapes[2][0] * _data[:, :, None]).sum(axis=1, dtype=FDT)
    Fc = project(Shapes[1][2], _data1)
    Fm = project(Shapes[2][2], _data2)
    Ff = bincount(idx.reshape(-1), _data3.reshape(-1), ne_nn[4])
    return Fc, Fm, Ff


def computeLevelSource(Levels, ne_nn, laser_position, LevelShape, properties, laserP):
    """cF:2667-2730: projected source averaged over the rows of laser_position."""
    FDT = config.FDT
    laser_position = np.asarray(laser_position, dtype=FDT)
    laserP = np.asarray(laserP, dtype=FDT)
    lshape = laser_position.shape[0]
    acc = None
    for il in range(lshape):
        Q, _, wq, _ = elementSourceAtGauss(
            Levels[3], laser_position[il], ne_nn[1], properties, laserP[il]
        )
        _data = Q * wq
        t = (LevelShape[0] * _data[:, :, None]).sum(axis=1, dtype=FDT)
        acc = t if acc is None else acc + t
    _data1 = acc / lshape
    return project(LevelShape[2], _data1)
