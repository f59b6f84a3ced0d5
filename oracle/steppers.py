"""Oracle (test infrastructure): step orchestrators and mesh movement of the reference.

Restates assignBCs cF:1568-1595 - assignBCsFine cF:1598-1620 - substitute_Tbar cF:1848-1865 -
jit_constrain_v cF:1672-1695 - move_fine_mesh cF:1698-1724 - update_overlap_nodes_coords{,_L1L2,_L2}
cF:1727-1845 - computeSolutions cF:2135-2204 - computeL1Temperature cF:2813-2854 -
computeL2Temperature cF:2917-2957 - computeSolutions_L3 cF:3015-3054 - updateStateProperties
cF:2513-2564 - stepGOMELT cF:2304-2397 - stepGOMELTDwellTime cF:2617-2664 - subcycleGOMELT
cF:3224-3632 - moveEverything cF:2400-2510 - melting_temp cF:3696-3712 - printLevelMaxMin
cF:3635-3665.
"""
import numpy as np

from . import config
from .fem import (
    computeConvRadBC,
    computeSourcesL3,
    computeStateProperties,
    solveMatrixFreeFE,
)
from .transfer import (
    computeCoarseFineShapeFunctions,
    computeCoarseTprimeMassTerm,
    computeCoarseTprimeTerm,
    computeL1TprimeTerms_Part1,
    computeL1TprimeTerms_Part2,
    computeL2TprimeTerms_Part1,
    computeL2TprimeTerms_Part2,
    computeLevelSource,
    computeSources,
    getBothNewTprimes,
    getNewTprime,
    getOverlapRegion,
    interpolate_w_matrix,
    interpolatePoints,
    interpolatePointsMatrix,
)


def assignBCs(RHS, Levels):
    """cF:1568-1595: y-, y+, x-, x+, z- faces of Level 1 <- conditions (in that order)."""
    _RHS = np.array(RHS, dtype=config.FDT, copy=True)
    c = Levels[1]["conditions"]
    BC = Levels[1]["BC"]
    _RHS[BC[2]] = c["y"][0]
    _RHS[BC[3]] = c["y"][1]
    _RHS[BC[0]] = c["x"][0]
    _RHS[BC[1]] = c["x"][1]
    _RHS[BC[4]] = c["z"][0]
    return _RHS


def assignBCsFine(RHS, TfAll, BC):
    """cF:1598-1620."""
    _RHS = np.array(RHS, dtype=config.FDT, copy=True)
    for b in (2, 3, 0, 1, 4):
        _RHS[BC[b]] = TfAll[BC[b]]
    return _RHS


def substitute_Tbar(Tbar, _idx, _val):
    """cF:1848-1865."""
    out = np.array(Tbar, dtype=config.FDT, copy=True)
    out[int(_idx):] = _val
    return out


def jit_constrain_v(vtot, Level):
    """cF:1672-1695."""
    FDT = config.FDT
    b = Level["bounds"]
    return [
        np.clip(FDT(vtot[0]), FDT(b["ix"][0]), FDT(b["ix"][1])),
        np.clip(FDT(vtot[1]), FDT(b["iy"][0]), FDT(b["iy"][1])),
        np.clip(FDT(vtot[2]), FDT(b["iz"][0]), FDT(b["iz"][1])),
    ]


def _shift(v, h):
    """(v / h + 1e-2).astype(int): truncation toward zero of an f32 quotient (cF:1716)."""
    FDT = config.FDT
    return int(np.asarray(FDT(v) / FDT(h) + FDT(1e-2)).astype(int))


def move_fine_mesh(node_coords, element_size, v):
    """cF:1698-1724."""
    FDT = config.FDT
    s = [_shift(v[i], element_size[i]) for i in range(3)]
    new = [(node_coords[i] + FDT(element_size[i]) * s[i]).astype(FDT) for i in range(3)]
    return new, s


def update_overlap_nodes_coords(Level, vcon, element_size, ele_ratio):
    """cF:1727-1764."""
    FDT = config.FDT
    shift = [_shift(vcon[i], element_size[i]) for i in range(3)]
    Level["overlapNodes"] = [
        Level["orig_overlap_nodes"][i] + int(ele_ratio[i]) * shift[i] for i in range(3)
    ]
    Level["overlapCoords"] = [
        (Level["orig_overlap_coors"][i] + FDT(element_size[i]) * shift[i]).astype(FDT)
        for i in range(3)
    ]
    return Level


def update_overlap_nodes_coords_L1L2(Level, vcon, element_size, powder_layer):
    """cF:1767-1804: z index shift in Level-1 cells, z coordinate shift in layer heights."""
    FDT = config.FDT
    sx = _shift(vcon[0], element_size[0])
    sy = _shift(vcon[1], element_size[1])
    sz = _shift(vcon[2], element_size[2])
    shift_z_p = FDT(powder_layer) * _shift(vcon[2], powder_layer)
    o = Level["orig_overlap_nodes"]
    c = Level["orig_overlap_coors"]
    Level["overlapNodes"] = [o[0] + sx, o[1] + sy, o[2] + sz]
    Level["overlapCoords"] = [
        (c[0] + FDT(element_size[0]) * sx).astype(FDT),
        (c[1] + FDT(element_size[1]) * sy).astype(FDT),
        (c[2] + shift_z_p).astype(FDT),
    ]
    return Level


def update_overlap_nodes_coords_L2(Level, vcon, element_size, ele_ratio):
    """cF:1807-1845."""
    FDT = config.FDT
    shift = [_shift(vcon[i], element_size[i]) for i in range(3)]
    Level["overlapNodes_L2"] = [
        Level["orig_overlap_nodes_L2"][i] + int(ele_ratio[i]) * shift[i] for i in range(3)
    ]
    Level["overlapCoords_L2"] = [
        (Level["orig_overlap_coors_L2"][i] + FDT(element_size[i]) * shift[i]).astype(FDT)
        for i in range(3)
    ]
    return Level


def computeSolutions(Levels, ne_nn, tmp_ne_nn, LF, L1V, LInterp, Lk, Lrhocp, L2V, dt, properties):
    """cF:2135-2204."""
    FDT = config.FDT
    L1T = solveMatrixFreeFE(
        Levels[1], ne_nn[2], tmp_ne_nn[0], Lk[1], Lrhocp[1], dt, Levels[1]["T0"], LF[1], L1V
    )
    L1T = substitute_Tbar(L1T, tmp_ne_nn[1], FDT(properties["T_amb"]))
    FinalL1 = assignBCs(L1T, Levels)
    TfAll = interpolate_w_matrix(LInterp[0], FinalL1)
    L2T = solveMatrixFreeFE(
        Levels[2], ne_nn[3], ne_nn[0], Lk[2], Lrhocp[2], dt, Levels[2]["T0"], LF[2], L2V
    )
    FinalL2 = assignBCsFine(L2T, TfAll, Levels[2]["BC"])
    TfAll = interpolate_w_matrix(LInterp[1], FinalL2)
    FinalL3 = solveMatrixFreeFE(
        Levels[3], ne_nn[4], ne_nn[1], Lk[3], Lrhocp[3], dt, Levels[3]["T0"], LF[3], 0
    )
    FinalL3 = assignBCsFine(FinalL3, TfAll, Levels[3]["BC"])
    return FinalL1, FinalL2, FinalL3


def computeL1Temperature(Levels, ne_nn, tmp_ne_nn, L1F, L1V, L1k, L1rhocp, dt, properties):
    """cF:2813-2854."""
    L1T = solveMatrixFreeFE(
        Levels[1], ne_nn[2], tmp_ne_nn[0], L1k, L1rhocp, dt, Levels[1]["T0"], L1F, L1V
    )
    L1T = substitute_Tbar(L1T, tmp_ne_nn[1], config.FDT(properties["T_amb"]))
    return assignBCs(L1T, Levels)


def computeL2Temperature(L1T, L1L2Interp, Levels, ne_nn, L2T0, L2F, L2V, L2k, L2rhocp, dt):
    """cF:2917-2957."""
    TfAll = interpolate_w_matrix(L1L2Interp, L1T)
    L2T = solveMatrixFreeFE(Levels[2], ne_nn[3], ne_nn[0], L2k, L2rhocp, dt, L2T0, L2F, L2V)
    return assignBCsFine(L2T, TfAll, Levels[2]["BC"])


def computeSolutions_L3(FinalL2, L2L3Interp, Levels, ne_nn, L3T0, L3F, L3k, L3rhocp, dt):
    """cF:3015-3054."""
    TfAll = interpolate_w_matrix(L2L3Interp, FinalL2)
    FinalL3 = solveMatrixFreeFE(Levels[3], ne_nn[4], ne_nn[1], L3k, L3rhocp, dt, L3T0, L3F, 0)
    return assignBCsFine(FinalL3, TfAll, Levels[3]["BC"])


def _push_S1_to_L1(Levels, substrate):
    """cF:2546-2556 / 3272-3278: L2.S1 -> Level-1 overlap nodes, substrate -> 1."""
    interpolated_S1 = interpolatePoints(Levels[2], Levels[2]["S1"], Levels[2]["overlapCoords"])
    overlap_idx_L1 = getOverlapRegion(
        Levels[2]["overlapNodes"], Levels[1]["nodes"][0], Levels[1]["nodes"][1]
    )
    S1 = np.array(Levels[1]["S1"], dtype=config.FDT, copy=True)
    S1[overlap_idx_L1] = interpolated_S1
    S1[: int(substrate[1])] = 1
    Levels[1]["S1"] = S1


def updateStateProperties(Levels, properties, substrate):
    """cF:2513-2564."""
    Levels[3]["S1"], Levels[3]["S2"], L3k, L3rhocp = computeStateProperties(
        Levels[3]["T0"], Levels[3]["S1"], properties, substrate[3]
    )
    Levels[2]["S1"], _L2S2, L2k, L2rhocp = computeStateProperties(
        Levels[2]["T0"], Levels[2]["S1"], properties, substrate[2]
    )
    _push_S1_to_L1(Levels, substrate)
    _, _, L1k, L1rhocp = computeStateProperties(
        Levels[1]["T0"], Levels[1]["S1"], properties, substrate[1]
    )
    return Levels, [0, L1k, L2k, L3k], [0, L1rhocp, L2rhocp, L3rhocp]


def _scatter_L0(Levels):
    """cF:2390-2392 / 3628-3630."""
    S1 = np.array(Levels[0]["S1"], copy=True)
    S1[Levels[0]["idx"]] = Levels[3]["S1"]
    Levels[0]["S1"] = S1
    S2 = np.zeros_like(Levels[0]["S2"])
    S2[Levels[0]["idx"]] = Levels[3]["S2"]
    Levels[0]["S2"] = S2


def stepGOMELT(Levels, ne_nn, tmp_ne_nn, Shapes, LInterp, v, properties, dt, laserP, substrate):
    """cF:2304-2397: one single-step predictor/corrector update of all three levels."""
    FDT = config.FDT
    T_amb = FDT(properties["T_amb"])
    preS2 = Levels[3]["S2"]
    Levels, Lk, Lrhocp = updateStateProperties(Levels, properties, substrate)
    L1, L2, L3 = Levels[1], Levels[2], Levels[3]
    Fc, Fm, Ff = computeSources(L3, v, Shapes, ne_nn, properties, laserP)
    Fc = computeConvRadBC(L1, L1["T0"], tmp_ne_nn[0], ne_nn[2], properties, Fc)
    Fm = computeConvRadBC(L2, L2["T0"], ne_nn[0], ne_nn[3], properties, Fm)
    Ff = computeConvRadBC(L3, L3["T0"], ne_nn[1], ne_nn[4], properties, Ff)
    F = [0, Fc, Fm, Ff]
    Vcu, Vmu = computeCoarseTprimeTerm(Levels, Lk[3], Lk[2], Shapes)
    L1T, L2T, L3T = computeSolutions(
        Levels, ne_nn, tmp_ne_nn, F, Vcu, LInterp, Lk, Lrhocp, Vmu, dt, properties
    )
    L1T = np.maximum(T_amb, L1T)
    L2T = np.maximum(T_amb, L2T)
    L3T = np.maximum(T_amb, L3T)
    L3Tp, L2Tp, L2T, L1T = getBothNewTprimes(Levels, L3T, L2T, LInterp[1], L1T, LInterp[0])
    Vcu, Vmu = computeCoarseTprimeMassTerm(
        Levels, L3Tp, L2Tp, Lrhocp[3], Lrhocp[2], dt, Shapes, Vcu, Vmu
    )
    L1T, L2T, L3T0 = computeSolutions(
        Levels, ne_nn, tmp_ne_nn, F, Vcu, LInterp, Lk, Lrhocp, Vmu, dt, properties
    )
    L1T = np.maximum(T_amb, L1T)
    L2T = np.maximum(T_amb, L2T)
    Levels[3]["T0"] = np.maximum(T_amb, L3T0)
    (
        Levels[3]["Tprime0"],
        Levels[2]["Tprime0"],
        Levels[2]["T0"],
        Levels[1]["T0"],
    ) = getBothNewTprimes(Levels, Levels[3]["T0"], L2T, LInterp[1], L1T, LInterp[0])
    _scatter_L0(Levels)
    _resetmask = ((1 - 2 * preS2.astype(np.int32)) * Levels[3]["S2"].astype(np.int32)) == 1
    return Levels, _resetmask


def stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, properties, dt, substrate):
    """cF:2617-2664: Level 1 only; no max(T_amb, .) clamp."""
    L1 = Levels[1]
    Fc = computeConvRadBC(L1, L1["T0"], tmp_ne_nn[0], ne_nn[2], properties, 0)
    _, _, k, rhocp = computeStateProperties(L1["T0"], L1["S1"], properties, substrate[1])
    T_new = solveMatrixFreeFE(L1, ne_nn[2], tmp_ne_nn[0], k, rhocp, dt, L1["T0"], Fc, 0)
    T_new = substitute_Tbar(T_new, tmp_ne_nn[1], config.FDT(properties["T_amb"]))
    Levels[1]["T0"] = assignBCs(T_new, Levels)
    return Levels


def subcycleGOMELT(
    Levels,
    ne_nn,
    Shapes,
    substrate,
    LInterp,
    tmp_ne_nn,
    laser_position,
    properties,
    laserP,
    subcycle,
    max_accum_L3,
    accum_L3,
    return_history=False,
):
    """cF:3224-3632: L1 once, L2 x N2, L3 x N2*N3, run as predictor pass then corrector pass."""
    FDT = config.FDT
    T_amb = FDT(properties["T_amb"])
    laser_position = np.asarray(laser_position, dtype=FDT)
    laserP = np.asarray(laserP, dtype=FDT)
    N2, N3 = int(subcycle[0]), int(subcycle[1])
    fN2, fN3 = FDT(subcycle[3]), FDT(subcycle[4])

    _, _, L3k_L1, L3rhocp_L1 = computeStateProperties(
        Levels[3]["T0"], Levels[3]["S1"], properties, substrate[3]
    )
    _, _, L2k_L1, L2rhocp_L1 = computeStateProperties(
        Levels[2]["T0"], Levels[2]["S1"], properties, substrate[2]
    )
    _push_S1_to_L1(Levels, substrate)
    _, _, L1k, L1rhocp = computeStateProperties(
        Levels[1]["T0"], Levels[1]["S1"], properties, substrate[1]
    )
    dt_all = laser_position[:, 5].sum(dtype=FDT)
    L1F = computeLevelSource(Levels, ne_nn, laser_position, Shapes[1], properties, laserP)
    L1F = computeConvRadBC(Levels[1], Levels[1]["T0"], tmp_ne_nn[0], ne_nn[2], properties, L1F)
    L1V = computeL1TprimeTerms_Part1(Levels, ne_nn, L3k_L1, Shapes, L2k_L1)
    L1T = computeL1Temperature(Levels, ne_nn, tmp_ne_nn, L1F, L1V, L1k, L1rhocp, dt_all, properties)
    L1T = np.maximum(T_amb, L1T)

    def L2_common(_L2carry, _L2sub):
        alpha_L2 = FDT(_L2sub + 1) / fN2
        beta_L2 = FDT(1) - alpha_L2
        Lidx = _L2sub * N3 + np.arange(N3)
        _, _, L3k_L2, L3rhocp_L2 = computeStateProperties(
            _L2carry[2], _L2carry[4], properties, substrate[3]
        )
        L2S1, _, L2k, L2rhocp = computeStateProperties(
            _L2carry[0], _L2carry[1], properties, substrate[2]
        )
        L2F = computeLevelSource(
            Levels, ne_nn, laser_position[Lidx, :], Shapes[2], properties, laserP[Lidx]
        )
        L2F = computeConvRadBC(Levels[2], _L2carry[0], ne_nn[0], ne_nn[3], properties, L2F)
        L2V = computeL2TprimeTerms_Part1(Levels, ne_nn, _L2carry[3], L3k_L2, Shapes)
        dt2 = laser_position[Lidx, 5].sum(dtype=FDT)
        return alpha_L2, beta_L2, L3rhocp_L2, L2S1, L2k, L2rhocp, L2F, L2V, dt2

    def L3_substep(T3, S13, _L3sub, _L2sub, L2T, L2T_prev):
        L3S1, L3S2, L3k, L3rhocp = computeStateProperties(T3, S13, properties, substrate[3])
        LLidx = _L3sub + _L2sub * N3
        L3F = computeSourcesL3(Levels[3], laser_position[LLidx, :], ne_nn, properties, laserP[LLidx])
        L3F = computeConvRadBC(Levels[3], T3, ne_nn[1], ne_nn[4], properties, L3F)
        alpha_L3 = FDT(_L3sub + 1) / fN3
        beta_L3 = FDT(1) - alpha_L3
        _BC = alpha_L3 * L2T + beta_L3 * L2T_prev
        L3T = computeSolutions_L3(
            _BC, LInterp[1], Levels, ne_nn, T3, L3F, L3k, L3rhocp, laser_position[LLidx, 5]
        )
        L3T = np.maximum(T_amb, L3T)
        return L3T, L3S1, L3S2, LLidx

    # ---- predictor pass (cF:3308-3430) ----
    carry = [Levels[2]["T0"], Levels[2]["S1"], Levels[3]["T0"], Levels[3]["Tprime0"], Levels[3]["S1"]]
    L3Tp_L2 = []
    for _L2sub in range(N2):
        a2, b2, _, L2S1, L2k, L2rhocp, L2F, L2V, dt2 = L2_common(carry, _L2sub)
        _BC = a2 * L1T + b2 * Levels[1]["T0"]
        L2T = computeL2Temperature(
            _BC, LInterp[0], Levels, ne_nn, carry[0], L2F, L2V, L2k, L2rhocp, dt2
        )
        L2T = np.maximum(T_amb, L2T)
        T3, S13 = carry[2], carry[4]
        for _L3sub in range(N3):
            T3, S13, _, _ = L3_substep(T3, S13, _L3sub, _L2sub, L2T, carry[0])
        L3Tp, L2T = getNewTprime(Levels[3], T3, L2T, Levels[2], LInterp[1])
        carry = [L2T, L2S1, T3, L3Tp, S13]
        L3Tp_L2.append(L3Tp)
    L2T, L3Tp = carry[0], carry[3]

    # ---- Level-1 corrector (cF:3432-3456) ----
    L2Tp, L1T = getNewTprime(Levels[2], L2T, L1T, Levels[1], LInterp[0])
    L1V = computeL1TprimeTerms_Part2(
        Levels, ne_nn, L3Tp, L2Tp, L3rhocp_L1, L2rhocp_L1, dt_all, Shapes, L1V
    )
    L1T = computeL1Temperature(Levels, ne_nn, tmp_ne_nn, L1F, L1V, L1k, L1rhocp, dt_all, properties)
    L1T = np.maximum(T_amb, L1T)

    # ---- corrector pass (cF:3458-3622) ----
    carry = [
        Levels[2]["T0"],
        Levels[2]["S1"],
        Levels[3]["T0"],
        Levels[3]["Tprime0"],
        Levels[3]["S1"],
        Levels[3]["S2"],
        np.asarray(max_accum_L3, dtype=FDT),
        np.asarray(accum_L3, dtype=FDT),
    ]
    L2all, L3all, L3pall = [], [], []
    for _L2sub in range(N2):
        a2, b2, L3rhocp_L2, L2S1, L2k, L2rhocp, L2F, L2V, dt2 = L2_common(carry, _L2sub)
        L2V = computeL2TprimeTerms_Part2(
            Levels, ne_nn, L3Tp_L2[_L2sub], carry[3], L3rhocp_L2, dt2, Shapes, L2V
        )
        _BC = a2 * L1T + b2 * Levels[1]["T0"]
        L2T = computeL2Temperature(
            _BC, LInterp[0], Levels, ne_nn, carry[0], L2F, L2V, L2k, L2rhocp, dt2
        )
        L2T = np.maximum(T_amb, L2T)
        T3, S13, S23, mx, ac = carry[2], carry[4], carry[5], carry[6], carry[7]
        for _L3sub in range(N3):
            T3n, S13n, L3S2, LLidx = L3_substep(T3, S13, _L3sub, _L2sub, L2T, carry[0])
            # cF:3568-3578: melt-time bookkeeping
            _resetmask = ((1 - 2 * S23.astype(np.int32)) * L3S2.astype(np.int32)) == 1
            _resetaccumtime = ac * _resetmask
            mx = np.maximum(_resetaccumtime, mx)
            ac = (ac + laser_position[LLidx, 5] * L3S2 - _resetaccumtime).astype(FDT)
            T3, S13, S23 = T3n, S13n, L3S2
        L3Tp, L2T = getNewTprime(Levels[3], T3, L2T, Levels[2], LInterp[1])
        carry = [L2T, L2S1, T3, L3Tp, S13, S23, mx, ac]
        if return_history:
            L2all.append(L2T)
            L3all.append(T3)
            L3pall.append(L3Tp)
    (
        Levels[2]["T0"],
        Levels[2]["S1"],
        Levels[3]["T0"],
        Levels[3]["Tprime0"],
        Levels[3]["S1"],
        Levels[3]["S2"],
        max_accum_L3,
        accum_L3,
    ) = carry
    Levels[2]["Tprime0"], Levels[1]["T0"] = getNewTprime(
        Levels[2], Levels[2]["T0"], L1T, Levels[1], LInterp[0]
    )
    _scatter_L0(Levels)
    return Levels, L2all, L3all, L3pall, max_accum_L3, accum_L3


def moveEverything(v, vstart, Levels, move_v, LInterp, L1L2Eratio, L2L3Eratio, height):
    """cF:2400-2510: integer-cell shift of the L3 / L2 windows, re-interpolation of T0 / T'0,
    overlap bookkeeping, S1/S2 regather from Level 0, rebuild of the transfer operators."""
    FDT = config.FDT
    v = np.asarray(v, dtype=FDT)
    vstart = np.asarray(vstart, dtype=FDT)
    vtot = v - vstart
    v_L3 = jit_constrain_v(vtot, Levels[3])
    new_coords_L3, _ = move_fine_mesh(Levels[3]["init_node_coors"], Levels[2]["h"], v_L3)
    Levels[3] = update_overlap_nodes_coords(Levels[3], v_L3, Levels[2]["h"], [1, 1, 1])
    T3 = interpolatePoints(Levels[1], Levels[1]["T0"], new_coords_L3)
    Tprime_L2 = interpolatePoints(Levels[2], Levels[2]["Tprime0"], new_coords_L3)
    Tprime_L3 = interpolatePoints(Levels[3], Levels[3]["Tprime0"], new_coords_L3)
    Levels[3]["T0"] = T3 + (Tprime_L2 + Tprime_L3)
    Levels[3]["Tprime0"] = Tprime_L3
    Levels[3]["node_coords"] = new_coords_L3

    v_L2 = jit_constrain_v(vtot, Levels[2])
    h_L1 = list(Levels[1]["h"][:2]) + [FDT(height)]
    new_coords_L2, move_v = move_fine_mesh(Levels[2]["init_node_coors"], h_L1, v_L2)
    move_v = [move_v[i] * int(L1L2Eratio[i]) for i in range(3)]
    Levels[2] = update_overlap_nodes_coords_L1L2(Levels[2], v_L2, Levels[1]["h"], height)
    T2 = interpolatePoints(Levels[1], Levels[1]["T0"], new_coords_L2)
    Levels[2]["Tprime0"] = interpolatePoints(Levels[2], Levels[2]["Tprime0"], new_coords_L2)
    Levels[2]["T0"] = T2 + Levels[2]["Tprime0"]
    Levels[2]["node_coords"] = new_coords_L2

    L2L1Shape = computeCoarseFineShapeFunctions(Levels[1], Levels[2])
    LInterp = list(LInterp)
    LInterp[0] = interpolatePointsMatrix(Levels[1], new_coords_L2)

    Levels[3]["overlapNodes"] = [Levels[3]["overlapNodes"][i] - move_v[i] for i in range(3)]

    Levels[0] = update_overlap_nodes_coords(Levels[0], v_L3, Levels[2]["h"], L2L3Eratio)
    Levels[0] = update_overlap_nodes_coords_L2(
        Levels[0],
        v_L2,
        [Levels[1]["h"][0], Levels[1]["h"][1], Levels[2]["h"][2]],
        [L1L2Eratio[0] * L2L3Eratio[0], L1L2Eratio[1] * L2L3Eratio[1], L2L3Eratio[2]],
    )
    Levels[0]["overlapNodes"][2] = Levels[0]["overlapNodes"][2] - move_v[2] * int(L2L3Eratio[2])
    Levels[0]["overlapNodes_L2"][2] = (
        Levels[0]["overlapNodes_L2"][2] - move_v[2] * int(L2L3Eratio[2])
    )
    Levels[0]["idx"] = getOverlapRegion(
        Levels[0]["overlapNodes"], Levels[0]["nodes"][0], Levels[0]["nodes"][1]
    )
    Levels[0]["idx_L2"] = getOverlapRegion(
        Levels[0]["overlapNodes_L2"], Levels[0]["nodes"][0], Levels[0]["nodes"][1]
    )
    Levels[2]["S1"] = np.array(Levels[0]["S1"][Levels[0]["idx_L2"]], dtype=FDT)
    Levels[3]["S1"] = np.array(Levels[0]["S1"][Levels[0]["idx"]], dtype=FDT)
    Levels[3]["S2"] = np.array(Levels[0]["S2"][Levels[0]["idx"]], dtype=bool)

    L3L1Shape = computeCoarseFineShapeFunctions(Levels[1], Levels[3])
    L3L2Shape = computeCoarseFineShapeFunctions(Levels[2], Levels[3])
    LInterp[1] = interpolatePointsMatrix(Levels[2], new_coords_L3)
    Shapes = [L2L1Shape, L3L1Shape, L3L2Shape]
    return Levels, Shapes, LInterp, move_v


def melting_temp(temps, delt_T, T_melt, accum_time, idx):
    """cF:3696-3712."""
    FDT = config.FDT
    above = np.asarray(temps) > FDT(T_melt)
    out = np.array(accum_time, dtype=FDT, copy=True)
    out[idx] = out[idx] + above * FDT(delt_T)
    return out


def levelMaxMin(Ls):
    """cF:3635-3665 without the prints / exit: [(min, max, ok)] for levels 1.."""
    res = []
    for i in range(1, len(Ls)):
        T = Ls[i]["T0"]
        lo, hi = float(np.min(T)), float(np.max(T))
        ok = all(np.isfinite(v) and 0 < v <= 1e5 for v in (lo, hi))
        res.append((lo, hi, ok))
    return res
