"""Oracle (test infrastructure): the array updates the reference *driver* performs inline
with jnp (go_melt.py), restated on NumPy so a driver loop can run against the oracle.

Each function cites the gm: lines it follows.
"""
import copy

import numpy as np

from . import config
from .transfer import interpolatePoints, interpolatePointsMatrix


def initAccum(Levels):
    """gm:93-95."""
    nn0 = int(Levels[0]["nn"])
    return np.zeros(nn0, dtype=config.FDT), np.zeros(nn0, dtype=config.FDT)


def findLayerCoords(Levels, laser_z, layer_height):
    """gm:199-211: raise Level-1's original z by layer_height until a plane meets laser z."""
    FDT = config.FDT
    tmp_coords = copy.deepcopy(Levels[1]["orig_node_coords"])
    idx = 0
    while not np.isclose(tmp_coords[2] - FDT(laser_z), 0, atol=1e-4).any():
        tmp_coords[2] = (tmp_coords[2] + FDT(layer_height)).astype(FDT)
        idx += 1
        if idx > 100000:
            raise RuntimeError("laser z never meets a Level-1 plane")
    return tmp_coords, idx


def layerChangeLevel1(Levels, tmp_coords, state_idx, T_amb):
    """gm:214-225: re-interpolate L1 T0 onto the raised grid, rotate S1 through S1_storage."""
    FDT = config.FDT
    Levels[1]["T0"] = np.maximum(
        interpolatePoints(Levels[1], Levels[1]["T0"], tmp_coords), FDT(T_amb)
    )
    st = np.array(Levels[1]["S1_storage"], copy=True)
    st[state_idx - 1, :] = Levels[1]["S1"]  # NB: index -1 when state_idx == 0 (wraps, as jnp does)
    Levels[1]["S1_storage"] = st
    Levels[1]["S1"] = np.array(st[state_idx, :], copy=True)
    Levels[1]["node_coords"] = copy.deepcopy(tmp_coords)
    return Levels


def rebuildInterp(Levels):
    """gm:228-234."""
    return [
        interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
        interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"]),
    ]


def shiftLevel0Down(Levels, laser_z, accum_time, max_accum_time):
    """gm:253-290: accum = max(accum, max_accum); shift L0 S1 and accum down by
    layer_idx_delta planes, zero the freed top planes, move L0 z-coordinates."""
    FDT = config.FDT
    L0 = Levels[0]
    accum_time = np.maximum(accum_time, max_accum_time)
    saved_accum = accum_time.copy()
    nxy = int(L0["nodes"][0]) * int(L0["nodes"][1])
    _0nn1 = nxy * int(L0["layer_idx_delta"])
    _0nn2 = nxy * (int(L0["nodes"][2]) - int(L0["layer_idx_delta"]))
    S1 = np.array(L0["S1"], copy=True)
    S1[:_0nn2] = L0["S1"][_0nn1:]
    S1[_0nn2:] = 0
    L0["S1"] = S1
    L0["node_coords"][2] = (
        L0["orig_node_coords"][2] + FDT(laser_z) - L0["orig_node_coords"][2][-1]
    ).astype(FDT)
    max_accum_time = np.zeros(int(L0["nn"]), dtype=FDT)
    a = accum_time.copy()
    a[:_0nn2] = accum_time[_0nn1:]
    a[_0nn2:] = 0
    return Levels, a, max_accum_time, saved_accum


def fillLevel0Substrate(Levels, substrate):
    """gm:313."""
    S1 = np.array(Levels[0]["S1"], copy=True)
    S1[: int(substrate[0])] = 1
    Levels[0]["S1"] = S1
    return Levels


def accumSingleStep(Levels, all_reset, accum_time, max_accum_time, dt, T_liquidus):
    """gm:339-357 (+ melting_temp cF:3696-3712)."""
    FDT = config.FDT
    idx = Levels[0]["idx"]
    _reset = accum_time[idx] * (np.asarray(all_reset) > 0)
    max_accum_time = np.array(max_accum_time, copy=True)
    max_accum_time[idx] = np.maximum(_reset, max_accum_time[idx])
    accum_time = np.array(accum_time, copy=True)
    accum_time[idx] = accum_time[idx] + (-_reset)
    above = np.asarray(Levels[3]["T0"]) > FDT(T_liquidus)
    accum_time[idx] = accum_time[idx] + above * FDT(dt)
    return accum_time.astype(FDT), max_accum_time.astype(FDT)


def tprimesAllZero(Levels):
    """gm:360-363 predicate (negated): True when either T'0 field is identically zero."""
    return bool((Levels[2]["Tprime0"] == 0).all() or (Levels[3]["Tprime0"] == 0).all())


def zeroTprimes(Levels):
    """gm:367-368."""
    Levels[2]["Tprime0"] = np.zeros_like(Levels[2]["Tprime0"])
    Levels[3]["Tprime0"] = np.zeros_like(Levels[3]["Tprime0"])
    return Levels


def gatherAccum(Levels, accum_time, max_accum_time):
    """gm:448-449."""
    idx = Levels[0]["idx"]
    return max_accum_time[idx], accum_time[idx]


def scatterAccum(Levels, accum_time, max_accum_time, _max_accum, _accum):
    """gm:453-455."""
    idx = Levels[0]["idx"]
    m = np.array(max_accum_time, copy=True)
    a = np.array(accum_time, copy=True)
    m[idx] = _max_accum
    a[idx] = _accum
    return a, m


def finalAccum(accum_time, max_accum_time):
    """gm:512."""
    return np.maximum(accum_time, max_accum_time)


def toHost(x):
    return np.asarray(x)
