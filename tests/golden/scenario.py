"""A short, fixed GO-MELT run expressed against a ``computeFunctions``-like namespace ``cf``.

The same code drives (a) the reference's own source through the NumPy ``jax`` shim
(``make_golden.py`` -> committed ``*.npz``), (b) the oracle (``tests/test_oracle_golden.py``) and
(c) the CUDA product's drop-in namespace (``tests/test_steppers_gpu.py``).  It follows the call
sequence of the reference driver ``go_melt.py`` (gm:134-459) on a scaled-down three-level
problem: layer-start single steps, moving windows, one subcycle block, dwell steps.

Test infrastructure only.
"""
import copy

import numpy as np

SMALL_INPUT = {
    "Level1": {"elements": [10, 4, 6], "bounds": {"x": [0, 2.0], "y": [0, 0.8], "z": [-0.8, 0.4]},
               "conditions": {"x": [298.15, 298.15], "y": [298.15, 298.15], "z": [298.15, 298.15]}},
    "Level2": {"elements": [20, 20, 5], "bounds": {"x": [0, 0.8], "y": [0, 0.8], "z": [-0.2, 0.0]}},
    "Level3": {"elements": [20, 20, 4], "bounds": {"x": [0.2, 0.6], "y": [0.2, 0.6], "z": [-0.08, 0.0]}},
    "properties": {
        "laser_center": [0.4, 0.4, 0.0, 0, 0, 0, 0],
        "thermal_conductivity_powder": 0.4, "thermal_conductivity_bulk_a0": 4.23,
        "thermal_conductivity_bulk_a1": 0.016, "thermal_conductivity_fluid_a0": 29.0,
        "heat_capacity_solid_a0": 383.1, "heat_capacity_solid_a1": 0.174, "heat_capacity_mushy": 3235.0,
        "heat_capacity_fluid": 769.0, "density": 8e-06, "laser_radius": 0.1, "laser_depth": 0.1,
        "laser_power": 285.0, "laser_absorptivity": 0.45, "T_amb": 298.15, "T_solidus": 1533,
        "T_liquidus": 1609, "T_boiling": 3038.0, "h_conv": 1.5e-05, "emissivity": 0.3,
        "evaporation_coefficient": 0.82, "boltzmann_constant": 1.38e-23, "atomic_mass": 9.746e-26,
        "latent_heat_evap": 6457000.0, "molar_mass": 58.69, "layer_height": 0.04,
    },
    "nonmesh": {"timestep_L3": 1e-5, "subcycle_num_L2": 2, "subcycle_num_L3": 2, "dwell_time": 0.1,
                "Level1_record_step": 1, "output_files": 0, "wait_time": 500, "layer_num": 0,
                "restart_layer_num": 10000, "info_T": 0, "laser_velocity": 500, "record_step": 1000},
}


# Two layers of three serpentine tracks joined by rapid moves (tests/golden/toolpath_serpentine.txt is the reference
# parser's output for it, see make_golden.py)
SERPENTINE_GCODE = ("G0 X0.40 Y0.36 Z0.04\nG1 X0.88 Y0.36 Z0.04\nG0 X0.88 Y0.40 Z0.04\nG1 X0.40 Y0.40 Z0.04\n"
                    "G0 X0.40 Y0.44 Z0.04\nG1 X0.88 Y0.44 Z0.04\n"
                    "G0 X0.88 Y0.44 Z0.08\nG1 X0.40 Y0.44 Z0.08\nG0 X0.40 Y0.40 Z0.08\nG1 X0.88 Y0.40 Z0.08\n"
                    "G0 X0.88 Y0.36 Z0.08\nG1 X0.40 Y0.36 Z0.08\n")
SERPENTINE_NONMESH = {"timestep_L3": 1e-05, "dwell_time": 0.002, "wait_time": 20, "output_files": 1, "record_step": 25,
                      "Level1_record_step": 1, "laser_velocity": 800, "layer_num": 0, "subcycle_num_L2": 3,
                      "subcycle_num_L3": 4, "info_T": 0, "dwell_time_multiplier": 5, "use_txt": 0}


def toolpath_rows():
    """x, y, z, Ljump, Ldwell, dt, P (cP:71-74).  Large x increments so that the windows shift
    within a handful of steps; dt kept at the physical 1e-5 s."""
    rows = []
    x = 0.4
    for i in range(3):                       # single steps
        rows.append([x, 0.4, 0.0, 1, 1, 1e-5, 285.0])
        x += 0.045
    for i in range(4):                       # one N2 x N3 = 2 x 2 subcycle block
        rows.append([x, 0.4 + 0.012 * i, 0.0, 1, 1, 1e-5, 285.0])
        x += 0.006
    for i in range(2):                       # dwell rows
        rows.append([x, 0.4, 0.0, 0, 0, 2e-3, 0.0])
    return np.array(rows, dtype=np.float32)


def _np(x):
    if hasattr(x, "detach"):  # torch tensor (CUDA product)
        return x.detach().cpu().numpy()
    return np.asarray(x)


def snapshot(Levels, extra=None):
    out = {
        "L1_T0": _np(Levels[1]["T0"]), "L2_T0": _np(Levels[2]["T0"]), "L3_T0": _np(Levels[3]["T0"]),
        "L2_Tp0": _np(Levels[2]["Tprime0"]), "L3_Tp0": _np(Levels[3]["Tprime0"]),
        "L1_S1": _np(Levels[1]["S1"]), "L2_S1": _np(Levels[2]["S1"]), "L3_S1": _np(Levels[3]["S1"]),
        "L3_S2": _np(Levels[3]["S2"]).astype(np.uint8), "L0_S1": _np(Levels[0]["S1"]),
        "L0_S2": _np(Levels[0]["S2"]).astype(np.uint8),
        "L3_x": _np(Levels[3]["node_coords"][0]), "L2_x": _np(Levels[2]["node_coords"][0]),
        "L0_idx": _np(Levels[0]["idx"]),
    }
    if extra:
        out.update({k: _np(v) for k, v in extra.items()})
    return {k: np.array(v, copy=True) for k, v in out.items()}


def run(cf, wrap=lambda a: a, save_path="/tmp/gomelt_golden/", inp=None, hooks=None):
    """Returns {phase: {name: array}}.  ``wrap`` re-types a NumPy array for ``cf`` (identity for the
    oracle / product, the shim's array class for the reference)."""
    inp = copy.deepcopy(inp or SMALL_INPUT)
    inp["nonmesh"]["save_path"] = save_path
    inp["nonmesh"]["toolpath"] = save_path + "toolpath.txt"
    P = cf.SetupProperties(inp["properties"])
    Levels = cf.SetupLevels(inp, P)
    N = cf.SetupNonmesh(inp["nonmesh"])
    ne_nn = cf.getStaticNodesAndElements(Levels)
    subcycle = cf.getStaticSubcycle(N)
    L1L2E = [int(np.round(float(Levels[1]["h"][i]) / float(Levels[2]["h"][i]))) for i in range(2)] + [
        int(np.round(P["layer_height"] / float(Levels[2]["h"][2])))]
    L2L3E = [int(np.round(float(Levels[2]["h"][i]) / float(Levels[3]["h"][i]))) for i in range(3)]
    laser_start = np.array(P["laser_center"])
    LInterp = [cf.interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
               cf.interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"])]
    nn0 = int(Levels[0]["nn"])
    accum = wrap(np.zeros(nn0, np.float32))
    max_accum = wrap(np.zeros(nn0, np.float32))
    move_hist = [wrap(np.array(0)), wrap(np.array(0)), wrap(np.array(0))]
    rows = toolpath_rows()
    out = {}

    # ---- single-step rows (gm:177-388); row 0 runs the layer-start path (gm:198-290) ----------
    laser_prev_z = float("inf")
    for irow in range(3):
        lp = wrap(rows[irow])
        if float(lp[2]) != laser_prev_z:
            tmp_coords = copy.deepcopy(Levels[1]["orig_node_coords"])
            state_idx = 0
            while not np.isclose(_np(tmp_coords[2]) - float(lp[2]), 0, atol=1e-4).any():
                tmp_coords[2] = wrap((_np(tmp_coords[2]) + np.float32(P["layer_height"])).astype(np.float32))
                state_idx += 1
            T0i = cf.interpolatePoints(Levels[1], Levels[1]["T0"], tmp_coords)
            Levels[1]["T0"] = wrap(np.maximum(_np(T0i), np.float32(P["T_amb"])))
            st = np.array(_np(Levels[1]["S1_storage"]), copy=True)
            st[state_idx - 1, :] = _np(Levels[1]["S1"])
            Levels[1]["S1_storage"] = wrap(st)
            Levels[1]["S1"] = wrap(st[state_idx, :].copy())
            Levels[1]["node_coords"] = copy.deepcopy(tmp_coords)
            LInterp = [cf.interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
                       cf.interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"])]
            tmp_ne_nn = cf.calcStaticTmpNodesAndElements(Levels, lp)
            laser_prev_z = float(lp[2])
            # gm:253-290 Level-0 shift
            accum = wrap(np.maximum(_np(accum), _np(max_accum)))
            nxy = int(Levels[0]["nodes"][0]) * int(Levels[0]["nodes"][1])
            n1 = nxy * int(Levels[0]["layer_idx_delta"])
            n2 = nxy * (int(Levels[0]["nodes"][2]) - int(Levels[0]["layer_idx_delta"]))
            S1 = np.array(_np(Levels[0]["S1"]), copy=True)
            S1[:n2] = _np(Levels[0]["S1"])[n1:]
            S1[n2:] = 0
            Levels[0]["S1"] = wrap(S1)
            Levels[0]["node_coords"][2] = wrap((_np(Levels[0]["orig_node_coords"][2]) + np.float32(lp[2])
                                                - _np(Levels[0]["orig_node_coords"][2])[-1]).astype(np.float32))
            max_accum = wrap(np.zeros(nn0, np.float32))
            a = np.array(_np(accum), copy=True)
            a[:n2] = _np(accum)[n1:]
            a[n2:] = 0
            accum = wrap(a)
            move_vert = True
        Levels, Shapes, LInterp, move_hist = cf.moveEverything(lp, laser_start, Levels, move_hist, LInterp, L1L2E,
                                                               L2L3E, P["layer_height"])
        if move_vert:
            move_vert = False
            substrate = cf.getSubstrateNodes(Levels)
            S1 = np.array(_np(Levels[0]["S1"]), copy=True)
            S1[: int(substrate[0])] = 1
            Levels[0]["S1"] = wrap(S1)
        if hooks and "before_step" in hooks:
            hooks["before_step"](irow, Levels, Shapes, LInterp, ne_nn, tmp_ne_nn, substrate, P)
        Levels, all_reset = cf.stepGOMELT(Levels, ne_nn, tmp_ne_nn, Shapes, LInterp, lp, P, lp[5], lp[6], substrate)
        # gm:338-357
        idx = _np(Levels[0]["idx"])
        a, m = np.array(_np(accum), copy=True), np.array(_np(max_accum), copy=True)
        reset = a[idx] * (_np(all_reset) > 0)
        m[idx] = np.maximum(reset, m[idx])
        a[idx] = a[idx] + (-reset)
        accum, max_accum = wrap(a.astype(np.float32)), wrap(m.astype(np.float32))
        accum = cf.melting_temp(Levels[3]["T0"], lp[5], P["T_liquidus"], accum, Levels[0]["idx"])
        out[f"step{irow}"] = snapshot(Levels, {"accum": accum, "max_accum": max_accum, "all_reset": _np(all_reset).astype(np.uint8),
                                                "move_hist": np.array([int(v) for v in move_hist])})

    # ---- one subcycle block (gm:413-459) ---------------------------------------------------------
    laser_all = wrap(rows[3:7])
    Levels, Shapes, LInterp, move_hist = cf.moveEverything(laser_all[0, :], laser_start, Levels, move_hist, LInterp,
                                                           L1L2E, L2L3E, P["layer_height"])
    idx = _np(Levels[0]["idx"])
    res = cf.subcycleGOMELT(Levels, ne_nn, Shapes, substrate, LInterp, tmp_ne_nn, laser_all, P, laser_all[:, 6],
                            subcycle, wrap(_np(max_accum)[idx]), wrap(_np(accum)[idx]))
    Levels, _max, _acc = res[0], res[4], res[5]
    a, m = np.array(_np(accum), copy=True), np.array(_np(max_accum), copy=True)
    m[idx] = _np(_max)
    a[idx] = _np(_acc)
    accum, max_accum = wrap(a), wrap(m)
    out["subcycle"] = snapshot(Levels, {"accum": accum, "max_accum": max_accum,
                                        "move_hist": np.array([int(v) for v in move_hist])})

    # ---- dwell rows (gm:358-380): Level 1 only ---------------------------------------------------------
    for irow in (7, 8):
        lp = wrap(rows[irow])
        Levels, Shapes, LInterp, move_hist = cf.moveEverything(lp, laser_start, Levels, move_hist, LInterp, L1L2E,
                                                               L2L3E, P["layer_height"])
        if not (_np(Levels[2]["Tprime0"]) == 0).all() and not (_np(Levels[3]["Tprime0"]) == 0).all():
            Levels[2]["Tprime0"] = wrap(np.zeros_like(_np(Levels[2]["Tprime0"])))
            Levels[3]["Tprime0"] = wrap(np.zeros_like(_np(Levels[3]["Tprime0"])))
        Levels = cf.stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, P, lp[5], substrate)
        out[f"dwell{irow}"] = snapshot(Levels)
    out["meta"] = {"ne_nn": np.array([int(v) for v in ne_nn]), "tmp_ne_nn": np.array([int(v) for v in tmp_ne_nn]),
                   "substrate": np.array([int(v) for v in substrate]), "rows": rows}
    return out
