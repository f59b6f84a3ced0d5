"""Oracle (test infrastructure): JSON -> Properties / Levels / Nonmesh, as the reference
builds them.  Restates SetupLevels cF:115-264 - SetupProperties cF:267-345 - SetupNonmesh
cF:348-419 - getStaticNodesAndElements cF:447-470 - getStaticSubcycle cF:473-492 -
calcStaticTmpNodesAndElements cF:495-517 - getBCindices cF:520-559 - getSubstrateNodes
cF:562-579 - getCoarseNodesIn(Large)FineRegion cF:858-925 - find_max_const cF:1886-1913 -
calc_length_h cF:1916-1942.  Levels are plain dicts of NumPy arrays with the reference's
field names (cF:258-263).
"""
import copy
import os

import numpy as np

from . import config
from .fem import createMesh3D
from .transfer import getOverlapRegion


def calcNumNodes(elements):
    """cF:18-29."""
    return [elements[0] + 1, elements[1] + 1, elements[2] + 1]


def calc_length_h(A):
    """cF:1916-1942 (Python float arithmetic, as in the reference)."""
    b = A["bounds"]
    Lx = b["x"][1] - b["x"][0]
    Ly = b["y"][1] - b["y"][0]
    Lz = b["z"][1] - b["z"][0]
    return [Lx, Ly, Lz], [Lx / A["elements"][0], Ly / A["elements"][1], Lz / A["elements"][2]]


def getBCindices(nodes, nn):
    """cF:520-559: [west, east, south, north, bottom, top] node id arrays."""
    nx, ny, nz = nodes
    bidx = np.arange(0, nx * ny)
    tidx = np.arange(nx * ny * (nz - 1), nn)
    widx = np.arange(0, nn, nx)
    eidx = np.arange(nx - 1, nn, nx)
    sidx = (np.arange(0, nx)[:, None] + (nx * ny * np.arange(0, nz))[None, :]).reshape(-1)
    nidx = (
        np.arange(nx * (ny - 1), nx * ny)[:, None] + (nx * ny * np.arange(0, nz))[None, :]
    ).reshape(-1)
    return [widx, eidx, sidx, nidx, bidx, tidx]


def find_max_const(CoarseLevel, FinerLevel):
    """cF:1886-1913 (Python floats)."""
    cb, fb = CoarseLevel["bounds"], FinerLevel["bounds"]
    iE = cb["x"][1] - fb["x"][1]
    iN = cb["y"][1] - fb["y"][1]
    iT = cb["z"][1] - fb["z"][1]
    iW = cb["x"][0] - fb["x"][0]
    iS = cb["y"][0] - fb["y"][0]
    iB = cb["z"][0] - fb["z"][0]
    return [iW, iE], [iS, iN], [iB, iT]


def getCoarseNodesInFineRegion(xnf, xnc):
    """cF:858-887."""
    FDT = config.FDT
    xfmin, xfmax, xcmin, xcmax = xnf.min(), xnf.max(), xnc.min(), xnc.max()
    nec = xnc.size - 1
    hc = (xcmax - xcmin) / FDT(nec)
    overlapMin = np.round((xfmin - xcmin) / hc)
    overlapMax = np.round((xfmax - xcmin) / hc) + 1
    return np.arange(overlapMin, overlapMax).astype(int)


def getCoarseNodesInLargeFineRegion(xnc, xnf):
    """cF:890-925."""
    FDT = config.FDT
    xfmin, xfmax, xcmin, xcmax = xnf.min(), xnf.max(), xnc.min(), xnc.max()
    hf = (xfmax - xfmin) / FDT(xnf.size - 1)
    hc = (xcmax - xcmin) / FDT(xnc.size - 1)
    overlapMin = np.round((xcmin - xfmin) / hf)
    overlapMax = np.round((xcmax - xfmin) / hf) + 1
    step = int(np.round(hc / hf))
    return np.arange(overlapMin, overlapMax, step).astype(int)


def SetupProperties(prop_obj):
    """cF:267-345.  Values stay Python floats here (they become f32 on entry to each jitted
    function in the reference; the oracle casts with fem.f32props at the same places)."""
    p = dict(copy.deepcopy(prop_obj))
    g = prop_obj.get
    p["k_powder"] = g("thermal_conductivity_powder", 0.4)
    p["k_bulk_coeff_a0"] = g("thermal_conductivity_bulk_a0", 4.23)
    p["k_bulk_coeff_a1"] = g("thermal_conductivity_bulk_a1", 0.016)
    p["k_fluid_coeff_a0"] = g("thermal_conductivity_fluid_a0", 29.0)
    p["cp_solid_coeff_a0"] = g("heat_capacity_solid_a0", 383.1)
    p["cp_solid_coeff_a1"] = g("heat_capacity_solid_a1", 0.174)
    p["cp_mushy"] = g("heat_capacity_mushy", 3235.0)
    p["cp_fluid"] = g("heat_capacity_fluid", 769.0)
    p["rho"] = g("density", 8.0e-6)
    p["laser_radius"] = g("laser_radius", 0.110)
    p["laser_depth"] = g("laser_depth", 0.05)
    p["laser_power"] = g("laser_power", 300.0)
    p["laser_eta"] = g("laser_absorptivity", 0.25)
    p["laser_center"] = g("laser_center", [])
    p["T_amb"] = g("T_amb", 353.15)
    p["T_solidus"] = g("T_solidus", 1554.0)
    p["T_liquidus"] = g("T_liquidus", 1625.0)
    p["T_boiling"] = g("T_boiling", 3038.0)
    p["h_conv"] = g("h_conv", 1.473e-5)
    p["h_conv"] *= 1e6  # cF:319 "quick fix"
    p["vareps"] = g("emissivity", 0.600)
    p["evc"] = g("evaporation_coefficient", 0.82)
    p["kb"] = g("boltzmann_constant", 1.38e-23)
    p["mA"] = g("atomic_mass", 7.9485017e-26)
    p["Lev"] = g("latent_heat_evap", 4.22e6)
    p["molar_mass"] = g("molar_mass", 58.69) * 1e-3
    p["layer_height"] = g("layer_height", 0.04)
    p["sigma_sb"] = 5.67e-8
    p["gas_const"] = 8.314
    p["atmospheric_pressure"] = 101325
    p["CM_coeff"] = p["molar_mass"] / (2.0 * np.pi * p["gas_const"])
    p["CT_coeff"] = p["Lev"] * p["molar_mass"] / p["gas_const"]
    p["CP_coeff"] = 0.54 * p["atmospheric_pressure"]
    return p


def SetupNonmesh(nonmesh_input, make_dirs=True):
    """cF:348-419."""
    n = dict(copy.deepcopy(nonmesh_input))
    g = nonmesh_input.get
    n["timestep_L3"] = g("timestep_L3", 1e-5)
    n["subcycle_num_L2"] = g("subcycle_num_L2", 1)
    n["subcycle_num_L3"] = g("subcycle_num_L3", 1)
    n["dwell_time"] = g("dwell_time", 0.1)
    n["Level1_record_step"] = g("Level1_record_step", 1)
    n["save_path"] = g("save_path", "results/")
    n["output_files"] = g("output_files", 1)
    n["toolpath"] = g("toolpath", "laserPath.txt")
    n["wait_time"] = g("wait_time", 500.0)
    n["layer_num"] = g("layer_num", 0)
    n["restart_layer_num"] = g("restart_layer_num", 10000)
    n["info_T"] = g("info_T", 0)
    n["laser_velocity"] = g("laser_velocity", 500)
    n["wait_track"] = g("wait_track", 0.0)
    n["record_step"] = g("record_step", n["subcycle_num_L2"] * n["subcycle_num_L3"])
    n["gcode"] = g("gcode", "./examples/gcodefiles/defaultName.gcode")
    n["dwell_time_multiplier"] = g("dwell_time_multiplier", 1)
    n["use_txt"] = g("use_txt", 0)
    if make_dirs and not os.path.exists(n["save_path"]):
        os.makedirs(n["save_path"])
    return n


def SetupLevels(solver_input, properties):
    """cF:115-264 -> [Level0, Level1, Level2, Level3] dicts."""
    FDT = config.FDT
    L = [None, None, None, None]
    for i, name in ((1, "Level1"), (2, "Level2"), (3, "Level3")):
        L[i] = copy.deepcopy(solver_input.get(name, {}))
    for level in L[1:]:
        level["length"], level["h"] = calc_length_h(level)
        level["nodes"] = calcNumNodes(level["elements"])
        level["ne"] = level["elements"][0] * level["elements"][1] * level["elements"][2]
        level["nn"] = level["nodes"][0] * level["nodes"][1] * level["nodes"][2]
        level["BC"] = getBCindices(level["nodes"], level["nn"])
        b = level["bounds"]
        level["node_coords"], level["connect"] = createMesh3D(
            (b["x"][0], b["x"][1], level["nodes"][0]),
            (b["y"][0], b["y"][1], level["nodes"][1]),
            (b["z"][0], b["z"][1], level["nodes"][2]),
        )
        nn = level["nn"]
        level["T"] = FDT(properties["T_amb"]) * np.ones(nn, dtype=FDT)
        level["T0"] = FDT(properties["T_amb"]) * np.ones(nn, dtype=FDT)
        level["S1"] = np.zeros(nn, dtype=FDT)
        level["S2"] = np.zeros(nn, dtype=bool)
        level["k"] = FDT(properties["k_powder"]) * np.ones(nn, dtype=FDT)
        level["rhocp"] = FDT(properties["cp_solid_coeff_a0"] * properties["rho"]) * np.ones(
            nn, dtype=FDT
        )
    Level1, Level2, Level3 = L[1], L[2], L[3]
    Level1["S1_storage"] = np.zeros(
        [int(round(Level1["h"][2] / properties["layer_height"])), Level1["nn"]], dtype=FDT
    )
    for level in (Level2, Level3):
        ix, iy, iz = find_max_const(Level1, level)
        level["bounds"]["ix"], level["bounds"]["iy"], level["bounds"]["iz"] = ix, iy, iz
        level["init_node_coors"] = copy.deepcopy(level["node_coords"])
        level["Tprime"] = np.zeros(level["nn"], dtype=FDT)
        level["Tprime0"] = copy.deepcopy(level["Tprime"])

    # cF:175-182: raise Level-1 z by layer_height until a node meets Level-2's top plane
    Level1["orig_node_coords"] = copy.deepcopy(Level1["node_coords"])
    tmp_coords = copy.deepcopy(Level1["orig_node_coords"])
    guard = 0
    while True:
        if np.isclose(tmp_coords[2] - Level2["node_coords"][-1][-1], 0, atol=1e-4).any():
            break
        tmp_coords[2] = (tmp_coords[2] + FDT(properties["layer_height"])).astype(FDT)
        guard += 1
        if guard > 100000:
            raise RuntimeError("Level1/Level2 z-planes never align (reference would spin)")
    Level1["node_coords"] = copy.deepcopy(tmp_coords)

    def get_overlap(level_fine, level_coarse):
        nodes = [
            getCoarseNodesInFineRegion(level_fine["node_coords"][i], level_coarse["node_coords"][i])
            for i in range(3)
        ]
        coors = [
            np.array([level_coarse["node_coords"][i][j] for j in nodes[i]], dtype=FDT)
            for i in range(3)
        ]
        return nodes, coors

    Level2["orig_overlap_nodes"], Level2["orig_overlap_coors"] = get_overlap(Level2, Level1)
    Level3["orig_overlap_nodes"], Level3["orig_overlap_coors"] = get_overlap(Level3, Level2)
    for lv in (Level2, Level3):
        lv["overlapNodes"] = copy.deepcopy(lv["orig_overlap_nodes"])
        lv["overlapCoords"] = copy.deepcopy(lv["orig_overlap_coors"])

    Level0 = {}
    Level0["elements"] = [
        round(Level1["length"][0] / Level3["h"][0]),
        round(Level1["length"][1] / Level3["h"][1]),
        round(Level2["length"][2] / Level3["h"][2]),
    ]
    Level0["nodes"] = calcNumNodes(Level0["elements"])
    Level0["ne"] = Level0["elements"][0] * Level0["elements"][1] * Level0["elements"][2]
    Level0["nn"] = Level0["nodes"][0] * Level0["nodes"][1] * Level0["nodes"][2]
    Level0["node_coords"], Level0["connect"] = createMesh3D(
        (Level1["bounds"]["x"][0], Level1["bounds"]["x"][1], Level0["nodes"][0]),
        (Level1["bounds"]["y"][0], Level1["bounds"]["y"][1], Level0["nodes"][1]),
        (Level2["bounds"]["z"][0], Level2["bounds"]["z"][1], Level0["nodes"][2]),
    )
    Level0["orig_node_coords"] = copy.deepcopy(Level0["node_coords"])
    Level0["orig_overlap_nodes"], Level0["orig_overlap_coors"] = get_overlap(Level3, Level0)
    Level0["overlapNodes"] = copy.deepcopy(Level0["orig_overlap_nodes"])
    Level0["overlapCoords"] = copy.deepcopy(Level0["orig_overlap_coors"])
    Level0["orig_overlap_nodes_L2"] = [
        getCoarseNodesInLargeFineRegion(Level2["node_coords"][i], Level0["node_coords"][i])
        for i in range(3)
    ]
    Level0["orig_overlap_coors_L2"] = [
        np.array(
            [Level0["node_coords"][i][j] for j in Level0["orig_overlap_nodes_L2"][i]], dtype=FDT
        )
        for i in range(3)
    ]
    Level0["overlapNodes_L2"] = copy.deepcopy(Level0["orig_overlap_nodes_L2"])
    Level0["overlapCoords_L2"] = copy.deepcopy(Level0["orig_overlap_coors_L2"])
    Level0["S1"] = np.zeros(Level0["nn"], dtype=FDT)
    Level0["S2"] = np.zeros(Level0["nn"], dtype=bool)
    Level0["idx"] = getOverlapRegion(Level0["overlapNodes"], Level0["nodes"][0], Level0["nodes"][1])
    Level0["idx_L2"] = getOverlapRegion(
        Level0["overlapNodes_L2"], Level0["nodes"][0], Level0["nodes"][1]
    )
    Level0["layer_idx_delta"] = int(round(properties["layer_height"] / Level3["h"][2]))
    L[0] = Level0

    # cF:248-255: numeric fields become (f32 / i32) arrays
    for level in L:
        for attr in ("h", "length"):
            if attr in level:
                level[attr] = [FDT(v) for v in level[attr]]
        for attr in ("elements", "nodes"):
            if attr in level:
                level[attr] = [int(v) for v in level[attr]]
    return L


def getStaticNodesAndElements(L):
    """cF:447-470."""
    return (int(L[2]["ne"]), int(L[3]["ne"]), int(L[1]["nn"]), int(L[2]["nn"]), int(L[3]["nn"]))


def getStaticSubcycle(N):
    """cF:473-492."""
    N2, N3 = N["subcycle_num_L2"], N["subcycle_num_L3"]
    N23 = N2 * N3
    return (N2, N3, N23, float(N2), float(N3), float(N23))


def calcStaticTmpNodesAndElements(L, v):
    """cF:495-517: active Level-1 element / node counts up to z = v[2] (+1e-5)."""
    Level1_mask = L[1]["node_coords"][2] <= config.FDT(v[2]) + config.FDT(1e-5)
    Level1_nn = int(Level1_mask.sum())
    tmp_ne = L[1]["elements"][0] * L[1]["elements"][1] * (Level1_nn - 1)
    tmp_nn = L[1]["nodes"][0] * L[1]["nodes"][1] * Level1_nn
    return (int(tmp_ne), int(tmp_nn))


def getSubstrateNodes(Levels):
    """cF:562-579: nodes with z < 1e-5, per level 0..3."""
    return tuple(
        int((L["node_coords"][2] < 1e-5).sum() * L["nodes"][0] * L["nodes"][1]) for L in Levels[:4]
    )
