"""Input schema of the drop-in: the reference's JSON blocks -> Properties / Nonmesh dicts with the
reference's key names (SetupProperties cF:267-345, SetupNonmesh cF:348-419).  Host-side only.

Table-driven: (reference key, JSON key, default).  Derived quantities follow cF:319-343.
"""
import copy
import math
import os

_PROPERTY_TABLE = (
    ("k_powder", "thermal_conductivity_powder", 0.4),
    ("k_bulk_coeff_a0", "thermal_conductivity_bulk_a0", 4.23),
    ("k_bulk_coeff_a1", "thermal_conductivity_bulk_a1", 0.016),
    ("k_fluid_coeff_a0", "thermal_conductivity_fluid_a0", 29.0),
    ("cp_solid_coeff_a0", "heat_capacity_solid_a0", 383.1),
    ("cp_solid_coeff_a1", "heat_capacity_solid_a1", 0.174),
    ("cp_mushy", "heat_capacity_mushy", 3235.0),
    ("cp_fluid", "heat_capacity_fluid", 769.0),
    ("rho", "density", 8.0e-6),
    ("laser_radius", "laser_radius", 0.110),
    ("laser_depth", "laser_depth", 0.05),
    ("laser_power", "laser_power", 300.0),
    ("laser_eta", "laser_absorptivity", 0.25),
    ("laser_center", "laser_center", []),
    ("T_amb", "T_amb", 353.15),
    ("T_solidus", "T_solidus", 1554.0),
    ("T_liquidus", "T_liquidus", 1625.0),
    ("T_boiling", "T_boiling", 3038.0),
    ("h_conv", "h_conv", 1.473e-5),
    ("vareps", "emissivity", 0.600),
    ("evc", "evaporation_coefficient", 0.82),
    ("kb", "boltzmann_constant", 1.38e-23),
    ("mA", "atomic_mass", 7.9485017e-26),
    ("Lev", "latent_heat_evap", 4.22e6),
    ("molar_mass", "molar_mass", 58.69),
    ("layer_height", "layer_height", 0.04),
)

_NONMESH_TABLE = (
    ("timestep_L3", 1e-5), ("subcycle_num_L2", 1), ("subcycle_num_L3", 1), ("dwell_time", 0.1),
    ("Level1_record_step", 1), ("save_path", "results/"), ("output_files", 1),
    ("toolpath", "laserPath.txt"), ("wait_time", 500.0), ("layer_num", 0),
    ("restart_layer_num", 10000), ("info_T", 0), ("laser_velocity", 500), ("wait_track", 0.0),
    ("record_step", None), ("gcode", "./examples/gcodefiles/defaultName.gcode"),
    ("dwell_time_multiplier", 1), ("use_txt", 0),
)


def SetupProperties(prop_obj):
    """JSON ``properties`` block -> dict with the reference's names (cF:267-345)."""
    out = dict(copy.deepcopy(prop_obj))
    for key, json_key, default in _PROPERTY_TABLE:
        out[key] = prop_obj.get(json_key, default)
    out["h_conv"] = out["h_conv"] * 1e6          # cF:319: W/mm^2 K -> W/m^2 K for the surface term
    out["molar_mass"] = out["molar_mass"] * 1e-3  # cF:327: g/mol -> kg/mol
    out["sigma_sb"] = 5.67e-8
    out["gas_const"] = 8.314
    out["atmospheric_pressure"] = 101325
    out["CM_coeff"] = out["molar_mass"] / (2.0 * math.pi * out["gas_const"])
    out["CT_coeff"] = out["Lev"] * out["molar_mass"] / out["gas_const"]
    out["CP_coeff"] = 0.54 * out["atmospheric_pressure"]
    return out


def SetupNonmesh(nonmesh_input, make_dirs=True):
    """JSON ``nonmesh`` block -> dict with defaults (cF:348-419)."""
    out = dict(copy.deepcopy(nonmesh_input))
    for key, default in _NONMESH_TABLE:
        out[key] = nonmesh_input.get(key, default)
    if nonmesh_input.get("record_step") is None:
        out["record_step"] = out["subcycle_num_L2"] * out["subcycle_num_L3"]
    if make_dirs and not os.path.exists(out["save_path"]):
        os.makedirs(out["save_path"])
    return out


def getStaticSubcycle(N):
    """cF:473-492."""
    n2, n3 = N["subcycle_num_L2"], N["subcycle_num_L3"]
    return (n2, n3, n2 * n3, float(n2), float(n3), float(n2 * n3))
