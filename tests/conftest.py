import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gm():
    """The product package (gomelt_b200/) with its C-ABI library loaded."""
    import gomelt_b200

    gomelt_b200.load()
    return gomelt_b200


@pytest.fixture(scope="session")
def example_props():
    """properties block of the reference's examples/example.json (ex:76-104), verbatim."""
    return {
        "laser_center": [2.0, 2.0, 0.0, 0, 0, 0, 0],
        "thermal_conductivity_powder": 0.4, "thermal_conductivity_bulk_a0": 4.23,
        "thermal_conductivity_bulk_a1": 0.016, "thermal_conductivity_fluid_a0": 29.0,
        "heat_capacity_solid_a0": 383.1, "heat_capacity_solid_a1": 0.174,
        "heat_capacity_mushy": 3235.0, "heat_capacity_fluid": 769.0, "density": 8e-06,
        "laser_radius": 0.1, "laser_depth": 0.1, "laser_power": 285.0, "laser_absorptivity": 0.45,
        "T_amb": 298.15, "T_solidus": 1533, "T_liquidus": 1609, "T_boiling": 3038.0,
        "h_conv": 1.5e-05, "emissivity": 0.3, "evaporation_coefficient": 0.82,
        "boltzmann_constant": 1.38e-23, "atomic_mass": 9.746e-26, "latent_heat_evap": 6457000.0,
        "molar_mass": 58.69, "layer_height": 0.04,
    }
