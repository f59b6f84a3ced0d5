"""Analytic known-answer tests of the oracle: properties that follow from the reference's
formulas themselves, independent of any execution of the reference.  CPU only, NumPy float32.

K1 uniform T stays uniform (element stiffness rows sum to zero); K2 a linear T with uniform k leaves the interior
unchanged; K3 energy balance sum_n M[n] (T_new - T)[n] = sum_n (F + Corr)[n] with natural faces; K4 the Gaussian
source integrates to P * eta over the half space below the laser; K5 interpolatePoints reproduces a trilinear field
and returns 0 outside the source box; K6 inject(prolong(u)) = u on the coincident nodes of nested grids;
K7 the projected source on a parent sums to the source on Level 3 (partition of unity of the parent shape functions)."""
import copy
import os
import sys

import numpy as np

from oracle import computeFunctions as cF
from oracle.util import make_level

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import scenario  # noqa: E402

F32 = np.float32
PROPS = cF.SetupProperties(dict(scenario.SMALL_INPUT["properties"]))


def _level(elements=(9, 7, 5), h=0.02):
    ex, ey, ez = elements
    return make_level(elements, ((0.0, ex * h), (0.0, ey * h), (-ez * h, 0.0)))


def _interior(lv):
    nx, ny, nz = lv["nodes"]
    m = np.zeros((nz, ny, nx), bool)
    m[1:-1, 1:-1, 1:-1] = True
    return m.ravel()


def _lumped_mass(lv, rhocp, dt):
    """sum_e Me per node, read off the solver: with k = 0 and T = 1 the update is (M * 1 + F) / M, so M = F / (T - 1)
    for a unit load - simpler: with T = 0, F = 1: T_new = 1 / M."""
    nn, ne = lv["nn"], lv["ne"]
    Tn = cF.solveMatrixFreeFE(lv, nn, ne, np.zeros(nn, F32), rhocp, dt, np.zeros(nn, F32), np.ones(nn, F32), 0)
    return (1.0 / Tn.astype(np.float64))


def test_K1_uniform_temperature_is_a_fixed_point():
    lv = _level()
    nn, ne = lv["nn"], lv["ne"]
    rng = np.random.default_rng(0)
    k = (0.01 + rng.random(nn) * 0.03).astype(F32)       # any conductivity field
    rc = (3e-3 + rng.random(nn) * 2e-3).astype(F32)
    T = np.full(nn, 1234.5, F32)
    Tn = cF.solveMatrixFreeFE(lv, nn, ne, k, rc, 1e-5, T, np.zeros(nn, F32), 0)
    assert float(np.max(np.abs(Tn - T))) <= 4 * np.spacing(F32(1234.5))   # f32 round-off of the M T - K T form: 2 ulp seen


def test_K2_linear_temperature_with_uniform_conductivity_keeps_the_interior():
    lv = _level((10, 8, 6))
    nn, ne = lv["nn"], lv["ne"]
    x, y, z = lv["node_coords"]
    T = (900.0 + 400.0 * x[None, None, :] - 250.0 * y[None, :, None] + 300.0 * z[:, None, None]).astype(F32).ravel()
    k = np.full(nn, 0.02, F32)
    rc = np.full(nn, 4e-3, F32)
    Tn = cF.solveMatrixFreeFE(lv, nn, ne, k, rc, 1e-5, T, np.zeros(nn, F32), 0)
    inner = _interior(lv)
    assert float(np.max(np.abs(Tn[inner] - T[inner]))) <= 3e-4
    assert float(np.max(np.abs(Tn[~inner] - T[~inner]))) > 1e-2   # the natural faces do see the gradient


def test_K3_energy_balance_with_natural_faces():
    lv = _level((8, 8, 6))
    nn, ne = lv["nn"], lv["ne"]
    rng = np.random.default_rng(1)
    T = (600.0 + 800.0 * rng.random(nn)).astype(F32)
    k = (0.01 + 0.02 * rng.random(nn)).astype(F32)
    rc = (3e-3 + 2e-3 * rng.random(nn)).astype(F32)
    F = (rng.random(nn) * 2e-2).astype(F32)
    dt = 1e-5
    Tn = cF.solveMatrixFreeFE(lv, nn, ne, k, rc, dt, T, F, 0)
    M = _lumped_mass(lv, rc, dt)
    lhs = float(np.sum(M * (Tn.astype(np.float64) - T.astype(np.float64))))
    rhs = float(np.sum(F.astype(np.float64)))
    assert abs(lhs - rhs) <= 2e-3 * abs(rhs), (lhs, rhs)    # K's columns sum to zero: conduction moves heat, adds none


def test_K4_source_integrates_to_absorbed_power_below_the_surface():
    r = float(PROPS["laser_radius"])
    h = r / 8.0
    n = int(round(6 * r / h))                                 # +-3 r laterally, 3 d deep: exp(-27) beyond
    lv = make_level((2 * n, 2 * n, n), ((-n * h, n * h), (-n * h, n * h), (-n * h, 0.0)))
    v = np.array([0.0, 0.0, 0.0], F32)
    F = cF.computeSourcesL3(lv, v, (0, lv["ne"], 0, 0, lv["nn"]), PROPS, 285.0)
    total = float(np.sum(F.astype(np.float64)))
    want = 285.0 * float(PROPS["laser_eta"])                  # half of 2 P eta (cF:1014-1025)
    assert abs(total - want) <= 2e-3 * want, (total, want)
    assert float(F.min()) >= 0.0


def test_K5_interpolation_reproduces_trilinear_fields_and_is_zero_outside():
    lv = _level((6, 5, 4), h=0.05)
    x, y, z = lv["node_coords"]
    f = lambda X, Y, Z: (300.0 + 70.0 * X - 40.0 * Y + 90.0 * Z + 500.0 * X * Y - 300.0 * Y * Z + 800.0 * X * Y * Z)
    u = f(x[None, None, :], y[None, :, None], z[:, None, None]).astype(F32).ravel()
    rng = np.random.default_rng(2)
    xs = np.sort(rng.uniform(x[0], x[-1], 7)).astype(F32)
    ys = np.sort(rng.uniform(y[0], y[-1], 6)).astype(F32)
    zs = np.sort(rng.uniform(z[0], z[-1], 5)).astype(F32)
    got = cF.interpolatePoints(lv, u, [xs, ys, zs]).reshape(5, 6, 7)
    want = f(xs[None, None, :].astype(np.float64), ys[None, :, None].astype(np.float64), zs[:, None, None].astype(np.float64))
    assert float(np.max(np.abs(got - want) / np.abs(want))) <= 3e-6
    out = cF.interpolatePoints(lv, u, [np.array([x[-1] + 0.1], F32), ys, zs])
    assert (np.asarray(out) == 0).all()                       # cF:1101-1102: weights zeroed outside the box


def test_K6_inject_after_prolong_is_the_identity_on_coincident_nodes():
    coarse = _level((4, 4, 3), h=0.04)
    xc, yc, zc = coarse["node_coords"]
    fine = make_level((8, 8, 6), ((xc[0], xc[-1]), (yc[0], yc[-1]), (zc[0], zc[-1])))
    rng = np.random.default_rng(3)
    uc = (300.0 + 1000.0 * rng.random(coarse["nn"])).astype(F32)
    uf = cF.interpolatePoints(coarse, uc, fine["node_coords"])               # prolong
    back = cF.interpolatePoints(fine, uf, coarse["node_coords"])             # inject at the coincident nodes
    assert float(np.max(np.abs(back - uc) / uc)) <= 2e-6


def test_K7_projected_source_conserves_the_total_load():
    P = cF.SetupProperties(dict(scenario.SMALL_INPUT["properties"]))
    Levels = cF.SetupLevels(copy.deepcopy(scenario.SMALL_INPUT), P)
    ne_nn = cF.getStaticNodesAndElements(Levels)
    Shapes = [None, cF.computeCoarseFineShapeFunctions(Levels[1], Levels[3]),
              cF.computeCoarseFineShapeFunctions(Levels[2], Levels[3])]
    x3, y3, z3 = [np.asarray(c) for c in Levels[3]["node_coords"]]
    v = np.array([0.5 * (x3[0] + x3[-1]), 0.5 * (y3[0] + y3[-1]), z3[-1]], F32)
    Fc, Fm, Ff = cF.computeSources(Levels[3], v, Shapes, ne_nn, P, 285.0)
    tot = [float(np.sum(np.asarray(a, np.float64))) for a in (Fc, Fm, Ff)]
    assert tot[2] > 0
    assert abs(tot[0] - tot[2]) <= 1e-5 * tot[2] and abs(tot[1] - tot[2]) <= 1e-5 * tot[2], tot


def test_torch_cpu_port_matches_the_numpy_oracle():
    """oracle/torch_cpu.py (the multi-threaded CPU baseline of bench.py: the reference's algorithm as dense tensor
    operations) against the NumPy oracle on a Level-3 substep with a melt pool: temperatures within 1e-5 relative (the
    scatter sums run in a different order), state bit-exact."""
    import torch

    from oracle import computeFunctions as cF
    from oracle.torch_cpu import L3SubstepCPU
    from oracle.util import make_level, smooth_field

    props_in = {"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15, "T_solidus": 1533,
                "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3, "latent_heat_evap": 6457000.0}
    P = cF.SetupProperties(props_in)
    lv = make_level((24, 18, 7), ((1.0, 1.48), (1.0, 1.36), (-0.14, 0.0)))
    rng = np.random.default_rng(3)
    T0 = smooth_field(lv, rng)
    S1 = (rng.random(lv["nn"]) > 0.5).astype(np.float32)
    v, dt, power = np.array([1.2, 1.15, 0.0], np.float32), 1e-5, 285.0
    S1r, _, k, rc = cF.computeStateProperties(T0, S1, P, 0)
    F = cF.computeSourcesL3(lv, v, (0, lv["ne"], 0, 0, lv["nn"]), P, power)
    F = cF.computeConvRadBC(lv, T0, lv["ne"], lv["nn"], P, F)
    want = np.maximum(np.float32(P["T_amb"]), cF.solveMatrixFreeFE(lv, lv["nn"], lv["ne"], k, rc, dt, T0, F, 0))
    port = L3SubstepCPU(lv, P, threads=2)
    got, S1g = port.substep(torch.from_numpy(T0), torch.from_numpy(S1), v, power, dt)
    assert (T0 >= P["T_liquidus"]).sum() > 10
    assert np.array_equal(S1g.numpy(), S1r)
    assert float(np.max(np.abs(got.numpy() - want) / np.abs(want))) <= 1e-5


def test_torch_cpu_dwell_step_matches_the_numpy_oracle():
    """The CPU arm of the N > 1 bench line (oracle/torch_cpu.py dwell_step: stepGOMELTDwellTime cF:2617-2664 as dense
    tensor operations) against the NumPy oracle's flux -> properties -> solve -> assignBCs on a part-scale grid."""
    import torch

    from oracle import computeFunctions as cF
    from oracle.torch_cpu import L3SubstepCPU
    from oracle.util import make_level, smooth_field

    props_in = {"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15, "T_solidus": 1533,
                "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3, "latent_heat_evap": 6457000.0}
    P = cF.SetupProperties(props_in)
    lv = make_level((14, 9, 6), ((0.0, 2.8), (0.0, 1.8), (-1.2, 0.0)))
    rng = np.random.default_rng(4)
    T0 = smooth_field(lv, rng)
    S1 = (rng.random(lv["nn"]) > 0.3).astype(np.float32)
    dt = 2e-3
    bc5 = [301.0, 302.0, 303.0, 304.0, 305.0]
    _, _, k, rc = cF.computeStateProperties(T0, S1, P, 0)
    F = cF.computeConvRadBC(lv, T0, lv["ne"], lv["nn"], P, np.zeros(lv["nn"], np.float32))
    want = cF.solveMatrixFreeFE(lv, lv["nn"], lv["ne"], k, rc, dt, T0, F, 0).reshape(lv["nodes"][2], lv["nodes"][1], lv["nodes"][0])
    want[:, 0, :] = bc5[0]; want[:, -1, :] = bc5[1]; want[:, :, 0] = bc5[2]; want[:, :, -1] = bc5[3]; want[0] = bc5[4]
    got = L3SubstepCPU(lv, P, threads=2).dwell_step(torch.from_numpy(T0), torch.from_numpy(S1), dt, bc5).numpy()
    assert float(np.max(np.abs(got - want.reshape(-1)) / np.abs(want.reshape(-1)))) <= 1e-5
    assert np.abs(got - T0).max() > 0.1
