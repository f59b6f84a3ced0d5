"""Edge-case inputs for the per-function parity pin (shared by make_golden.py, which runs the REFERENCE's functions
on them through the NumPy jax shim, and by tests/test_oracle_golden.py, which runs the oracle on them).

Each case returns a dict of named float32 / bool arrays.  ``cf`` is either the reference's computeFunctions module
(through the shim) or ``oracle.computeFunctions``; ``wrap`` converts NumPy inputs to the array type ``cf`` expects."""
import numpy as np

import scenario

F32 = np.float32


def _props(cf):
    return cf.SetupProperties(dict(scenario.SMALL_INPUT["properties"]))


def _level(cf):
    P = _props(cf)
    return cf.SetupLevels(__import__("copy").deepcopy(scenario.SMALL_INPUT), P), P


def state_properties_at_thresholds(cf, wrap):
    """computeStateProperties cF:2567-2614 on temperatures AT, one ulp below and one ulp above the solidus and the
    liquidus, states at / around the 0.499 threshold, negative zero, and a substrate prefix."""
    P = _props(cf)
    ts, tl = F32(P["T_solidus"]), F32(P["T_liquidus"])
    T = np.array([ts, np.nextafter(ts, F32(0)), np.nextafter(ts, F32(1e9)), tl, np.nextafter(tl, F32(0)),
                  np.nextafter(tl, F32(1e9)), 298.15, 5000.0, 0.5 * (ts + tl), 250.0], F32)
    S = np.array([0.0, 1.0, 0.499, np.nextafter(F32(0.499), F32(1)), np.nextafter(F32(0.499), F32(0)), -0.0, 0.5,
                  0.25, 0.75, 1.0], F32)
    TT, SS = np.meshgrid(T, S, indexing="ij")
    TT, SS = TT.ravel().astype(F32), SS.ravel().astype(F32)
    out = {}
    for nsub in (0, 37):
        S1, S2, k, rc = cf.computeStateProperties(wrap(TT), wrap(SS), P, nsub)
        out.update({f"S1_{nsub}": np.asarray(S1, F32), f"S2_{nsub}": np.asarray(S2).astype(bool),
                    f"k_{nsub}": np.asarray(k, F32), f"rhocp_{nsub}": np.asarray(rc, F32)})
    return out


def surface_flux_hot_and_capped(cf, wrap):
    """computeConvRadBC cF:2207-2301 on a top surface that spans ambient .. beyond the T_boiling + 1000 cap (the
    evaporation term and its min(T, T_b + 1000) clamp), on the scenario's Level 3."""
    Levels, P = _level(cf)
    L = Levels[3]
    nn, ne = int(L["nn"]), int(L["ne"])
    nx, ny, nz = [int(v) for v in L["nodes"]]
    rng = np.random.default_rng(7)
    T = (300.0 + 200.0 * rng.random(nn)).astype(F32).reshape(nz, ny, nx)
    T[-1] = np.linspace(298.15, 5200.0, nx * ny, dtype=F32).reshape(ny, nx)
    F = cf.computeConvRadBC(L, wrap(T.ravel()), ne, nn, P, wrap(np.zeros(nn, F32)))
    return {"F": np.asarray(F, F32)}


def interpolation_outside_the_parent(cf, wrap):
    """interpolatePoints cF:1131-1210 at target grids that stick out of the parent on every side (the +-1e-2 validity
    window zeroes the weights there), that touch its faces exactly, and that sit a hair inside / outside."""
    Levels, P = _level(cf)
    L = Levels[1]
    x, y, z = [np.asarray(c, F32) for c in L["node_coords"]]
    nn = int(L["nn"])
    u = (300.0 + np.arange(nn, dtype=F32) * F32(0.37)).astype(F32)
    hx, hy, hz = [F32(v) for v in L["h"]]
    eps = F32(1e-6)
    xs = np.array([x[0] - hx, x[0] - F32(0.011) * hx, x[0] - F32(0.009) * hx, x[0], x[0] + eps, 0.5 * (x[0] + x[1]),
                   x[-1] - eps, x[-1], x[-1] + F32(0.009) * hx, x[-1] + F32(0.011) * hx, x[-1] + hx], F32)
    ys = np.array([y[0] - hy, y[0], 0.25 * (y[0] + 3 * y[1]), y[-1], y[-1] + F32(0.5) * hy], F32)
    zs = np.array([z[0] - F32(0.5) * hz, z[0], 0.5 * (z[-2] + z[-1]), z[-1], z[-1] + F32(0.02) * hz], F32)
    out = cf.interpolatePoints(L, wrap(u), [wrap(xs), wrap(ys), wrap(zs)])
    return {"u_new": np.asarray(out, F32)}


CASES = {"state_properties_at_thresholds": state_properties_at_thresholds,
         "surface_flux_hot_and_capped": surface_flux_hot_and_capped,
         "interpolation_outside_the_parent": interpolation_outside_the_parent}


def run(cf, wrap=lambda a: a):
    flat = {}
    for name, fn in CASES.items():
        for k, v in fn(cf, wrap).items():
            flat[f"{name}/{k}"] = v
    return flat
