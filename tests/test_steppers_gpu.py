"""GPU parity of the drop-in step orchestrators (stepGOMELT, subcycleGOMELT, stepGOMELTDwellTime,
moveEverything) against (a) the golden vectors produced by the reference's own source and (b) the oracle,
on the shared scenario of tests/golden/scenario.py.

Tolerances (BASELINE.json north_star): temperatures within 1e-5 relative max error in float32; state /
index fields and the melt-pool extent (S2) bit-exact.
"""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import scenario  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL_T = 1e-5


@pytest.fixture(scope="module")
def product_run(tmp_path_factory):
    cf = importlib.import_module("gomelt_b200.computeFunctions")
    return scenario.run(cf, save_path=str(tmp_path_factory.mktemp("prod")) + "/")


def _check(out, ref_get, keys):
    worst = {}
    for phase, name in keys:
        a, b = np.asarray(out[phase][name]), ref_get(phase, name)
        assert a.shape == b.shape, (phase, name, a.shape, b.shape)
        if b.dtype.kind in "iub":
            assert np.array_equal(a.astype(b.dtype), b), f"{phase}/{name}: {int(np.sum(a.astype(b.dtype) != b))} differ"
        else:
            den = np.maximum(np.abs(b), 1.0) if name.endswith("_T0") else max(float(np.abs(b).max()), 1e-30)
            err = float(np.max(np.abs(a.astype(np.float64) - b) / den))
            worst[f"{phase}/{name}"] = err
            tol = RTOL_T if name.endswith("_T0") else 2e-5   # T' fields: relative to their max
            assert err <= tol, f"{phase}/{name}: max rel err {err:.3e}"
    return worst


def test_against_reference_golden(product_run):
    ref = np.load(os.path.join(HERE, "golden", "small_run_reference.npz"))
    keys = [tuple(k.split("/")) for k in ref.files if not k.startswith("meta/")]
    worst = _check(product_run, lambda p, n: ref[f"{p}/{n}"], keys)
    print("worst float errors vs reference source:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])


def test_against_oracle(product_run, tmp_path):
    from oracle import computeFunctions as cF

    ora = scenario.run(cF, save_path=str(tmp_path) + "/")
    keys = [(p, n) for p in ora if p != "meta" for n in ora[p]]
    _check(product_run, lambda p, n: np.asarray(ora[p][n]), keys)


def test_melt_pool_extent_bit_exact(product_run):
    ref = np.load(os.path.join(HERE, "golden", "small_run_reference.npz"))
    for phase in ("step0", "step1", "step2", "subcycle"):
        assert np.array_equal(product_run[phase]["L3_S2"], ref[f"{phase}/L3_S2"])
        assert np.array_equal(product_run[phase]["L0_S2"], ref[f"{phase}/L0_S2"])
    assert product_run["subcycle"]["L3_S2"].sum() > 0
