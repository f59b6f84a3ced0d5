"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCE.

``tests/golden/small_run_reference.npz`` was written by ``tests/golden/make_golden.py``: the
unmodified ``/root/reference/go_melt/computeFunctions.py`` executed through a NumPy stand-in for the
``jax`` API (``tests/golden/jax_numpy_shim.py``) on the scaled-down three-level run of
``tests/golden/scenario.py`` (layer start, 3 single steps with window shifts, one 2x2 subcycle block,
2 dwell steps).  This pins the oracle: every integer / state / index field bit-exact, every float
field within 5e-6 relative (the shim does not reproduce XLA's summation order, see its header).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import scenario  # noqa: E402
from oracle import computeFunctions as cF  # noqa: E402

FLOAT_RTOL = 5e-6


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    out = scenario.run(cF, save_path=str(tmp_path_factory.mktemp("golden")) + "/")
    ref = np.load(os.path.join(HERE, "golden", "small_run_reference.npz"))
    return out, ref


def test_golden_file_covers_every_phase(runs):
    _, ref = runs
    phases = {k.split("/")[0] for k in ref.files}
    assert phases == {"step0", "step1", "step2", "subcycle", "dwell7", "dwell8", "meta"}
    assert len(ref.files) >= 100


def test_oracle_matches_reference_source(runs):
    out, ref = runs
    worst = 0.0
    for key in ref.files:
        phase, name = key.split("/")
        a, b = np.asarray(out[phase][name]), ref[key]
        assert a.shape == b.shape, key
        if b.dtype.kind in "iub":
            assert np.array_equal(a.astype(b.dtype), b), f"{key}: {int(np.sum(a != b))} entries differ"
        else:
            den = np.maximum(np.abs(b), 1.0) if name.endswith("_T0") else max(float(np.abs(b).max()), 1e-30)
            err = float(np.max(np.abs(a - b) / den))
            worst = max(worst, err)
            assert err <= FLOAT_RTOL, f"{key}: max rel err {err:.3e}"
    assert worst > 0.0  # the run is not trivially identical (different summation orders)


def test_melt_pool_extent_is_bit_exact(runs):
    """north_star: the set of nodes above liquidus must be identical."""
    out, ref = runs
    for phase in ("step0", "step1", "step2", "subcycle"):
        assert np.array_equal(out[phase]["L3_S2"], ref[f"{phase}/L3_S2"])
        assert np.array_equal(out[phase]["L0_S2"], ref[f"{phase}/L0_S2"])
    assert ref["subcycle/L3_S2"].sum() > 0  # the scenario does melt
    assert ref["subcycle/accum"].max() > 0


def test_windows_moved(runs):
    _, ref = runs
    assert ref["step2/L3_x"][0] > ref["step0/L3_x"][0]  # the Level-3 window followed the laser
    assert ref["subcycle/move_hist"].shape == (3,)


def test_oracle_matches_reference_on_edge_cases():
    """Per-function pin on crafted inputs (tests/golden/edge_cases.py; the reference's own functions, run through the
    shim by make_golden.py --edge, wrote edge_cases_reference.npz): computeStateProperties exactly at / one ulp around
    the solidus, the liquidus and the 0.499 state threshold (incl. -0.0 and a substrate prefix) - bit-exact;
    computeConvRadBC on a surface running from ambient past the T_boiling + 1000 cap; interpolatePoints at targets
    outside, on and a hair inside / outside the parent's faces (the +-1e-2 validity window)."""
    import edge_cases

    ref = np.load(os.path.join(HERE, "golden", "edge_cases_reference.npz"))
    got = edge_cases.run(cF)
    assert set(got) == set(ref.files) and len(got) == 10
    for k in ref.files:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        assert a.shape == b.shape, k
        if k.startswith("state_properties_at_thresholds/"):
            assert np.array_equal(a, b), k                      # selects of constants / one FMA: no rounding freedom
        else:
            scale = float(np.max(np.abs(b)))
            assert float(np.max(np.abs(a.astype(np.float64) - b))) <= FLOAT_RTOL * scale, k
            assert np.array_equal(a == 0, b == 0), k            # same support (zeroed weights outside the parent)
    assert (ref["interpolation_outside_the_parent/u_new"] == 0).sum() > 100   # the case does reach outside
