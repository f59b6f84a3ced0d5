"""Host-side level geometry of the product (gomelt_b200/levels.py, JAX-free) against the oracle's SetupLevels
(cF:115-264, itself pinned to the reference's own source by tests/test_oracle_golden.py): coordinates, overlap
index sets, Level-0 scatter indices, travel limits, static sizes - on examples/example.json and on the scaled-down
scenario of tests/golden/scenario.py.  CPU only: no field is allocated."""
import copy
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import setup as osetup  # noqa: E402


def _inputs():
    ex = json.load(open(os.path.join(ROOT, "examples", "example.json")))
    yield "example", ex
    import scenario

    yield "scenario", copy.deepcopy(scenario.SMALL_INPUT)


@pytest.mark.parametrize("name,inp", list(_inputs()))
def test_geometry_matches_oracle(name, inp):
    lev = importlib.import_module("gomelt_b200.levels")
    sc = importlib.import_module("gomelt_b200.schema")
    P_o = osetup.SetupProperties(copy.deepcopy(inp["properties"]))
    P_p = sc.SetupProperties(copy.deepcopy(inp["properties"]))
    assert set(P_o) == set(P_p)
    for k in P_o:
        assert np.array_equal(np.asarray(P_o[k], np.float32), np.asarray(P_p[k], np.float32)), k
    ref = osetup.SetupLevels(copy.deepcopy(inp), P_o)
    got = lev.build_levels(copy.deepcopy(inp), P_p)
    for i in range(4):
        R, G = ref[i], got[i]
        assert [int(v) for v in R["nodes"]] == [int(v) for v in G["nodes"]], i
        assert int(R["nn"]) == int(G["nn"]) and int(R["ne"]) == int(G["ne"]), i
        for d in range(3):
            assert np.array_equal(np.asarray(R["node_coords"][d], np.float32), np.asarray(G["node_coords"][d], np.float32)), (i, d)
    for i in (1, 2, 3):
        assert np.array_equal(np.asarray(ref[i]["h"], np.float32), np.asarray(got[i]["h"], np.float32)), i
    for i in (2, 3):
        for a in ("ix", "iy", "iz"):
            assert np.allclose(np.asarray(ref[i]["bounds"][a], np.float32), np.asarray(got[i]["bounds"][a], np.float32),
                               rtol=0, atol=1e-6), (i, a)
        for d in range(3):
            assert np.array_equal(np.asarray(ref[i]["orig_overlap_nodes"][d]), np.asarray(got[i]["orig_overlap_nodes"][d])), (i, d)
            assert np.array_equal(np.asarray(ref[i]["orig_overlap_coors"][d], np.float32),
                                  np.asarray(got[i]["orig_overlap_coors"][d], np.float32)), (i, d)
    for key in ("idx", "idx_L2"):
        assert np.array_equal(np.asarray(ref[0][key]).astype(np.int64), np.asarray(got[0][key]).astype(np.int64)), key
    assert int(ref[0]["layer_idx_delta"]) == int(got[0]["layer_idx_delta"])
    for d in range(3):
        assert np.array_equal(np.asarray(ref[0]["orig_overlap_nodes_L2"][d]), np.asarray(got[0]["orig_overlap_nodes_L2"][d])), d
    # static sizes the steppers take as (static) arguments
    assert tuple(int(v) for v in osetup.getStaticNodesAndElements(ref)) == lev.static_sizes(got)
    assert tuple(int(v) for v in osetup.getSubstrateNodes(ref)) == lev.substrate_counts(got)
    z1 = np.asarray(got[1]["node_coords"][2])
    for zl in (float(z1[len(z1) // 2]), float(z1[-1]), float(z1[1]) + 1e-6):
        v = np.array([0.0, 0.0, zl], np.float32)
        assert tuple(int(q) for q in osetup.calcStaticTmpNodesAndElements(ref, v)) == lev.active_sizes(got, v), zl
