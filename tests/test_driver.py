"""gomelt_b200/driver.py: (CPU) the loop's control flow on the reference's own example toolpath, and (GPU)
the whole drop-in run against the oracle driven through the same loop."""
import importlib
import json
import os

import numpy as np
import pytest

from driver_support import NumpyArrays, RecordingStub, serpentine_input, small_two_layer_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mode_sequence_on_the_example_toolpath(tmp_path):
    """the default example is 25 single-step rows (layer start), 30 subcycle blocks, 25
    single-step rows (wait counter near wait_time), 499 Level-1 dwell rows; moveEverything runs on every
    single-step / dwell row and once per subcycle block (gm:292, 422)."""
    drv = importlib.import_module("gomelt_b200.driver")
    sc = importlib.import_module("gomelt_b200.schema")
    tp = importlib.import_module("gomelt_b200.toolpath")

    class Real:
        SetupProperties, SetupNonmesh, getStaticSubcycle = sc.SetupProperties, sc.SetupNonmesh, sc.getStaticSubcycle
        count_lines, parsingGcode = tp.count_lines, tp.parsingGcode

    inp = json.load(open(os.path.join(ROOT, "examples", "example.json")))
    inp["nonmesh"].update(save_path=str(tmp_path) + "/", toolpath=str(tmp_path / "toolpath.txt"),
                          gcode=os.path.join(ROOT, "examples", "gcodefiles", "example.gcode"))
    stub = RecordingStub(Real)
    out = drv.go_melt(inp, cf=stub, xp=NumpyArrays(), write_final=False)
    assert out["total_t_inc"] == 1299 and out["time_inc"] == 1299
    assert abs(out["sim_seconds"] - 1.006) < 1e-4
    c = out["counts"]
    assert (c["stepGOMELT"], c["subcycleGOMELT"], c["stepGOMELTDwellTime"], c["moveEverything"]) == (50, 30, 499, 579)
    calls = [x for x in stub.calls if x != "move"]
    assert calls[:25] == ["step"] * 25 and calls[25:55] == ["subcycle"] * 30
    assert calls[55:80] == ["step"] * 25 and calls[80:] == ["dwell"] * 499
    # the same schedule computed ahead of the run from the toolpath file alone (N2: host-side block scheduler)
    plan = drv.plan_toolpath(out["Nonmesh"]["toolpath"], out["Nonmesh"], sc.getStaticSubcycle(out["Nonmesh"]))
    assert sum(b["rows"] for b in plan) == 1299
    assert [b["mode"] for b in plan[:32]] == ["single"] + ["subcycle"] * 30 + ["single"]
    assert sum(b["steps"] for b in plan) == 50 and sum(b["dwells"] for b in plan) == 499
    assert sum(b["layer_changes"] for b in plan) == 1
    # the pause is runs of identical rows (one per 25-row block): what dwellRows replays as CUDA graphs
    assert sum(sum(b["dwell_runs"]) for b in plan) >= 490


@pytest.mark.gpu
def test_graph_replay_of_dwell_runs_is_bit_identical(tmp_path):
    """N2: runs of identical dwell rows go through computeFunctions.dwellRows (two eager rows, two captured as one
    CUDA graph, the rest replayed); the run must equal the row-by-row run bit for bit on every level."""
    import torch

    import gomelt_b200 as gm

    drv = importlib.import_module("gomelt_b200.driver")
    inp = json.load(open(os.path.join(ROOT, "examples", "example.json")))

    def run(sub, graphs):
        d = tmp_path / sub
        d.mkdir()
        i = json.loads(json.dumps(inp))
        i["nonmesh"].update(save_path=str(d) + "/", toolpath=str(d / "toolpath.txt"), output_files=0, info_T=0,
                            gcode=os.path.join(ROOT, "examples", "gcodefiles", "example.gcode"))
        g0 = gm.ops.GRAPH_LAUNCHES
        res = drv.go_melt(i, write_final=False, graphs=graphs)
        torch.cuda.synchronize()
        return res, gm.ops.GRAPH_LAUNCHES - g0

    a, ga = run("graphs", True)
    b, gb = run("eager", False)
    assert a["counts"] == b["counts"] and a["time_inc"] == b["time_inc"] == 1299
    assert abs(a["dwell_seconds"] - b["dwell_seconds"]) < 1e-12
    assert gb == 0 and ga > 2000, (ga, gb)   # most of the 499 x ~6 launches of the pause ran as replays
    for lvl in (1, 2, 3):
        for f in ("T0", "S1"):
            assert torch.equal(a["Levels"][lvl][f], b["Levels"][lvl][f]), (lvl, f)
    for lvl in (2, 3):
        assert torch.equal(a["Levels"][lvl]["Tprime0"], b["Levels"][lvl]["Tprime0"])
    assert torch.equal(a["accum_time"], b["accum_time"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["two_layers", "serpentine"])
def test_whole_run_matches_oracle(tmp_path, case):
    """G-code through go_melt(): product (CUDA) vs oracle (NumPy), same driver loop - two layers with dwell
    (BASELINE.json configs[3] in miniature) and a one-layer serpentine scan with rapid moves (configs[2]).
    Temperatures within 1e-5 relative, states / melt flags / melt-time bookkeeping exact or to rounding."""
    from oracle import computeFunctions as cF

    drv = importlib.import_module("gomelt_b200.driver")
    make = small_two_layer_input if case == "two_layers" else serpentine_input
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    got = drv.go_melt(make(str(tmp_path / "gpu")), write_final=True)
    ref = drv.go_melt(make(str(tmp_path / "cpu")), cf=cF, xp=NumpyArrays(), write_final=False)
    assert got["counts"] == ref["counts"] and got["time_inc"] == ref["time_inc"]
    c = got["counts"]
    # the schedule computed ahead of the run from the toolpath alone (driver.plan_toolpath) is the one that was executed
    sc = importlib.import_module("gomelt_b200.schema")
    plan = drv.plan_toolpath(got["Nonmesh"]["toolpath"], got["Nonmesh"], sc.getStaticSubcycle(got["Nonmesh"]))
    assert sum(b["rows"] for b in plan) == got["time_inc"]
    assert sum(b["steps"] for b in plan) == c["stepGOMELT"] and sum(b["dwells"] for b in plan) == c["stepGOMELTDwellTime"]
    assert sum(b["mode"] == "subcycle" for b in plan) == c["subcycleGOMELT"]
    assert sum(b["layer_changes"] for b in plan) == c["layers"]
    assert c["stepGOMELT"] > 0 and c["subcycleGOMELT"] > 0 and c["stepGOMELTDwellTime"] > 0
    assert c["layers"] == (2 if case == "two_layers" else 1)
    host = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    for lvl in (1, 2, 3):
        a, b = host(got["Levels"][lvl]["T0"]), host(ref["Levels"][lvl]["T0"])
        err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
        assert err <= 1e-5, (lvl, err)
        assert np.array_equal(host(got["Levels"][lvl]["S1"]), host(ref["Levels"][lvl]["S1"])), lvl
    assert np.array_equal(host(got["Levels"][3]["S2"]).astype(bool), host(ref["Levels"][3]["S2"]).astype(bool))
    assert np.array_equal(host(got["Levels"][0]["S1"]), host(ref["Levels"][0]["S1"]))
    acc_g, acc_r = host(got["accum_time"]), host(ref["accum_time"])
    assert acc_r.max() > 0  # the run melts
    assert np.allclose(acc_g, acc_r, rtol=1e-5, atol=1e-9)
    final = np.load(tmp_path / "gpu" / "FinalTemperatureFields.npz")
    assert np.array_equal(final["L3T"], host(got["Levels"][3]["T0"]))
    # ... and directly against what the REFERENCE'S OWN driver + computeFunctions wrote for the same G-code run
    # (tests/golden/*_reference_driver.npz, see test_driver_loop_reproduces_the_reference_driver): north-star tolerance
    gold = np.load(os.path.join(ROOT, "tests", "golden",
                                ("two_layer" if case == "two_layers" else "serpentine") + "_reference_driver.npz"))
    for lvl in (1, 2, 3):
        a, b = host(got["Levels"][lvl]["T0"]), np.asarray(gold[f"L{lvl}T"], np.float32)
        err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
        assert err <= 1.2e-5, (lvl, err)
    assert np.allclose(acc_g, np.asarray(gold["accum_time"], np.float32), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("case", ["two_layers", "serpentine"])
def test_driver_loop_reproduces_the_reference_driver(tmp_path, case):
    """gomelt_b200/driver.py (the loop) + oracle/ (the arithmetic) against the REFERENCE'S OWN driver: the goldens
    tests/golden/{two_layer,serpentine}_reference_driver.npz hold the FinalTemperatureFields / accum_time that the
    unmodified go_melt.go_melt (gm:16-530) wrote on top of the unmodified computeFunctions.py, both executed through
    the NumPy jax shim (tests/golden/make_golden.py --driver <case>): two layers with single steps, subcycle blocks,
    dwell rows and a layer change; one layer of serpentine tracks with rapid moves (jump rows, the 100x-velocity
    single-step trigger).  CPU only."""
    from oracle import computeFunctions as cF

    name = {"two_layers": "two_layer", "serpentine": "serpentine"}[case]
    ref = np.load(os.path.join(ROOT, "tests", "golden", name + "_reference_driver.npz"))
    make = small_two_layer_input if case == "two_layers" else serpentine_input
    drv = importlib.import_module("gomelt_b200.driver")
    got = drv.go_melt(make(str(tmp_path)), cf=cF, xp=NumpyArrays(), write_final=False)
    c = got["counts"]
    assert c["layers"] == (2 if case == "two_layers" else 1)
    assert c["stepGOMELT"] > 0 and c["subcycleGOMELT"] > 0 and c["stepGOMELTDwellTime"] > 0
    for lvl in (1, 2, 3):
        a, b = np.asarray(got["Levels"][lvl]["T0"], np.float32), np.asarray(ref[f"L{lvl}T"], np.float32)
        assert a.shape == b.shape, lvl
        err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
        assert err <= 5e-6, (lvl, err)      # measured on two_layers: 1.1e-6 / 8.7e-7 / 3.8e-7 on Levels 1 / 2 / 3
        assert np.array_equal(a >= 1609.0, b >= 1609.0), lvl  # the molten set (T >= T_liquidus), bit-exact
    a, b = np.asarray(got["accum_time"], np.float32), np.asarray(ref["accum_time"], np.float32)
    assert a.shape == b.shape and b.max() > 0
    assert np.allclose(a, b, rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_example_json_first_rows_match_oracle(tmp_path):
    """BASELINE.json configs[0]: examples/example.json verbatim (Level 1 50x20x30, Levels 2/3 100x100x10 elements,
    N2 = N3 = 5) through go_melt() on the GPU against the oracle through the same loop, on the first 75 toolpath rows
    = the 25 layer-start single steps (window shifts included) + two 25-row subcycle blocks.  The oracle needs about
    a minute of one host core for them.  Temperatures within 1e-5 relative on every level, the melt pool (S2, T >=
    T_liquidus), the states and the Level-0 index sets identical."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
    import run_example
    from oracle import computeFunctions as cF

    drv = importlib.import_module("gomelt_b200.driver")
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    got = drv.go_melt(run_example.load_input(str(tmp_path / "gpu"), 75), write_final=False)
    ref = drv.go_melt(run_example.load_input(str(tmp_path / "cpu"), 75), cf=cF, xp=NumpyArrays(), write_final=False)
    assert got["counts"] == ref["counts"] and got["time_inc"] == ref["time_inc"] == 75
    assert (got["counts"]["stepGOMELT"], got["counts"]["subcycleGOMELT"]) == (25, 2)
    host = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    for lvl in (1, 2, 3):
        a, b = host(got["Levels"][lvl]["T0"]), host(ref["Levels"][lvl]["T0"])
        err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
        assert err <= 1e-5, (lvl, err)
        assert np.array_equal(host(got["Levels"][lvl]["S1"]), host(ref["Levels"][lvl]["S1"])), lvl
    s2g, s2r = host(got["Levels"][3]["S2"]).astype(bool), host(ref["Levels"][3]["S2"]).astype(bool)
    assert s2r.sum() > 100 and np.array_equal(s2g, s2r)          # the melt pool exists and is identical
    assert np.array_equal(host(got["Levels"][0]["S2"]).astype(bool), host(ref["Levels"][0]["S2"]).astype(bool))
    assert np.array_equal(host(got["Levels"][0]["idx"]), host(ref["Levels"][0]["idx"]))
    assert np.allclose(host(got["accum_time"]), host(ref["accum_time"]), rtol=1e-5, atol=1e-9)


def _restart_vs_uninterrupted(tmp_path, cf, xp):
    """gm:108-129 + 390-411: a two-layer run stopped at its first checkpoint (restart_layer_num = 1: the driver
    returns right after writing Checkpoint0001, no final output) and restarted with layer_num = 1 must end exactly
    where the uninterrupted run ends; the per-layer melt-time file accum_time0000.npz (gm:253-261) is written at the
    layer change in both."""
    drv = importlib.import_module("gomelt_b200.driver")
    out = importlib.import_module("gomelt_b200.output")
    hooks = {k: v for k, v in out.driver_hooks().items() if k in ("on_checkpoint", "load_checkpoint", "on_layer_accum")}
    kw = dict(hooks=hooks, write_final=True)
    if cf is not None:
        kw.update(cf=cf, xp=xp)
    for d in ("whole", "split"):
        (tmp_path / d).mkdir()
    whole = drv.go_melt(small_two_layer_input(str(tmp_path / "whole")), **kw)
    assert whole["counts"]["layers"] == 2 and not whole["stopped_at_layer_check"]
    assert os.path.exists(tmp_path / "whole" / "checkpoint" / "Checkpoint0001" / "header.json")
    assert os.path.exists(tmp_path / "whole" / "accum_time0000.npz") and os.path.exists(tmp_path / "whole" / "accum_time0001.npz")
    inp = small_two_layer_input(str(tmp_path / "split"))
    inp["nonmesh"]["restart_layer_num"] = 1
    first = drv.go_melt(inp, **kw)
    assert first["stopped_at_layer_check"] and first["time_inc"] < whole["time_inc"]
    assert not os.path.exists(tmp_path / "split" / "FinalTemperatureFields.npz")   # early return: no final files
    inp = small_two_layer_input(str(tmp_path / "split"))
    inp["nonmesh"].update(layer_num=1, use_txt=1)
    second = drv.go_melt(inp, **kw)
    assert second["time_inc"] == whole["time_inc"]
    host = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    for lvl in (1, 2, 3):
        for f in ("T0", "S1"):
            assert np.array_equal(host(second["Levels"][lvl][f]), host(whole["Levels"][lvl][f])), (lvl, f)
    assert np.array_equal(host(second["Levels"][3]["S2"]), host(whole["Levels"][3]["S2"]))
    assert np.array_equal(host(second["Levels"][0]["S1"]), host(whole["Levels"][0]["S1"]))
    assert host(whole["accum_time"]).max() > 0 and np.array_equal(host(second["accum_time"]), host(whole["accum_time"]))
    a = np.load(tmp_path / "whole" / "accum_time0000.npz")["accum_time"]
    b = np.load(tmp_path / "split" / "accum_time0000.npz")["accum_time"]
    assert a.max() > 0 and np.array_equal(a, b)
    # a restart without the checkpoint hook is an error, not a silent run from t = 0
    with pytest.raises(RuntimeError, match="load_checkpoint"):
        drv.go_melt(inp, **{**kw, "hooks": {}})


def test_restart_from_checkpoint_equals_uninterrupted_run_oracle(tmp_path):
    from oracle import computeFunctions as cF

    _restart_vs_uninterrupted(tmp_path, cF, NumpyArrays())


@pytest.mark.gpu
def test_restart_from_checkpoint_equals_uninterrupted_run_gpu(tmp_path):
    _restart_vs_uninterrupted(tmp_path, None, None)
