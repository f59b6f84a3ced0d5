"""Generate tests/golden/*.npz by executing the reference's own, unmodified source
(/root/reference/go_melt/computeFunctions.py) through the NumPy ``jax`` shim.

    python tests/golden/make_golden.py            (needs /root/reference; run in the build container)

The fixtures travel with the repo; the reference and the shim are not needed to run the tests.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jax_numpy_shim as shim  # noqa: E402
import scenario  # noqa: E402


def main():
    cF = shim.load_reference()
    t0 = time.time()
    out = scenario.run(cF, wrap=shim._wrap)
    flat = {}
    for phase, d in out.items():
        for k, v in d.items():
            flat[f"{phase}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "small_run_reference.npz"), **flat)
    print(f"small_run_reference.npz: {len(flat)} arrays, {time.time() - t0:.1f} s")
    toolpath_golden(cF)
    toolpath_golden_serpentine(cF)
    toolpath_golden_dwell(cF)


def driver_golden(case="two_layers"):
    """The reference's own DRIVER (go_melt.go_melt gm:16-530, unmodified) on top of its own computeFunctions, both
    through the NumPy ``jax`` shim, on a G-code run of tests/driver_support.py -> <case>_reference_driver.npz = its
    FinalTemperatureFields + the Level-0 melt-time field(s).
      two_layers: small_two_layer_input, 57 toolpath rows - layer start, single steps, subcycle blocks, dwell, a
                  layer change (28 min here);
      serpentine: serpentine_input, 70 rows - one layer, three tracks joined by rapid (G0) moves: jump rows and the
                  faster-than-100x-velocity single-step trigger gm:168-171 (~35 min)."""
    import importlib
    import tempfile

    sys.path.insert(0, os.path.join(HERE, ".."))
    import driver_support

    make = {"two_layers": driver_support.small_two_layer_input, "serpentine": driver_support.serpentine_input}[case]
    shim.load_reference()
    gm = importlib.import_module("go_melt")
    tmp = tempfile.mkdtemp()
    t0 = time.time()
    gm.go_melt(make(tmp))
    out = collect_driver_outputs(tmp)
    np.savez_compressed(os.path.join(HERE, {"two_layers": "two_layer", "serpentine": "serpentine"}[case]
                                     + "_reference_driver.npz"), **out)
    print(f"{case}: {sorted(out)} in {time.time() - t0:.0f} s")


def collect_driver_outputs(tmp):
    """FinalTemperatureFields + accum_time files the reference driver left in its save_path."""
    out = dict(np.load(os.path.join(tmp, "FinalTemperatureFields.npz")))
    # melt-time field of Level 0: written at every layer change and at the end of the run (the last one = final)
    acc = sorted(f for f in os.listdir(tmp) if f.startswith("accum_time") and f.endswith(".npz"))
    for f in acc[:-1]:
        out["accum_time_layer" + str(int(f[len("accum_time"):-4]))] = np.load(os.path.join(tmp, f))["accum_time"]
    out["accum_time"] = np.load(os.path.join(tmp, acc[-1]))["accum_time"]
    return out


def toolpath_golden_serpentine(cF):
    """The reference's parser on a two-layer serpentine scan with rapid (G0) moves between the tracks, a layer change
    and dwell rows (scenario.SERPENTINE_GCODE / SERPENTINE_NONMESH): BASELINE.json configs[2] / [3] in miniature."""
    import importlib
    import shutil
    import tempfile

    cP = importlib.import_module("createPath")
    tmp = tempfile.mkdtemp() + "/"
    with open(tmp + "serp.gcode", "w") as fh:
        fh.write(scenario.SERPENTINE_GCODE)
    nm = dict(scenario.SERPENTINE_NONMESH, save_path=tmp, toolpath=tmp + "toolpath.txt", gcode=tmp + "serp.gcode")
    n = cP.parsingGcode(cF.SetupNonmesh(nm), {"laser_power": 285.0}, [0.04, 0.04, 0.04])
    shutil.copy(tmp + "toolpath.txt", os.path.join(HERE, "toolpath_serpentine.txt"))
    print(f"toolpath_serpentine.txt: {n} rows")


DWELL_CASES = [  # round-number dwell configurations: int(dwell / dt / coef) vs int(dwell / (dt * coef)) differ in some
    {"dwell_time": 0.1, "timestep_L3": 2e-5, "wait_time": 1000, "subcycle_num_L2": 5, "subcycle_num_L3": 5,
     "dwell_time_multiplier": 1},
    {"dwell_time": 0.3, "timestep_L3": 1e-5, "wait_time": 100, "subcycle_num_L2": 3, "subcycle_num_L3": 4,
     "dwell_time_multiplier": 5},
    {"dwell_time": 0.7, "timestep_L3": 1e-5, "wait_time": 500, "subcycle_num_L2": 5, "subcycle_num_L3": 5,
     "dwell_time_multiplier": 8},
    {"dwell_time": 0.06, "timestep_L3": 3e-5, "wait_time": 50, "subcycle_num_L2": 2, "subcycle_num_L3": 2,
     "dwell_time_multiplier": 10},
    {"dwell_time": 0.002, "timestep_L3": 1e-5, "wait_time": 500, "subcycle_num_L2": 5, "subcycle_num_L3": 5,
     "dwell_time_multiplier": 8},
    {"dwell_time": 1.2, "timestep_L3": 2e-5, "wait_time": 200, "subcycle_num_L2": 4, "subcycle_num_L3": 6,
     "dwell_time_multiplier": 3},
    {"dwell_time": 0.1, "timestep_L3": 1e-5, "wait_time": 100, "subcycle_num_L2": 3, "subcycle_num_L3": 4,
     "dwell_time_multiplier": 1},
    {"dwell_time": 0.1, "timestep_L3": 2e-5, "wait_time": 200, "subcycle_num_L2": 3, "subcycle_num_L3": 4,
     "dwell_time_multiplier": 8},
]


def toolpath_golden_dwell(cF):
    """Row count and text digest of the reference parser's output on the two-layer serpentine G-code for DWELL_CASES
    -> toolpath_dwell_cases.json (the dwell row count depends on the float operation order of cP:82-84 / 173)."""
    import hashlib
    import importlib
    import json
    import tempfile

    cP = importlib.import_module("createPath")
    out = []
    for case in DWELL_CASES:
        tmp = tempfile.mkdtemp() + "/"
        with open(tmp + "serp.gcode", "w") as fh:
            fh.write(scenario.SERPENTINE_GCODE)
        nm = dict(scenario.SERPENTINE_NONMESH, save_path=tmp, toolpath=tmp + "toolpath.txt", gcode=tmp + "serp.gcode")
        nm.update(case)
        n = cP.parsingGcode(cF.SetupNonmesh(nm), {"laser_power": 285.0}, [0.04, 0.04, 0.04])
        text = open(tmp + "toolpath.txt", "rb").read()
        out.append({"nonmesh": case, "rows": int(n), "sha256": hashlib.sha256(text).hexdigest()})
    json.dump(out, open(os.path.join(HERE, "toolpath_dwell_cases.json"), "w"), indent=1)
    print("toolpath_dwell_cases.json:", [c["rows"] for c in out])


def toolpath_golden(cF):
    """The reference's own G-code parser (createPath.parsingGcode cP:6-188, unmodified) on its own example."""
    import importlib
    import json
    import shutil
    import tempfile

    cP = importlib.import_module("createPath")
    inp = json.load(open("/root/reference/examples/example.json"))
    tmp = tempfile.mkdtemp() + "/"
    nm = dict(inp["nonmesh"], save_path=tmp, toolpath=tmp + "toolpath.txt",
              gcode="/root/reference/examples/gcodefiles/example.gcode")
    n = cP.parsingGcode(cF.SetupNonmesh(nm), cF.SetupProperties(inp["properties"]), [0.04, 0.04, 0.04])
    shutil.copy(tmp + "toolpath.txt", os.path.join(HERE, "toolpath_example.txt"))
    print(f"toolpath_example.txt: {n} rows")


if __name__ == "__main__":
    if "--edge" in sys.argv:  # per-function edge cases (tests/golden/edge_cases.py), seconds
        import edge_cases

        np.savez_compressed(os.path.join(HERE, "edge_cases_reference.npz"),
                            **edge_cases.run(shim.load_reference(), wrap=shim._wrap))
    elif "--driver" in sys.argv:  # --driver [two_layers|serpentine]
        i = sys.argv.index("--driver")
        driver_golden(sys.argv[i + 1] if i + 1 < len(sys.argv) else "two_layers")
    elif "--toolpaths-only" in sys.argv:  # the .npz is left as committed
        _cF = shim.load_reference()
        toolpath_golden(_cF)
        toolpath_golden_serpentine(_cF)
        toolpath_golden_dwell(_cF)
    else:
        main()
