"""Test infrastructure for gomelt_b200/driver.py: a NumPy array adapter (drives the oracle through the
same loop) and a recording stub namespace (exercises the control flow without arithmetic)."""
import copy  # noqa: F401
import os

import numpy as np

F32 = np.float32


class NumpyArrays:
    def zeros(self, n):
        return np.zeros(int(n), F32)

    def f32(self, x):
        return np.asarray(x, dtype=F32)

    def maximum(self, a, b):
        return np.maximum(a, F32(b) if np.ndim(b) == 0 else b).astype(F32)

    def take(self, a, idx):
        return np.asarray(a)[np.asarray(idx)]

    def put(self, a, idx, v):
        a = np.array(a, copy=True)
        a[np.asarray(idx)] = v
        return a

    def add_at(self, a, idx, v):
        a = np.array(a, copy=True)
        a[np.asarray(idx)] = a[np.asarray(idx)] + v
        return a.astype(F32)

    def shift_down(self, a, n1, n2):
        out = np.zeros_like(a)
        out[:n2] = a[n1:]
        return out

    def fill_prefix(self, a, n, value):
        a = np.array(a, copy=True)
        a[: int(n)] = value
        return a

    def all_zero(self, a):
        return bool((np.asarray(a) == 0).all())

    def zeros_like(self, a):
        return np.zeros_like(a)

    def set_row(self, m, i, v):
        m = np.array(m, copy=True)
        m[i, :] = v
        return m

    def get_row(self, m, i):
        return np.array(m[i, :], copy=True)

    def positive(self, a):
        return (np.asarray(a) > 0).astype(F32)

    def host(self, a):
        return np.asarray(a)


class RecordingStub:
    """A ``computeFunctions``-shaped namespace whose steppers only record that they were called."""

    def __init__(self, real):
        self.real, self.calls = real, []
        for name in ("SetupProperties", "SetupNonmesh", "getStaticSubcycle", "count_lines", "parsingGcode"):
            setattr(self, name, getattr(real, name))

    def SetupLevels(self, inp, P):
        z = np.linspace(-4, 2, 31, dtype=F32)
        L = [{"nn": 8, "nodes": [2, 2, 2], "layer_idx_delta": 1, "S1": np.zeros(8, F32), "idx": np.arange(4),
              "node_coords": [z, z, z.copy()], "orig_node_coords": [z, z, z.copy()]}]
        for h in (0.2, 0.04, 0.02):
            L.append({"nn": 4, "h": [h] * 3, "T0": np.zeros(4, F32), "S1": np.zeros(4, F32), "Tprime0": np.ones(4, F32),
                      "S1_storage": np.zeros((5, 4), F32), "node_coords": [z, z, z.copy()],
                      "orig_node_coords": [z, z, z.copy()]})
        return L

    def getStaticNodesAndElements(self, L):
        return (1, 1, 4, 4, 4)

    def interpolatePointsMatrix(self, L, c):
        return None

    def interpolatePoints(self, L, u, c):
        return np.asarray(u)

    def calcStaticTmpNodesAndElements(self, L, lp):
        return (1, 4)

    def getSubstrateNodes(self, L):
        return (2, 2, 2, 2)

    def moveEverything(self, v, vs, L, mh, LI, r12, r23, lh):
        self.calls.append("move")
        return L, {}, LI, mh

    def stepGOMELT(self, L, *a):
        self.calls.append("step")
        return L, np.zeros(4, bool)

    def stepGOMELTDwellTime(self, L, *a):
        self.calls.append("dwell")
        return L

    def subcycleGOMELT(self, L, *a):
        self.calls.append("subcycle")
        return L, None, None, None, a[-2], a[-1]

    def melting_temp(self, T, dt, Tm, acc, idx):
        return acc


def small_two_layer_input(tmp):
    """Scaled-down three-level problem (tests/golden/scenario.py sizes) driven from G-code: two layers, a
    short track each, pauses long enough to reach the Level-1-only dwell mode."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import scenario

    inp = copy.deepcopy(scenario.SMALL_INPUT)
    g = os.path.join(tmp, "two_layers.gcode")
    with open(g, "w") as fh:
        fh.write("G0 X0.40 Y0.40 Z0.0\nG1 X0.50 Y0.40 Z0.0\nG0 X0.50 Y0.44 Z0.04\nG1 X0.42 Y0.44 Z0.04\n")
    inp["nonmesh"].update(save_path=tmp + "/", toolpath=os.path.join(tmp, "toolpath.txt"), gcode=g, use_txt=0,
                          wait_time=6, dwell_time=2e-4, dwell_time_multiplier=1, subcycle_num_L2=2,
                          subcycle_num_L3=2, record_step=4, info_T=0, laser_velocity=500)
    return inp


def serpentine_input(tmp):
    """BASELINE.json configs[2] in miniature: one layer, three serpentine tracks joined by rapid (G0) moves, so
    that jump rows and the faster-than-100x-velocity single-step trigger (gm:168-171) are exercised."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import scenario

    inp = copy.deepcopy(scenario.SMALL_INPUT)
    g = os.path.join(tmp, "serpentine.gcode")
    with open(g, "w") as fh:
        fh.write("G0 X0.40 Y0.36 Z0.0\nG1 X0.48 Y0.36 Z0.0\nG0 X0.48 Y0.40 Z0.0\nG1 X0.40 Y0.40 Z0.0\n"
                 "G0 X0.40 Y0.44 Z0.0\nG1 X0.48 Y0.44 Z0.0\n")
    inp["nonmesh"].update(save_path=tmp + "/", toolpath=os.path.join(tmp, "toolpath.txt"), gcode=g, use_txt=0,
                          wait_time=4, dwell_time=8e-5, dwell_time_multiplier=1, subcycle_num_L2=2,
                          subcycle_num_L3=2, record_step=4, info_T=0, laser_velocity=500)
    return inp


def wide_part_input(tmp):
    """examples/example.json with a part-scale level wide enough for the fast level-step kernel (72 x 20 x 30 elements:
    73 nodes per row) - the Level-1 sweeps of the slab-decomposed runs then go through level_step_v3 with the fused
    halo exchange - driven by a short two-layer G-code with a pause that reaches the dwell mode (the bench's N>1 line
    runs the same case: bench_tools/dist_check.py)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench_tools.dist_check import wide_part_input as make

    return make(tmp)
