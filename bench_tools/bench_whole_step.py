"""Whole steps at BASELINE.json configs[1] scale (C2 of SURVEY.md 8d) through the drop-in entry points, not just K1:
a 10.29 M-node Level-3 window (500 x 500 x 40 elements, h = 0.02 mm) inside a same-count Level-2 window at h = 0.04 mm
inside a 120 x 120 x 30-element Level 1 (h = 0.2 mm), Level 0 = 1201 x 1201 x 81 state nodes.

  * one *subcycle block* = moveEverything (cF:2400-2510) + subcycleGOMELT with N2 = N3 = 5 (cF:3224-3632: 50 Level-3, 10
    Level-2 and 2 Level-1 sweeps, the T' projections, getNewTprime, face prolongations, melt-time bookkeeping) + the
    gather / scatter of the melt-time windows (gm:448-455);
  * one *single step* = moveEverything + stepGOMELT (cF:2304-2397) + the melt-time update (gm:339-357).

Timed with CUDA events over whole calls (device-resident state, host issue included; the median call is reported, every call listed);
the per-kernel table comes from
a CUPTI trace of one block and one step (torch.profiler sees every kernel of the process).  Returns a dict for the
bench line; run as a script it prints it.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2_INPUT = {
    "Level1": {"elements": [120, 120, 30], "bounds": {"x": [0, 24.0], "y": [0, 24.0], "z": [-4.0, 2.0]},
               "conditions": {"x": [298.15, 298.15], "y": [298.15, 298.15], "z": [298.15, 298.15]}},
    "Level2": {"elements": [500, 500, 40], "bounds": {"x": [2.0, 22.0], "y": [2.0, 22.0], "z": [-1.6, 0.0]}},
    "Level3": {"elements": [500, 500, 40], "bounds": {"x": [7.0, 17.0], "y": [7.0, 17.0], "z": [-0.8, 0.0]}},
}
N2 = N3 = 5
DT = 1e-5
V_LASER = 1000.0
B_INTERP = 8       # algorithmic bytes per target node of an interpolation pass: parent read (shared by ~8 targets at
                   # ratio 2) ~ 0 + result write 4 + base / second field 4 ("~8 B/target node")
B_PROJECT = 12     # per fine node of a correction projection: T' read 4 + T read 4 + S1 read 4 (the coefficient is
                   # evaluated in the kernel); the parent-side write is 1/8 .. 1/1000 of that


def _kernel_table(prof, total_us):
    rows = {}
    for ev in prof.events():
        if getattr(ev, "device_type", None) is None or "cuda" not in str(ev.device_type).lower():
            continue
        name = ev.name
        us = float(getattr(ev, "device_time", 0.0) or getattr(ev, "cuda_time", 0.0) or 0.0)
        short = name.replace("(anonymous namespace)::", "").split("<")[0].split("(")[0].replace("gomelt::", "").replace("void ", "")
        if short.startswith("at::") or "elementwise" in name or "vectorized" in name:
            short = "torch:" + short[:48]
        r = rows.setdefault(short, [0, 0.0])
        r[0] += 1
        r[1] += us
    tot = sum(v[1] for v in rows.values()) or 1.0
    table = [{"kernel": k, "launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)}
             for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])]
    return {"kernels": table[:16], "sum_kernel_us": round(tot, 1), "wall_us_of_the_traced_call": round(total_us, 1),
            "torch_kernels": sum(v[0] for k, v in rows.items() if k.startswith("torch:") or "Memcpy" in k or "Memset" in k)}


def run(props_in, peaks_gbs, blocks=5, warm=2):
    import numpy as np
    import torch

    import gomelt_b200 as gm

    cf = gm.computeFunctions
    inp = dict(C2_INPUT)
    inp["properties"] = dict(props_in, laser_center=[12.0, 12.0, 0.0, 0, 0, 0, 0])
    t0 = time.time()
    P = cf.SetupProperties(inp["properties"])
    Levels = cf.SetupLevels(inp, P)
    nn = [int(Levels[i]["nn"]) for i in range(4)]
    ne_nn = cf.getStaticNodesAndElements(Levels)
    subcycle = (N2, N3, N2 * N3, float(N2), float(N3), float(N2 * N3))
    h = [None] + [[float(v) for v in Levels[i]["h"]] for i in (1, 2, 3)]
    r12 = [int(round(h[1][0] / h[2][0])), int(round(h[1][1] / h[2][1])), int(round(P["layer_height"] / h[2][2]))]
    r23 = [int(round(h[2][i] / h[3][i])) for i in range(3)]
    laser_start = np.array(P["laser_center"], np.float32)
    laser = laser_start.copy()
    LInterp = [cf.interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
               cf.interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"])]
    tmp = cf.calcStaticTmpNodesAndElements(Levels, laser)
    move = [0, 0, 0]
    # substrate below z = 0 is bulk: S1 of Level 0 (the state grid the windows regather from)
    Levels, Shapes, LInterp, move = cf.moveEverything(laser, laser_start, Levels, move, LInterp, r12, r23, P["layer_height"])
    substrate = cf.getSubstrateNodes(Levels)
    Levels[0]["S1"][: substrate[0]] = 1.0
    accum, max_accum = torch.zeros(nn[0], device="cuda"), torch.zeros(nn[0], device="cuda")
    setup_s = time.time() - t0

    def rows_block():
        rows = np.zeros((N2 * N3, 7), np.float32)
        for i in range(N2 * N3):
            laser[0] += V_LASER * DT
            rows[i] = (laser[0], laser[1], laser[2], 1, 1, DT, P["laser_power"])
        return rows

    def block():
        nonlocal Levels, Shapes, LInterp, move
        rows = rows_block()
        Levels, Shapes, LInterp, move = cf.moveEverything(rows[0], laser_start, Levels, move, LInterp, r12, r23,
                                                          P["layer_height"])
        idx = Levels[0]["idx"]
        res = cf.subcycleGOMELT(Levels, ne_nn, Shapes, substrate, LInterp, tmp, rows, P, rows[:, 6], subcycle,
                                cf.take_box(max_accum, idx), cf.take_box(accum, idx))
        Levels = res[0]
        cf.put_box(max_accum, idx, res[4])
        cf.put_box(accum, idx, res[5])

    def step():
        nonlocal Levels, Shapes, LInterp, move
        laser[0] += V_LASER * DT
        row = np.array((laser[0], laser[1], laser[2], 1, 1, DT, P["laser_power"]), np.float32)
        Levels, Shapes, LInterp, move = cf.moveEverything(row, laser_start, Levels, move, LInterp, r12, r23,
                                                          P["layer_height"])
        Levels, reset = cf.stepGOMELT(Levels, ne_nn, tmp, Shapes, LInterp, row, P, row[5], row[6], substrate)
        cf.accumSingleStepFused(Levels, reset, accum, max_accum, row[5], P["T_liquidus"])

    def timed(fn, n):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        l0 = gm.ops.LAUNCHES
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        t = time.perf_counter()
        evs[0].record()
        for i in range(n):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t) / n
        each = sorted(evs[i].elapsed_time(evs[i + 1]) * 1e-3 for i in range(n))
        per_call.append([round(1e3 * v, 3) for v in each])
        # the median call: a call that lands on a first-time allocation of the caching allocator (a new window position
        # brings a new 41 MB cell-sum buffer) or a host hiccup takes several times longer and would own the mean
        return each[n // 2], wall, (gm.ops.LAUNCHES - l0) / n

    per_call = []
    # heat the window up first (a few steps with the laser on), so that the melt pool exists in what is timed
    for _ in range(3):
        step()
    blk_s, blk_wall, blk_launch = timed(block, blocks)
    stp_s, stp_wall, stp_launch = timed(step, max(blocks, 5))
    out = {
        "workload": "C2: Level 3 500x500x40 el (%d nodes, h=0.02) in Level 2 500x500x40 el (%d nodes, h=0.04) in Level 1 "
                    "120x120x30 el (%d nodes, h=0.2); Level 0 %d state nodes; N2=N3=5, dt=1e-5, laser moving at 1 m/s"
                    % (nn[3], nn[2], nn[1], nn[0]),
        "setup_s": round(setup_s, 2),
        "subcycle_block": {
            "what": "moveEverything + subcycleGOMELT (one native call: 50 L3 + 10 L2 + 2 L1 sweeps, projections, "
                    "getNewTprime, faces, bookkeeping) + gather / scatter of the melt-time windows",
            "ms": blk_s * 1e3, "ms_each_call_sorted": per_call[0], "host_wall_ms_mean": blk_wall * 1e3, "lib_launches": blk_launch,
            "L3_DOF_updates_per_s": 2 * N2 * N3 * nn[3] / blk_s,
            "all_levels_DOF_updates_per_s": (2 * N2 * N3 * nn[3] + 2 * N2 * nn[2] + 2 * nn[1]) / blk_s,
            "algorithmic_GBps": (2 * N2 * N3 * nn[3] * 16 + 2 * N2 * nn[2] * 20 + 2 * nn[1] * 16) / blk_s / 1e9,
            "frac_of_hbm_peak": (2 * N2 * N3 * nn[3] * 16 + 2 * N2 * nn[2] * 20 + 2 * nn[1] * 16) / blk_s / 1e9 / peaks_gbs,
            "sim_s_per_block": N2 * N3 * DT, "wall_s_per_sim_s": blk_s / (N2 * N3 * DT),
        },
        "single_step": {
            "what": "moveEverything + stepGOMELT (one native call: 2 x (L1 + L2 solves, faces), 1 L3 solve, 6 projections, "
                    "4 getNewTprime) + melt-time update",
            "ms": stp_s * 1e3, "ms_each_call_sorted": per_call[1], "host_wall_ms_mean": stp_wall * 1e3, "lib_launches": stp_launch,
            "wall_s_per_sim_s": stp_s / DT,
        },
    }
    # per-kernel table (CUPTI) + bandwidth rooflines of the transfer kernels
    try:
        from torch.profiler import ProfilerActivity, profile

        tables = {}
        for name, fn in (("subcycle_block", block), ("single_step", step)):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
            tables[name] = _kernel_table(prof, e0.elapsed_time(e1) * 1e3)
        out["kernel_tables"] = tables
        roof = {}
        for k in tables["subcycle_block"]["kernels"]:
            if k["kernel"].startswith("interp_kernel"):
                roof["interp_kernel"] = {"bytes_per_target_node": B_INTERP, "avg_us": k["us"] / k["launches"],
                                         "note": "mix of full-window passes (T' = T - I(parent): %d targets) and small "
                                                 "overlap / face passes; GB/s below is for the full-window RSUB passes "
                                                 "only if they dominate" % nn[3]}
            if k["kernel"].startswith("project_cells_kernel"):
                gbs = B_PROJECT * nn[3] / (k["us"] / k["launches"] * 1e-6) / 1e9
                roof["project_cells_kernel"] = {"bytes_per_fine_node": B_PROJECT, "avg_us": k["us"] / k["launches"],
                                                "achieved_GBps_if_all_L3": gbs, "frac": gbs / peaks_gbs,
                                                "note": "15 of 19 launches per block project the 10 M-node Level 3"}
            if k["kernel"].startswith("shift_window_kernel"):
                gbs = B_INTERP * (nn[3] + nn[2]) / (k["us"] * 1e-6) / 1e9
                roof["shift_window_kernel"] = {"bytes_per_target_node": B_INTERP, "us_both_windows": k["us"],
                                               "achieved_GBps": gbs, "frac": gbs / peaks_gbs}
        out["transfer_rooflines"] = roof
    except Exception as exc:  # the timings above stand on their own
        out["kernel_tables"] = {"error": repr(exc)}
    del Levels, accum, max_accum
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import bench

    print(json.dumps(run(bench.EXAMPLE_PROPS, bench.read_peaks()["hbm_gbs"])))
