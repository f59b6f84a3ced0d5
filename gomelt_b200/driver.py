"""``go_melt(solver_input)``: the reference driver loop (go_melt.py gm:16-530) restated JAX-free on top of
the drop-in ``computeFunctions`` namespace, with every field resident on the GPU.

The reference driver mixes control flow with ~20 inline ``jax.numpy`` updates of ``Levels`` fields
(gm:215-224, 264-290, 313, 339-357, 360-368, 448-455).  Here those updates go through a small array
adapter (``TorchArrays``: torch CUDA tensors), the control flow - block reads of N2*N3 toolpath rows, the
single-step / subcycle / dwell mode predicate (gm:162-172), layer change, wait counter - is the
reference's.  The same function drives the NumPy oracle in the tests (``cf`` / ``xp`` arguments), which
is how the loop itself is parity-checked.

File output is behind hooks (``on_record`` / ``on_layer_state`` / ``on_checkpoint`` / ``on_info`` /
``on_final``), called where the reference writes its files; ``output.driver_hooks()`` supplies the
dependency-free implementations (``.vtr`` files with the reference's names, raw-dump checkpoints, fused
min / max monitor) and the CLI installs them.  The checkpoint *flags* that influence the stepping mode
(gm:179-193, 390-411) are kept in the loop itself.

CLI (gm:533-580):  python gomelt_b200/driver.py [DEVICE_ID] [input.json]
"""
import copy
import json
import os
import sys
import time

import numpy as np

F32 = np.float32


class TorchArrays:
    """Array adapter of the driver: float32 fields as torch CUDA tensors, updated in place."""

    def __init__(self):
        from . import _lib

        self.torch = _lib.require_cuda()

    def _idx(self, idx):
        t = self.torch
        return idx.long() if isinstance(idx, t.Tensor) else t.as_tensor(np.asarray(idx, dtype=np.int64)).cuda()

    def zeros(self, n):
        return self.torch.zeros(int(n), device="cuda", dtype=self.torch.float32)

    def f32(self, x):
        t = self.torch
        return x.float() if isinstance(x, t.Tensor) else t.as_tensor(np.asarray(x, dtype=F32)).cuda()

    def maximum(self, a, b):
        t = self.torch
        return t.maximum(a, b) if isinstance(b, t.Tensor) else t.clamp_min(a, float(b))

    def take(self, a, idx):
        if hasattr(idx, "vectors"):  # computeFunctions.BoxIndex: gather through the three index vectors
            from .computeFunctions import take_box

            return take_box(a, idx)
        return a[self._idx(idx)]

    def put(self, a, idx, v):
        if hasattr(idx, "vectors"):
            from .computeFunctions import put_box

            return put_box(a, idx, v)
        a[self._idx(idx)] = v
        return a

    def add_at(self, a, idx, v):  # idx has no repeats (a window's node set)
        a[self._idx(idx)] += v
        return a

    def shift_down(self, a, n1, n2):
        """a[:n2] = a[n1:]; a[n2:] = 0  (gm:275-278, 288-289)."""
        out = self.torch.zeros_like(a)
        out[:n2] = a[n1:]
        return out

    def fill_prefix(self, a, n, value):
        a[: int(n)] = value
        return a

    def all_zero(self, a):
        return not bool((a != 0).any())

    def zeros_like(self, a):
        return self.torch.zeros_like(a)

    def set_row(self, m, i, v):
        m[i, :] = v
        return m

    def get_row(self, m, i):
        return m[i, :].clone()

    def positive(self, a):
        return (self.f32(a) > 0).to(self.torch.float32)

    def host(self, a):
        return a.detach().cpu().numpy() if isinstance(a, self.torch.Tensor) else np.asarray(a)

    def sync(self):
        self.torch.cuda.synchronize()


def _read_block(fh, nrows):
    """gm:140-160: the next ``nrows`` toolpath rows -> (float32 array [n, 7], end_of_file)."""
    rows = []
    for _ in range(nrows):
        line = fh.readline()
        if line.strip() == "":
            return np.array(rows, dtype=F32).reshape(-1, 7), True
        rows.append([float(v) for v in line.split(",")])
    return np.array(rows, dtype=F32), False


def _needs_single_step(laser_all, laser_prev_z, load_chkpt, wait_inc, nonmesh, subcycle, ongoing):
    """gm:162-172."""
    if (laser_all[:, 2] != F32(laser_prev_z)).any() or load_chkpt or not ongoing or laser_all.shape[0] == 1:
        return True
    if wait_inc > max(0, nonmesh["wait_time"] - subcycle[0] * subcycle[1] * 2):
        return True
    speed = np.abs(np.diff(laser_all, axis=0)[:, :2] / laser_all[:-1, 5].max())
    return bool((speed > F32(100 * nonmesh["laser_velocity"])).any())


def plan_toolpath(toolpath_file, Nonmesh, subcycle):
    """The stepping schedule of a toolpath file, computed on the host ahead of the run: the mode
    predicate of gm:162-172 and the wait counter of gm:179 / 415-419 depend on the toolpath rows only (for a run that
    does not restart from a checkpoint), so the sequence of calls the driver will make is known before the first kernel
    is launched.  Returns a list of blocks: {"mode": "single" | "subcycle", "rows": n, "steps": s, "dwells": d,
    "layer_changes": c, "dwell_runs": [lengths of runs of identical dwell rows: what dwellRows replays as graphs]}."""
    nblock = subcycle[0] * subcycle[1]
    blocks = []
    wait_inc, laser_prev_z, ongoing, load_chkpt, time_inc = 0, float("inf"), True, False, 0
    with open(toolpath_file, "r") as fh:
        while ongoing:
            laser_all, eof = _read_block(fh, nblock)
            if eof:
                ongoing = False
                if laser_all.shape[0] == 0:
                    break
            if _needs_single_step(laser_all, laser_prev_z, load_chkpt, wait_inc, Nonmesh, subcycle, ongoing):
                b = {"mode": "single", "rows": int(laser_all.shape[0]), "steps": 0, "dwells": 0, "layer_changes": 0,
                     "dwell_runs": []}
                new_checkpoint, prev, run = False, None, 0
                for row in laser_all:
                    wait_inc = wait_inc + 1 if row[4] == 0 else 0
                    if row[2] != F32(laser_prev_z):
                        if time_inc > 0 and not load_chkpt:
                            new_checkpoint = True
                        b["layer_changes"] += 1
                        laser_prev_z = float(row[2])
                        wait_inc = 0
                    dwell = wait_inc > Nonmesh["wait_time"]
                    b["dwells" if dwell else "steps"] += 1
                    if dwell and prev is not None and np.array_equal(prev, row):
                        run += 1
                    else:
                        if run > 1:
                            b["dwell_runs"].append(run)
                        run = 1 if dwell else 0
                    prev = row if dwell else None
                    time_inc += 1
                if run > 1:
                    b["dwell_runs"].append(run)
                load_chkpt = new_checkpoint
                blocks.append(b)
            else:
                off = laser_all[:, 4] == 0
                wait_inc = wait_inc + len(laser_all) - int(laser_all[:, 4].sum()) if off.any() else 0
                time_inc += laser_all.shape[0]
                blocks.append({"mode": "subcycle", "rows": int(laser_all.shape[0]), "steps": 0, "dwells": 0,
                               "layer_changes": 0, "dwell_runs": []})
    return blocks


def go_melt(solver_input, cf=None, xp=None, hooks=None, verbose=False, write_final=True, graphs=True):
    """Run the whole simulation.  Returns a dict: ``Levels``, ``accum_time``, counters and timings.  ``graphs=False``
    issues every row of a pause eagerly instead of replaying CUDA graphs (A/B; the results are identical)."""
    if cf is None:
        from . import computeFunctions as cf
    if xp is None:
        xp = TorchArrays()
    hooks = hooks or {}
    say = print if verbose else (lambda *a, **k: None)
    tstart = time.time()

    Properties = cf.SetupProperties(solver_input.get("properties", {}))
    Levels = cf.SetupLevels(solver_input, Properties)
    Nonmesh = cf.SetupNonmesh(solver_input.get("nonmesh", {}))
    ne_nn = cf.getStaticNodesAndElements(Levels)
    subcycle = cf.getStaticSubcycle(Nonmesh)
    h = [None] + [[float(v) for v in Levels[i]["h"]] for i in (1, 2, 3)]
    L1L2Eratio = [int(np.round(F32(h[1][i]) / F32(h[2][i]))) for i in range(2)] + [
        int(np.round(F32(Properties["layer_height"]) / F32(h[2][2])))]
    L2L3Eratio = [int(np.round(F32(h[2][i]) / F32(h[3][i]))) for i in range(3)]

    total_t_inc = cf.count_lines(Nonmesh["toolpath"]) if Nonmesh["use_txt"] else cf.parsingGcode(
        Nonmesh, Properties, Levels[2]["h"])
    if not Properties["laser_center"]:
        with open(Nonmesh["toolpath"], "r") as fh:
            laser_start = np.array([float(v) for v in fh.readline().split(",")])
    else:
        laser_start = np.array(Properties["laser_center"])
    LInterp = [cf.interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
               cf.interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"])]

    time_inc = record_inc = wait_inc = 0
    t_output = 0.0
    laser_prev_z = float("inf")
    dwell_count = 0.0
    # slab-decomposed Level 1 (computeFunctions.enable_distributed): every rank runs this loop; `worker` ranks hold only a
    # Level-1 slab - no windows, no Level 0, no melt-time arrays - and serve the Level-1 side of every call
    dist_on = hasattr(cf, "distOf") and cf.distOf(Levels) is not None
    worker = dist_on and cf.isWorker(Levels)
    batch_dwell = bool(graphs) and hasattr(cf, "dwellRows") and not dist_on

    def call(name, *args, storage=False):
        """A file-output hook.  In a distributed run the Level-1 fields it sees are assembled from the slabs first (a
        collective: every rank passes here; the ranks without windows then skip the hook)."""
        if name not in hooks:
            return
        if dist_on:
            view = cf.outputView(Levels, with_storage=storage)
            if view is None:
                return
            args = (view,) + args[1:]
        hooks[name](*args)
    nn0 = int(Levels[0]["nn"])
    accum_time, max_accum_time = (None, None) if worker else (xp.zeros(nn0), xp.zeros(nn0))
    move_hist = [0, 0, 0]
    force_move = move_vert = new_checkpoint = load_chkpt = False
    ongoing = True
    stopped_at_layer_check = False
    tprime_test_done = False
    layer_check = Nonmesh["layer_num"] + Nonmesh["restart_layer_num"]
    call("on_record", Levels, Nonmesh, int(time_inc / Nonmesh["record_step"]) + 1)  # the initial saveResults of gm:83-84
    counts = {"stepGOMELT": 0, "subcycleGOMELT": 0, "stepGOMELTDwellTime": 0, "moveEverything": 0, "layers": 0}
    Shapes = tmp_ne_nn = substrate = None
    nblock = subcycle[0] * subcycle[1]
    T_amb = Properties["T_amb"]

    fh = open(Nonmesh["toolpath"], "r")
    if Nonmesh["layer_num"] > 0:  # ---- restart from the checkpoint of this layer gm:108-129 ----
        if "load_checkpoint" not in hooks:
            fh.close()
            raise RuntimeError(f"nonmesh.layer_num = {Nonmesh['layer_num']} asks for a restart from "
                               f"Checkpoint{str(Nonmesh['layer_num']).zfill(4)}, but no 'load_checkpoint' hook was given "
                               "(use hooks=output.driver_hooks())")
        say(f"Checkpoint loading for start of Layer {Nonmesh['layer_num']}")
        loaded, accum_time, max_accum_time, time_inc_loaded, record_inc = hooks["load_checkpoint"](
            Nonmesh, "cuda" if (isinstance(xp, TorchArrays) and not worker) else None)
        # (a checkpoint holds the whole part-scale field: a distributed run cuts its slabs out of it)
        Levels = cf.adoptCheckpoint(Levels, loaded) if dist_on else loaded
        accum_time, max_accum_time = (None, None) if worker else (xp.f32(accum_time), xp.f32(max_accum_time))
        load_chkpt = True
        line_len = len(fh.readline())
        fh.seek(int(time_inc_loaded) * line_len)
        time_inc += int(time_inc_loaded)
    try:
        while ongoing:
            t_loop = time.time()
            laser_all, eof = _read_block(fh, nblock)
            t_add = laser_all.shape[0]
            if eof:
                ongoing = False
                if t_add == 0:
                    break
            single_step = _needs_single_step(laser_all, laser_prev_z, load_chkpt, wait_inc, Nonmesh, subcycle, ongoing)
            if single_step:
                jrow = 0
                while jrow < laser_all.shape[0]:
                    laser_pos = laser_all[jrow]
                    jrow += 1
                    # a run of identical rows in Level-1-only mode (the pause between tracks / layers): handed to the
                    # stepper side as ONE call, which replays them as CUDA graphs (computeFunctions.dwellRows; N2)
                    if (batch_dwell and laser_pos[4] == 0 and wait_inc + 1 > Nonmesh["wait_time"] and not move_vert
                            and laser_pos[2] == F32(laser_prev_z) and tprime_test_done):
                        m = 1
                        while jrow - 1 + m < laser_all.shape[0] and np.array_equal(laser_all[jrow - 1 + m], laser_pos):
                            m += 1
                        if m >= 3:
                            Levels, Shapes, LInterp, move_hist = cf.dwellRows(
                                Levels, m, laser_pos, laser_start, move_hist, LInterp, L1L2Eratio, L2L3Eratio,
                                Properties["layer_height"], tmp_ne_nn, ne_nn, Properties, laser_pos[5], substrate)
                            counts["moveEverything"] += m
                            counts["stepGOMELTDwellTime"] += m
                            for _ in range(m):
                                dwell_count += float(laser_pos[5])
                            wait_inc += m
                            time_inc += m
                            record_inc += m
                            jrow += m - 1
                            continue
                    wait_inc = wait_inc + 1 if laser_pos[4] == 0 else 0
                    if laser_pos[2] != F32(laser_prev_z) and time_inc > 0 and not load_chkpt:
                        new_checkpoint = True
                    if laser_pos[2] != F32(laser_prev_z):  # ---- layer change gm:198-290 ----
                        tmp_coords = copy.deepcopy(Levels[1]["orig_node_coords"])
                        state_idx = 0
                        while not np.isclose(np.asarray(tmp_coords[2]) - laser_pos[2], 0, atol=1e-4).any():
                            tmp_coords[2] = (np.asarray(tmp_coords[2], F32) + F32(Properties["layer_height"])).astype(F32)
                            state_idx += 1
                        if not load_chkpt and dist_on:
                            Levels = cf.layerShiftL1(Levels, tmp_coords, state_idx, F32(T_amb))
                        elif not load_chkpt:
                            Levels[1]["T0"] = xp.maximum(xp.f32(cf.interpolatePoints(Levels[1], Levels[1]["T0"], tmp_coords)),
                                                         F32(T_amb))
                            Levels[1]["S1_storage"] = xp.set_row(xp.f32(Levels[1]["S1_storage"]), state_idx - 1,
                                                                 xp.f32(Levels[1]["S1"]))
                            Levels[1]["S1"] = xp.get_row(Levels[1]["S1_storage"], state_idx)
                            Levels[1]["node_coords"] = copy.deepcopy(tmp_coords)
                        LInterp = [cf.interpolatePointsMatrix(Levels[1], Levels[2]["node_coords"]),
                                   cf.interpolatePointsMatrix(Levels[2], Levels[3]["node_coords"])]
                        tmp_ne_nn = cf.calcStaticTmpNodesAndElements(Levels, laser_pos)
                        laser_prev_z = float(laser_pos[2])
                        force_move = True
                        wait_inc = 0
                        move_vert = True
                        counts["layers"] += 1
                        if not load_chkpt and not worker:
                            if "on_layer_state" in hooks:  # saveState(Level0) gm:245 (Level 0 lives on the owner)
                                hooks["on_layer_state"](Levels, Nonmesh)
                            accum_time = xp.maximum(accum_time, max_accum_time)
                            if "on_layer_accum" in hooks:  # accum_time<layer_num>.npz gm:253-261, before the shift
                                hooks["on_layer_accum"](accum_time, Nonmesh)
                            L0 = Levels[0]
                            nxy = int(L0["nodes"][0]) * int(L0["nodes"][1])
                            n1 = nxy * int(L0["layer_idx_delta"])
                            n2 = nxy * (int(L0["nodes"][2]) - int(L0["layer_idx_delta"]))
                            L0["S1"] = xp.shift_down(xp.f32(L0["S1"]), n1, n2)
                            L0["node_coords"][2] = (np.asarray(L0["orig_node_coords"][2], F32) + laser_pos[2]
                                                    - F32(np.asarray(L0["orig_node_coords"][2])[-1])).astype(F32)
                            max_accum_time = xp.zeros(nn0)
                            accum_time = xp.shift_down(accum_time, n1, n2)
                    force_move = True  # gm:292: the windows are re-placed on every single-step row
                    if force_move:
                        force_move = False
                        Levels, Shapes, LInterp, move_hist = cf.moveEverything(
                            laser_pos, laser_start, Levels, move_hist, LInterp, L1L2Eratio, L2L3Eratio,
                            Properties["layer_height"])
                        counts["moveEverything"] += 1
                        if move_vert:
                            move_vert = False
                            substrate = cf.getSubstrateNodes(Levels)
                            if not worker:
                                Levels[0]["S1"] = xp.fill_prefix(xp.f32(Levels[0]["S1"]), substrate[0], 1.0)
                    if wait_inc <= Nonmesh["wait_time"]:
                        Levels, all_reset = cf.stepGOMELT(Levels, ne_nn, tmp_ne_nn, Shapes, LInterp, laser_pos, Properties,
                                                          laser_pos[5], laser_pos[6], substrate)
                        counts["stepGOMELT"] += 1
                        tprime_test_done = False
                        if worker:
                            pass
                        elif hasattr(cf, "accumSingleStepFused"):  # gm:339-357 + melting_temp as one kernel, in place
                            accum_time, max_accum_time = cf.accumSingleStepFused(
                                Levels, all_reset, accum_time, max_accum_time, laser_pos[5], Properties["T_liquidus"])
                        else:
                            idx = Levels[0]["idx"]  # gm:339-357
                            reset = xp.take(accum_time, idx) * xp.positive(all_reset)
                            max_accum_time = xp.put(max_accum_time, idx, xp.maximum(reset, xp.take(max_accum_time, idx)))
                            accum_time = xp.add_at(accum_time, idx, -reset)
                            accum_time = cf.melting_temp(Levels[3]["T0"], laser_pos[5], Properties["T_liquidus"],
                                                         accum_time, idx)
                    else:
                        # gm:360-368.  The outcome of this test (a device -> host round trip) can only change when a
                        # stepper has written T'0 again, so it is evaluated once per run of dwell rows.
                        if not tprime_test_done and not worker:
                            tprime_test_done = True
                            if not xp.all_zero(xp.f32(Levels[2]["Tprime0"])) and not xp.all_zero(xp.f32(Levels[3]["Tprime0"])):
                                dwell_count = Nonmesh["wait_time"] * Nonmesh["timestep_L3"]
                                Levels[2]["Tprime0"] = xp.zeros_like(xp.f32(Levels[2]["Tprime0"]))
                                Levels[3]["Tprime0"] = xp.zeros_like(xp.f32(Levels[3]["Tprime0"]))
                        Levels = cf.stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, Properties, laser_pos[5], substrate)
                        counts["stepGOMELTDwellTime"] += 1
                        dwell_count += float(laser_pos[5])
                    time_inc += 1
                    record_inc += 1
                if new_checkpoint:
                    Nonmesh["layer_num"] += 1
                    call("on_checkpoint", Levels, accum_time, max_accum_time, time_inc, record_inc, Nonmesh,
                         storage=True)  # dill dump gm:390-402
                    if Nonmesh["layer_num"] == layer_check:  # gm:405-406: return right here, no final output
                        stopped_at_layer_check = True
                        break
                    new_checkpoint = False
                    load_chkpt = True
                else:
                    load_chkpt = False
            else:  # ---- subcycling gm:413-459 ----
                off = laser_all[:, 4] == 0
                wait_inc = wait_inc + len(laser_all) - int(laser_all[:, 4].sum()) if off.any() else 0
                Levels, Shapes, LInterp, move_hist = cf.moveEverything(
                    laser_all[0, :], laser_start, Levels, move_hist, LInterp, L1L2Eratio, L2L3Eratio,
                    Properties["layer_height"])
                counts["moveEverything"] += 1
                idx = Levels[0]["idx"]
                res = cf.subcycleGOMELT(Levels, ne_nn, Shapes, substrate, LInterp, tmp_ne_nn, laser_all, Properties,
                                        laser_all[:, 6], subcycle, None if worker else xp.take(max_accum_time, idx),
                                        None if worker else xp.take(accum_time, idx))
                Levels, _max_accum, _accum = res[0], res[4], res[5]
                counts["subcycleGOMELT"] += 1
                tprime_test_done = False
                if not worker:
                    max_accum_time = xp.put(max_accum_time, idx, xp.f32(_max_accum))
                    accum_time = xp.put(accum_time, idx, xp.f32(_accum))
                time_inc += t_add
                record_inc += t_add
            t_output += float(laser_all[:, 5].sum(dtype=F32))
            if record_inc >= Nonmesh["record_step"]:
                record_inc = 0
                call("on_record", Levels, Nonmesh, int(time_inc / Nonmesh["record_step"]) + 1)  # saveResults gm:467-470
            if Nonmesh["info_T"]:  # printLevelMaxMin gm:473-474 (forces a sync)
                call("on_info", Levels)
            if verbose:
                tend = time.time()
                say("%d/%d, Real: %.6f s, Wall: %.2f s, Loop: %5.2f ms, Avg: %5.2f ms/dt"
                    % (time_inc, total_t_inc, t_output, tend - tstart, 1000 * (tend - t_loop),
                       1000 * (tend - tstart) / max(time_inc, 1)))
    finally:
        fh.close()
    if not worker:
        accum_time = xp.maximum(accum_time, max_accum_time)  # gm:512
    if hasattr(xp, "sync"):
        xp.sync()
    wall = time.time() - tstart
    if not stopped_at_layer_check:  # saveState(Level 0) + saveResultsFinal gm:501-502
        call("on_final", Levels, Nonmesh)
    L1_full = None
    if dist_on:  # the owner's Level-1 arrays are mirrors of the box under the windows: assemble the part-scale field
        L1_full = cf.gatherL1(Levels)
        if worker:
            return {"Levels": Levels, "accum_time": None, "time_inc": time_inc, "total_t_inc": total_t_inc,
                    "sim_seconds": t_output, "wall_seconds": wall, "counts": counts, "Properties": Properties,
                    "Nonmesh": Nonmesh, "ne_nn": ne_nn, "dwell_seconds": dwell_count, "worker": True,
                    "stopped_at_layer_check": stopped_at_layer_check}
        Levels[1]["T0"] = L1_full
    if write_final and not stopped_at_layer_check:  # gm:504-516
        np.savez(f"{Nonmesh['save_path']}FinalTemperatureFields", L1T=xp.host(Levels[1]["T0"]),
                 L2T=xp.host(Levels[2]["T0"]), L3T=xp.host(Levels[3]["T0"]))
        np.savez(Nonmesh["save_path"] + "accum_time" + str(Nonmesh["layer_num"]).zfill(4), accum_time=xp.host(accum_time))
    return {"Levels": Levels, "accum_time": accum_time, "time_inc": time_inc, "total_t_inc": total_t_inc,
            "sim_seconds": t_output, "wall_seconds": wall, "counts": counts, "Properties": Properties, "Nonmesh": Nonmesh,
            "ne_nn": ne_nn, "dwell_seconds": dwell_count, "stopped_at_layer_check": stopped_at_layer_check}


def main(argv):
    """gm:533-580: ``[DEVICE_ID] [input.json]``."""
    device = argv[1] if len(argv) > 1 else "0"
    path = argv[2] if len(argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                       "examples", "example.json")
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(device))
    try:
        with open(path, "r") as fh:
            solver_input = json.load(fh)
    except FileNotFoundError:
        print(f"Input file not found: {path}")
        return 1
    from . import output  # the reference's file output (.vtr records, checkpoints, min / max monitor)

    out = go_melt(solver_input, verbose=True, hooks=output.driver_hooks())
    print(f"End of simulation: {out['time_inc']} steps, {out['sim_seconds']:.6f} s simulated in "
          f"{out['wall_seconds']:.2f} s wall ({out['wall_seconds'] / max(out['sim_seconds'], 1e-30):.1f} wall-s per sim-s)")
    return 0


if __name__ == "__main__":
    import importlib

    _root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, _root)
    sys.exit(importlib.import_module("gomelt_b200.driver").main(sys.argv))
