"""Bit-parity of the drop-in on a slab-decomposed Level 1 (gomelt_b200/dist.py) inside the multi-GPU bench line: every
rank runs the driver on ``wide_part_input`` (examples/example.json with a 73 x 21 x 31-node part-scale level, a short
two-layer G-code with a pause: window moves, single steps, subcycle blocks, dwell steps, a layer change), the laser owner
repeats the run alone (plain single-GPU drop-in) and compares every level bit for bit."""
import json
import os
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def wide_part_input(tmp):
    with open(os.path.join(ROOT, "examples", "example.json")) as fh:
        inp = json.load(fh)
    inp["Level1"]["elements"] = [72, 20, 30]
    inp["Level1"]["bounds"]["x"] = [0, 14.4]
    g = os.path.join(tmp, "wide.gcode")
    with open(g, "w") as fh:
        fh.write("G0 X2.0 Y2.0 Z0.04\nG1 X2.24 Y2.0 Z0.04\nG0 X2.24 Y2.2 Z0.08\nG1 X2.04 Y2.2 Z0.08\n")
    inp["nonmesh"].update(save_path=tmp + "/", toolpath=os.path.join(tmp, "toolpath.txt"), gcode=g, use_txt=0,
                          wait_time=6, dwell_time=2e-4, dwell_time_multiplier=1, subcycle_num_L2=2,
                          subcycle_num_L3=2, record_step=1000, info_T=0, output_files=0)
    return inp


def run(torch, dist, gm, rank, world):
    """Collective.  Returns the check's dict on rank 0 (None elsewhere)."""
    cf = gm.computeFunctions
    tmp = tempfile.mkdtemp(prefix=f"gomelt_dist_r{rank}_")
    res = None
    try:
        cf.enable_distributed(rank, world)
        out = gm.driver.go_melt(wide_part_input(tmp), write_final=False)
        torch.cuda.synchronize()
        S1 = cf.gatherL1(out["Levels"], "S1")
        d = cf.distOf(out["Levels"])
        owner = d.owner
        if not out.get("worker"):
            L = out["Levels"]
            got = {"L1T": L[1]["T0"], "L1S1": S1, "L2T": L[2]["T0"], "L3T": L[3]["T0"], "accum": out["accum_time"]}
            cf.disable_distributed()
            tmp2 = tempfile.mkdtemp(prefix="gomelt_plain_")
            ref = gm.driver.go_melt(wide_part_input(tmp2), write_final=False)
            torch.cuda.synchronize()
            R = ref["Levels"]
            want = {"L1T": R[1]["T0"], "L1S1": R[1]["S1"], "L2T": R[2]["T0"], "L3T": R[3]["T0"], "accum": ref["accum_time"]}
            diff = {k: int((got[k] != want[k]).sum().item()) for k in got}
            res = {"ok": all(v == 0 for v in diff.values()) and float(want["L3T"].max()) > 1000.0,
                   "values_differing": diff, "owner_rank": owner, "ranks": world,
                   "L1_nodes": int(np.prod(R[1]["nodes"])), "toolpath_rows": int(out["time_inc"]), "counts": out["counts"],
                   "level1_solves": d.stats["solves"], "boxes_down": d.stats["boxes_down"], "boxes_up": d.stats["boxes_up"],
                   "box_bytes": d.stats["bytes"], "max_T3_K": float(want["L3T"].max()),
                   "how": "driver.go_melt on all ranks (Level 1 in z-slabs, windows on the laser owner, Level-1 solves through "
                          "gomelt_hier_t.l1_solve) vs the plain single-GPU drop-in on the owner's GPU: every value of Level-1 T "
                          "and S1 (assembled from the slabs), Level-2 / Level-3 T and the melt-time field compared bit for bit"}
    finally:
        cf.disable_distributed()
    box = [res if rank == owner else None]
    if world > 1:
        dist.broadcast_object_list(box, src=owner)
    return box[0] if rank == 0 else None
