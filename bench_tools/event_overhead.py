"""How much of bench.py's per-launch CUDA-event time of K1 is the host's issue path?  The L3-10M substep call timed
(a) through ops.level_step (Python wrapper builds the argument struct between the two event records), (b) through the raw
C-ABI call on a prebuilt struct, (c) five such launches between one pair of events, (d) the raw call with the GPU kept
busy by a preceding kernel, so that the start event cannot be stamped before the launch is in the queue."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import gomelt_b200 as gm  # noqa: E402
from gomelt_b200 import _lib  # noqa: E402

gm.load()
torch.cuda.set_device(0)
blk = bench.L3Block()
for _ in range(3):
    blk.block()
ops = gm.ops
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
rows = blk._rows()
coef = ops.source_tables(blk.props, blk.grid, blk.coords, rows[0, :3], float(rows[0, 6]), blk.tx, blk.ty, blk.tz)
cur, other = blk.cur, (blk.Tb if blk.cur is blk.Ta else blk.Ta)
flags = blk.step_flags | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX
a = _lib.StepArgs()
a.grid = blk.grid
a.src_x, a.src_y, a.src_z, a.src_coef = blk.tx.data_ptr(), blk.ty.data_ptr(), blk.tz.data_ptr(), float(coef)
a.dt, a.nz_active, a.n_substrate, a.flags = bench.DT, blk.nodes[2], int(blk.n_sub), int(flags)
a.S1, a.S1_out = blk.S1.data_ptr(), blk.S1.data_ptr()
stream = _lib.stream_ptr()
out = {}


def med(ts):
    ts = sorted(ts)
    return ts[len(ts) // 2]


def timed(fn, n=1, busy=False):
    ts = []
    for it in range(14):
        flush.zero_()
        if busy:
            flush.mul_(1.0)   # ~80 us of GPU work queued ahead of the start event
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 4:
            ts.append(e0.elapsed_time(e1) * 1e3 / n)
    return med(ts)


state = {"cur": cur, "other": other}


def wrapper():
    ops.level_step(blk.props, blk.grid, state["cur"], blk.S1, state["other"], bench.DT, src=(blk.tx, blk.ty, blk.tz, coef),
                   n_substrate=blk.n_sub, flags=flags, S1_out=blk.S1)
    state["cur"], state["other"] = state["other"], state["cur"]


def raw():
    a.T0, a.T_out = state["cur"].data_ptr(), state["other"].data_ptr()
    rc = lib.gomelt_level_step_f32(C.byref(blk.props), C.byref(a), stream)
    assert rc == 0
    state["cur"], state["other"] = state["other"], state["cur"]


out["a_python_wrapper_us"] = timed(wrapper)
out["b_raw_c_abi_us"] = timed(raw)
out["c_five_raw_launches_per_event_pair_us"] = timed(raw, n=5)
out["d_raw_behind_queued_work_us"] = timed(raw, busy=True)
out["d5_five_raw_behind_queued_work_us"] = timed(raw, n=5, busy=True)
print(json.dumps(out))
