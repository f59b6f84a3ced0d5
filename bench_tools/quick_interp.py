"""Full-grid interpolation (getNewTprime's T' = child - I(parent), cF:2060-2099) and the window shift of moveEverything
(cF:2439-2443 / 2460-2464) at C2 size, one call at a time: the marching kernel against the per-target kernel (the same
call with an identity index map), and whether they give the same bits.  CUDA events, L2 flushed between calls.
python bench_tools/quick_interp.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gomelt_b200 as gm  # noqa: E402


def make_level(ne, origin, h):
    nodes = [n + 1 for n in ne]
    coords = [(np.float32(o) + np.arange(n, dtype=np.float32) * np.float32(h)).astype(np.float32) for o, n in zip(origin, nodes)]
    return {"nodes": nodes, "nn": int(np.prod(nodes)), "node_coords": coords}


def dev(lv):
    return [torch.as_tensor(c).cuda() for c in lv["node_coords"]]


def timed(fn, flush, n=8):
    if "--once" in sys.argv:   # under ncu: one launch per case and variant
        fn()
        torch.cuda.synchronize()
        return 1.0
    ts = []
    for it in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return round(ts[len(ts) // 2], 1)


def main():
    gm.load()
    torch.cuda.set_device(0)
    ops, lib = gm.ops, gm._lib
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    L1 = make_level((120, 120, 30), (0.0, 0.0, -6.0), 0.2)
    L2 = make_level((500, 500, 40), (2.0, 2.0, -1.6), 0.04)
    L3 = make_level((500, 500, 40), (7.0, 7.0, -0.8), 0.02)
    L3n = make_level((500, 500, 40), (7.0 + 3 * 0.04, 7.0 - 2 * 0.04, -0.8), 0.02)   # moved by (3, -2, 0) Level-2 cells
    L2n = make_level((500, 500, 40), (2.0 + 0.2, 2.0 - 0.2, -1.6), 0.04)             # moved by one Level-1 cell
    f = lambda lv, lo, hi: lo + (hi - lo) * torch.rand(lv["nn"], device="cuda", generator=g)
    T1, T2, T3 = f(L1, 300, 900), f(L2, 300, 1500), f(L3, 300, 2000)
    Tp2, Tp3 = f(L2, -30, 30), f(L3, -30, 30)
    c1, c2, c3, c3n, c2n = dev(L1), dev(L2), dev(L3), dev(L3n), dev(L2n)
    out3, out2 = torch.empty(L3["nn"], device="cuda"), torch.empty(L2["nn"], device="cuda")
    a3, b3, a2, b2 = (torch.empty(L3["nn"], device="cuda") for _ in range(4))
    ident3 = [torch.arange(n, dtype=torch.int32, device="cuda") for n in L3["nodes"]] + [L3["nodes"][0], L3["nodes"][1]]
    ident2 = [torch.arange(n, dtype=torch.int32, device="cuda") for n in L2["nodes"]] + [L2["nodes"][0], L2["nodes"][1]]
    cases = {
        "tprime_L2_to_L3 (ratio 2, RSUB)": (lambda **kw: ops.interp(c2, T2, c3, out3, mode=lib.INTERP_RSUB, base=T3, **kw), ident3, out3, 8 * L3["nn"]),
        "tprime_L1_to_L2 (ratio 5, RSUB)": (lambda **kw: ops.interp(c1, T1, c2, out2, mode=lib.INTERP_RSUB, base=T2, **kw), ident2, out2, 8 * L2["nn"]),
        "blend_L2_to_L3 (SET, two parents)": (lambda **kw: ops.interp(c2, T2, c3, out3, u2=Tp2, alpha=0.4, beta=0.6, **kw), ident3, out3, 4 * L3["nn"]),
    }
    res = {}
    for name, (fn, ident, out, nbytes) in cases.items():
        r = {}
        us = timed(lambda: fn(index_map=tuple(ident)), flush)
        r["per_target_us"], r["per_target_GBps"] = us, round(nbytes / us * 1e-3, 1)
        ref = out.clone()
        us = timed(fn, flush)
        r["march_us"], r["march_GBps"] = us, round(nbytes / us * 1e-3, 1)
        r["same_bits"] = bool(torch.equal(ref, out))
        res[name] = r
    for name, fn, nbytes in (
            ("shift_L3 (old window + L1 + L2)", lambda: ops.shift_window(c1, T1, c3, Tp3, c3n, a3, b3, mid_coords=c2, Tp_mid=Tp2), 12 * L3["nn"]),
            ("shift_L2 (old window + L1)", lambda: ops.shift_window(c1, T1, c2, Tp2, c2n, a2, b2), 12 * L2["nn"])):
        us = timed(fn, flush)
        res[name] = {"us": us, "GBps": round(nbytes / us * 1e-3, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
