"""K3 (correction projections) at C2 size, one call at a time: the marching kernel against the tile kernel for element
ratios 2 / 5 / 10 (a 500 x 500 x 40-element fine window), both modes, coefficient evaluated in the kernel as the steppers
do.  CUDA events, L2 flushed between calls.  python bench_tools/quick_k3.py [ratio ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import gomelt_b200 as gm  # noqa: E402
from gomelt_b200 import schema  # noqa: E402


def make_level(ne, bounds):
    nodes = [n + 1 for n in ne]
    coords = [np.linspace(b[0], b[1], n, dtype=np.float64).astype(np.float32) for b, n in zip(bounds, nodes)]
    h = [np.float32((b[1] - b[0]) / n) for b, n in zip(bounds, ne)]
    return {"nodes": nodes, "nn": int(np.prod(nodes)), "node_coords": coords, "h": h}


def main():
    gm.load()
    torch.cuda.set_device(0)
    cf = gm.computeFunctions
    P = schema.SetupProperties(bench.EXAMPLE_PROPS)
    props = gm._lib.make_props(P)
    ratios = [int(a) for a in sys.argv[1:]] or [2, 5, 10]
    ne, hf = (500, 500, 40), 0.02
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    out = {}
    for ratio in ratios:
        hc = hf * ratio
        fine = make_level(ne, ((4.0, 4.0 + ne[0] * hf), (4.0, 4.0 + ne[1] * hf), (-ne[2] * hf, 0.0)))
        npar = (ne[0] // ratio + 16, ne[1] // ratio + 16, ne[2] // ratio + 2)
        parent = make_level(npar, ((4.0 - 8 * hc, 4.0 - 8 * hc + npar[0] * hc), (4.0 - 8 * hc, 4.0 - 8 * hc + npar[1] * hc),
                                   (-npar[2] * hc, 0.0)))
        cells = cf._pair_cells(fine, parent)
        g = torch.Generator(device="cuda").manual_seed(1)
        nn = fine["nn"]
        Tf = 300.0 + 1500.0 * torch.rand(nn, device="cuda", generator=g)
        S1 = (torch.rand(nn, device="cuda", generator=g) > 0.4).float()
        Tp0 = 60.0 * torch.rand(nn, device="cuda", generator=g) - 30.0
        Tp1 = Tp0 + 4.0 * torch.rand(nn, device="cuda", generator=g)
        V = torch.zeros(parent["nn"], device="cuda")
        res = {"rmax": [int(v) for v in cells["rmax"]], "fine_nodes": nn}
        for how in (True, "tile"):
            for mode in (0, 1):
                ts = []
                for it in range(8):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    if mode == 0:
                        cf._project(cells, Tp0, None, V, mode=0, coef_from=(props, Tf, S1, 0), tiled=how)
                    else:
                        cf._project(cells, Tp1, None, V, mode=1, scale=1e5, A2=Tp0, coef_from=(props, Tf, S1, 0), tiled=how)
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                us = ts[len(ts) // 2]
                nbytes = nn * (12 if mode == 0 else 16)
                res[f"{'march' if how is True else 'tile'}_mode{mode}_us"] = round(us, 1)
                res[f"{'march' if how is True else 'tile'}_mode{mode}_GBps"] = round(nbytes / us * 1e-3, 1)
        out[f"ratio{ratio}"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
