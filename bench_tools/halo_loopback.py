"""Probe (one GPU): the fused halo protocol of level_step_v3 in LOOPBACK - the slab's own ghost planes stand in for
the neighbours' (peer_lo / peer_hi and the counter blocks point into this GPU's memory), so everything the protocol
adds to a sweep except the NVLink transfer itself is timed against the plain sweep of the same slab."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gomelt_b200 as gm  # noqa: E402
from bench import host_properties  # noqa: E402

NX = NY = 1001
PLANES = int(os.environ.get("GOMELT_SLAB_PLANES", "50"))
DT = 2e-3


def main():
    torch.cuda.set_device(0)
    gm.load()
    ops = gm.ops
    P = host_properties()
    props = gm._lib.make_props(P)
    nzl = PLANES + 2
    grid = gm._lib.make_grid((NX, NY, nzl), (0.2, 0.2, 0.2))
    plane = NX * NY
    n = plane * nzl
    bufs = [torch.full((n,), 400.0, device="cuda") + 5 * torch.rand(n, device="cuda") for _ in range(2)]
    S1 = torch.ones(n, device="cuda")
    nsync = int(gm._lib.load().gomelt_halo_sync_words())
    sync = torch.zeros(nsync, dtype=torch.int32, device="cuda")
    bc5 = [P["T_amb"]] * 5
    flags = ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX

    def run(halo, K=20, W=5):
        sync.zero_()
        seq, cur = 0, 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(W + K):
            if it == W:
                torch.cuda.synchronize()
                ev[0].record()
            nxt = (cur + 1) % 2
            kw = {}
            if halo:
                base = bufs[nxt].data_ptr()
                # loopback: my first owned plane -> my own upper ghost (plane nzl-1), my last owned plane -> plane 0
                kw = dict(peer_lo=base + 4 * plane * (nzl - 1), peer_hi=base,
                          halo=(sync.data_ptr(), sync.data_ptr(), sync.data_ptr(), seq))
                seq += 1
            ops.level_step(props, grid, bufs[cur], S1, bufs[nxt], DT, nz_active=nzl, flags=flags, bc5=bc5,
                           z_range=(1, nzl - 1), **kw)
            cur = nxt
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) * 1e3 / K

    out = {"planes": PLANES, "exp": os.environ.get("GOMELT_K1_EXP", "0"), "plain_us": run(False),
           "halo_loopback_us": run(True), "plain_again_us": run(False)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
