"""Where a fused multi-GPU dwell sweep spends its time (run under torchrun, 2+ ranks):
K1 alone on the symmetric buffers with / without the peer stores, and the whole sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import gomelt_b200 as gm
from bench_tools.bench_l1_slab import make_slab, DT_DWELL
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
gm.load()
P = bench.host_properties()
props = gm._lib.make_props(P)
sl = make_slab(gm, props, rank, world, dev, P)
ops = gm.ops


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def k1(peers):
    nxt = 1 - sl._cur
    lo = sl._peer(rank - 1, nxt, sl._ghost_lo) if (peers and rank > 0) else None
    hi = sl._peer(rank + 1, nxt, sl._ghost_hi) if (peers and rank < world - 1) else None
    ops.level_step(props, sl.grid, sl._halves[sl._cur], sl.S1, sl._halves[nxt], DT_DWELL, nz_active=sl.nz_active,
                   n_substrate=sl.n_substrate, flags=ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX, bc5=sl.bc5,
                   z_range=(sl.zb, sl.ze), peer_lo=lo, peer_hi=hi)


plain_T = torch.empty_like(sl._halves[0]); plain_T.copy_(sl._halves[0]); plain_Tn = torch.empty_like(plain_T)


def k1_plain():
    ops.level_step(props, sl.grid, plain_T, sl.S1, plain_Tn, DT_DWELL, nz_active=sl.nz_active,
                   n_substrate=sl.n_substrate, flags=ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX, bc5=sl.bc5,
                   z_range=(sl.zb, sl.ze))


res = {"k1_plain_memory": timed(k1_plain), "k1_symm_no_peers": timed(lambda: k1(False)),
       "k1_symm_peer_stores": timed(lambda: k1(True)), "barrier_only": timed(lambda: sl._hdl.barrier(channel=0)),
       "sweep": timed(lambda: sl.dwell_sweep(DT_DWELL))}
print(f"rank {rank} TMA={os.environ.get('GOMELT_K1_TMA', '1')}: " + "  ".join(f"{k}={v:.1f}us" for k, v in res.items()), flush=True)
dist.barrier()
dist.destroy_process_group()
