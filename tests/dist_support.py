"""Worker of the distributed-driver tests: one rank of a slab-decomposed run of gomelt_b200/driver.py (every rank runs
the same loop; the laser owner writes its final fields to ``out``)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_rank(rank, world, port, out, case, backend, one_device):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import driver_support

    import gomelt_b200 as gm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if one_device else rank
    torch.cuda.set_device(dev)
    kw = {"device_id": torch.device("cuda", dev)} if backend == "nccl" else {}
    dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    try:
        gm.load()
        cf = gm.computeFunctions
        tmp = os.path.join(out, f"rank{rank}")
        os.makedirs(tmp, exist_ok=True)
        inp = getattr(driver_support, case)(tmp)
        cf.enable_distributed(rank, world)
        res = gm.driver.go_melt(inp, write_final=False)
        torch.cuda.synchronize()
        if not res.get("worker"):
            L = res["Levels"]
            d = cf.distOf(L)
            np.savez(os.path.join(out, "owner.npz"), L1T=L[1]["T0"].cpu().numpy(), L2T=L[2]["T0"].cpu().numpy(),
                     L3T=L[3]["T0"].cpu().numpy(), accum=res["accum_time"].cpu().numpy(), owner=rank,
                     boxes_down=d.stats["boxes_down"], boxes_up=d.stats["boxes_up"], solves=d.stats["solves"],
                     counts=np.array([res["counts"][k] for k in sorted(res["counts"])]))
        S1 = cf.gatherL1(res["Levels"], "S1")
        if S1 is not None:
            np.save(os.path.join(out, "owner_S1.npy"), S1.cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()
