"""Level-1 z-slabs on real GPUs (needs >= 2 devices; skipped on a 1-GPU box): the fused halo path (K1 stores
its boundary planes into the neighbours' ghost planes in peer-mapped symmetric memory) and the NCCL
send/recv path both reproduce the single-GPU sweeps bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

NODES = (131, 67, 23)
H = (0.2, 0.2, 0.2)
NZ_ACTIVE = 20
NSWEEPS = 4
DT = 2e-3
PROPS_IN = {"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15,
            "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3,
            "latent_heat_evap": 6457000.0}
BC5 = [301.0, 302.0, 303.0, 304.0, 305.0]


def _state():
    rng = np.random.default_rng(5)
    nx, ny, nz = NODES
    z = np.arange(nz, dtype=np.float32)[:, None, None]
    T = 400.0 + 900.0 * np.exp(-(nz - 1 - z) / 6.0) + 30.0 * rng.random((nz, ny, nx))
    S1 = (rng.random((nz, ny, nx)) > 0.3).astype(np.float32)
    return T.astype(np.float32).reshape(-1), S1.reshape(-1)


def _worker(rank, world, port, out, symmetric):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import gomelt_b200 as gm

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        gm.load()
        P = gm.schema.SetupProperties(PROPS_IN)
        props = gm._lib.make_props(P)
        T0, S1 = _state()
        nx, ny, nz = NODES
        sl = gm.slab.Level1Slab(gm, props, NODES, H, rank, world, BC5, nz_active=NZ_ACTIVE, n_substrate=3 * nx * ny,
                                device=dev, symmetric=symmetric)
        assert sl.symmetric == symmetric
        pl = nx * ny
        sl.set_owned(torch.as_tensor(T0[sl.k0 * pl:sl.k1 * pl]).to(dev), torch.as_tensor(S1[sl.k0 * pl:sl.k1 * pl]).to(dev))
        if symmetric:
            sl._hdl.barrier(channel=0)
        for _ in range(NSWEEPS):
            sl.dwell_sweep(DT)
        torch.cuda.synchronize()
        np.save(os.path.join(out, f"rank{rank}.npy"), sl.owned(sl.T).cpu().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("symmetric", [True, False])
def test_slabs_match_single_gpu_bitwise(tmp_path, gm, symmetric):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    # single-GPU reference: the same sweeps on the whole grid
    P = gm.schema.SetupProperties(PROPS_IN)
    props = gm._lib.make_props(P)
    T0, S1 = _state()
    nx, ny, nz = NODES
    ref = gm.slab.Level1Slab(gm, props, NODES, H, 0, 1, BC5, nz_active=NZ_ACTIVE, n_substrate=3 * nx * ny,
                             device=torch.device("cuda", 0))
    ref.set_owned(torch.as_tensor(T0).cuda(), torch.as_tensor(S1).cuda())
    for _ in range(NSWEEPS):
        ref.dwell_sweep(DT)
    want = ref.T.cpu().numpy()
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path), symmetric), nprocs=world, join=True,
                       start_method="spawn")
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert got.shape == want.shape
    assert np.isfinite(got).all() and np.abs(got - T0).max() > 1.0  # the sweeps did something
    assert np.array_equal(got, want), int((got != want).sum())
