"""Level-1 z-slabs on real GPUs (needs >= 2 devices; skipped on a 1-GPU box; uses EVERY device of the box): the
device-side halo protocol (the step, then one exchange kernel: peer stores + release / acquire counters; no barrier
launch, no NCCL), the round-1 form (push kernel + barrier launch) and the NCCL send/recv path all reproduce the
single-GPU sweeps bit for bit - in dwell mode and in the load-vector + clamp shape of the stepGOMELT / subcycleGOMELT
Level-1 sweeps.  The driver-visible copy of this check is the ``parity_check`` of the multi-GPU bench line."""
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
NSWEEPS = 8
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


def _rhs():
    rng = np.random.default_rng(6)
    nx, ny, nz = NODES
    return (2e-4 * rng.standard_normal(nx * ny * nz)).astype(np.float32)


MODES = {"fused": dict(symmetric=True, fused=True), "push_barrier": dict(symmetric=True, fused=False),
         "nccl": dict(symmetric=False)}


def _worker(rank, world, port, out, mode, shape):
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
                                device=dev, **MODES[mode])
        assert sl.symmetric == MODES[mode]["symmetric"] and sl.fused == (mode == "fused")
        pl = nx * ny
        sl.set_owned(torch.as_tensor(T0[sl.k0 * pl:sl.k1 * pl]).to(dev), torch.as_tensor(S1[sl.k0 * pl:sl.k1 * pl]).to(dev))
        rhs = None
        if shape == "rhs":
            rhs = torch.as_tensor(_rhs()[sl.g0 * pl:(sl.g0 + sl.nzl) * pl]).to(dev)
        l0 = gm.ops.LAUNCHES
        for _ in range(NSWEEPS):
            sl.sweep(DT, rhs=rhs, clamp=(shape == "rhs"))
        if mode == "fused":
            assert gm.ops.LAUNCHES - l0 == 2 * NSWEEPS, "step + halo exchange: two launches per sweep, nothing else"
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


@pytest.mark.parametrize("shape", ["dwell", "rhs"])
@pytest.mark.parametrize("mode", list(MODES))
def test_slabs_match_single_gpu_bitwise(tmp_path, gm, mode, shape):
    import torch
    import torch.multiprocessing as mp

    world = torch.cuda.device_count()
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
    rhs = torch.as_tensor(_rhs()).cuda() if shape == "rhs" else None
    for _ in range(NSWEEPS):
        ref.sweep(DT, rhs=rhs, clamp=(shape == "rhs"))
    want = ref.T.cpu().numpy()
    for attempt in range(3):  # (a free port can be taken between the probe and the rendezvous)
        try:
            mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path), mode, shape), nprocs=world, join=True,
                               start_method="spawn")
            break
        except Exception as exc:
            if "EADDRINUSE" not in str(exc) or attempt == 2:
                raise
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert got.shape == want.shape
    assert np.isfinite(got).all() and np.abs(got - T0).max() > 1.0  # the sweeps did something
    assert np.array_equal(got, want), int((got != want).sum())
