"""N>1 host logic of the Level-1 z-slab decomposition on CPU (gloo, world_size 2 and 3): plane
partition, ghost exchange, and - with the oracle standing in for the CUDA kernel - that slab-wise
dwell sweeps reproduce the single-domain result (stepGOMELTDwellTime cF:2617-2664)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROPS_IN = {"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15,
            "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3,
            "latent_heat_evap": 6457000.0}
ELEMENTS = (9, 7, 11)
BOUNDS = ((0.0, 1.8), (0.0, 1.4), (-2.0, 0.2))
NZ_ACTIVE = 10
COND = {"x": [301.0, 302.0], "y": [303.0, 304.0], "z": [305.0, 306.0]}
NSWEEPS = 3
DT = 2e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleOps:
    """Test double for gomelt_b200.ops on CPU tensors: the oracle's dwell step on the local slab."""
    STEP_BC_CONST = 0x08
    STEP_FUSED_FLUX = 0x40
    STEP_CLAMP = 0x01
    LAUNCHES = 0

    def __init__(self, cF, P, make_level, h):
        self.cF, self.P, self.make_level, self.h = cF, P, make_level, h

    def surface_flux(self, props, grid, T0, flux, nz_active=None, add=False):
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        lv = self._level(nx, ny, nz)
        ne = (nx - 1) * (ny - 1) * (nz_active - 1)
        F = self.cF.computeConvRadBC(lv, T0.numpy(), ne, lv["nn"], self.P, 0)
        flux.copy_(torch.from_numpy(F[(nz_active - 1) * nx * ny:nz_active * nx * ny].copy()))
        return flux

    def _level(self, nx, ny, nz):
        hx, hy, hz = self.h
        return self.make_level((nx - 1, ny - 1, nz - 1), ((0, hx * (nx - 1)), (0, hy * (ny - 1)), (0, hz * (nz - 1))))

    def level_step(self, props, grid, T0, S1, T_out, dt, *, topflux=None, nz_active=None, n_substrate=0,
                   flags=0, bc5=None, z_range=None, **kw):
        cF = self.cF
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        lv = self._level(nx, ny, nz)
        P_ = nx * ny
        nn = lv["nn"]
        T = T0.numpy()
        _, _, k, rc = cF.computeStateProperties(T, S1.numpy(), self.P, n_substrate)
        F = np.zeros(nn, np.float32)
        if topflux is not None:
            F[(nz_active - 1) * P_:nz_active * P_] = topflux.numpy()
        ne = (nx - 1) * (ny - 1) * max(nz_active - 1, 0)
        if (flags & self.STEP_FUSED_FLUX) and nz_active >= 2:  # computeConvRadBC inside the step
            F = cF.computeConvRadBC(lv, T, ne, nn, self.P, F)
        rhs = kw.get("rhs")
        Tn = cF.solveMatrixFreeFE(lv, nn, ne, k, rc, dt, T, F, 0 if rhs is None else rhs.numpy())
        if flags & self.STEP_CLAMP:
            Tn = np.maximum(Tn, np.float32(self.P["T_amb"]))
        Tn[nz_active * P_:] = np.float32(self.P["T_amb"])
        T3 = Tn.reshape(nz, ny, nx)
        T3[:, 0, :] = bc5[0]; T3[:, -1, :] = bc5[1]; T3[:, :, 0] = bc5[2]; T3[:, :, -1] = bc5[3]
        z0, z1 = z_range
        if z0 == 0:
            T3[0] = bc5[4]
        T_out[z0 * P_:z1 * P_] = torch.from_numpy(Tn[z0 * P_:z1 * P_].copy())
        return T_out


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib

        from oracle import computeFunctions as cF
        from oracle.util import make_level, smooth_field

        slab = importlib.import_module("gomelt_b200.slab")
        lib = importlib.import_module("gomelt_b200._lib")
        P = cF.SetupProperties(PROPS_IN)
        lv = make_level(ELEMENTS, BOUNDS)
        nx, ny, nz = lv["nodes"]
        rng = np.random.default_rng(0)
        T0 = smooth_field(lv, rng)
        S1 = (rng.random(lv["nn"]) > 0.4).astype(np.float32)
        nsub = 3 * nx * ny

        class _GM:
            pass

        gm = _GM()
        gm._lib = lib
        gm.ops = _OracleOps(cF, P, make_level, lv["h"])
        bc5 = [COND["y"][0], COND["y"][1], COND["x"][0], COND["x"][1], COND["z"][0]]
        sl = slab.Level1Slab(gm, None, lv["nodes"], lv["h"], rank, world, bc5, nz_active=NZ_ACTIVE,
                             n_substrate=nsub, device=None)
        sl.comm = None
        pl = nx * ny
        sl.set_owned(torch.from_numpy(T0[sl.k0 * pl:sl.k1 * pl].copy()), torch.from_numpy(S1[sl.k0 * pl:sl.k1 * pl].copy()))
        # ghosts of T and S1 now hold the neighbours' planes
        if rank > 0:
            assert np.array_equal(sl.T[:pl].numpy(), T0[(sl.k0 - 1) * pl:sl.k0 * pl])
        for _ in range(NSWEEPS):
            _cpu_sweep(sl, DT)
        got = sl.owned(sl.T).numpy().copy()
        np.save(os.path.join(out, f"rank{rank}.npy"), got)
    finally:
        dist.destroy_process_group()


def _cpu_sweep(sl, dt):
    """Level1Slab.dwell_sweep without CUDA streams (same call sequence)."""
    slab_mod = sys.modules[type(sl).__module__]
    top = None  # the surface load is evaluated inside the level step (GOMELT_STEP_FUSED_FLUX)
    zb, ze = sl.zb, sl.ze
    lo = zb + 1 if sl.rank > 0 else zb
    hi = max(ze - 1 if sl.rank < sl.world - 1 else ze, lo)
    for a, b in ((zb, lo), (hi, ze)):
        if b > a:
            sl._k1(dt, a, b, top)
    works = slab_mod.exchange_planes(sl.Tn, sl.plane, zb, ze, sl.rank, sl.world)
    if hi > lo:
        sl._k1(dt, lo, hi, top)
    for w in works:
        w.wait()
    sl.T, sl.Tn = sl.Tn, sl.T


def test_partition_planes():
    import importlib

    slab = importlib.import_module("gomelt_b200.slab")
    assert slab.partition_planes(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert slab.partition_planes(8, 8) == [(i, i + 1) for i in range(8)]
    with pytest.raises(ValueError):
        slab.partition_planes(2, 3)
    assert slab.local_extent(0, 3, 0, 4) == (0, 5, 0, 4)
    assert slab.local_extent(1, 3, 4, 7) == (3, 5, 1, 4)
    assert slab.local_extent(2, 3, 7, 10) == (6, 4, 1, 4)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_sweeps_match_single_domain(tmp_path, world):
    from oracle import computeFunctions as cF
    from oracle.util import make_level, smooth_field

    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="fork")
    P = cF.SetupProperties(PROPS_IN)
    lv = make_level(ELEMENTS, BOUNDS)
    nx, ny, nz = lv["nodes"]
    rng = np.random.default_rng(0)
    T0 = smooth_field(lv, rng)
    S1 = (rng.random(lv["nn"]) > 0.4).astype(np.float32)
    Levels = [None, dict(lv, T0=T0, S1=S1, conditions=COND)]
    tmp = (ELEMENTS[0] * ELEMENTS[1] * (NZ_ACTIVE - 1), nx * ny * NZ_ACTIVE)
    for _ in range(NSWEEPS):
        Levels = cF.stepGOMELTDwellTime(Levels, tmp, (0, 0, lv["nn"]), P, DT, (0, 3 * nx * ny))
    ref = Levels[1]["T0"]
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert got.shape == ref.shape
    # the slab-local oracle sums element contributions in a different order at slab edges: f32 round-off only
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 2e-6


def test_active_plane_partition_balances_the_powder_bed():
    """dist.py cuts the slabs over the planes that are active when the build starts; the planes above the powder bed
    (a store per node, no stencil) ride on the last rank."""
    from gomelt_b200.slab import partition_active_planes, partition_planes

    for nz, act, world in ((156, 141, 8), (31, 22, 4), (7, 5, 3), (40, 40, 4), (9, 2, 4)):
        parts = partition_active_planes(nz, act, world)
        assert parts[0][0] == 0 and parts[-1][1] == nz and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(b > a for a, b in parts)
        active = [max(0, min(b, max(act, world)) - a) for a, b in parts]
        assert max(active) - min(active) <= 1   # the active planes are balanced
    assert partition_active_planes(40, 40, 4) == partition_planes(40, 4)
