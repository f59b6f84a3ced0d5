"""Timing of the subcycleL3_Part2 call shape (SKIP_FACES + WRITE_S2 + ACCUM, everything in place) at the 10 M-node
Level-3 size: fast kernel vs general kernel, for scattered and clustered molten nodes."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gomelt_b200 as gm
ops = gm.ops
P = gm.schema.SetupProperties({"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15, "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3, "latent_heat_evap": 6457000.0})
props = gm._lib.make_props(P)
nx, ny, nz = 513, 513, 39
grid = gm._lib.make_grid((nx, ny, nz), (0.02, 0.02, 0.02)); nn = nx*ny*nz
g = torch.Generator(device="cuda").manual_seed(0)
xx = torch.linspace(-1, 1, nx, device="cuda")[None, None, :]; yy = torch.linspace(-1, 1, ny, device="cuda")[None, :, None]
zz = torch.linspace(-1, 0, nz, device="cuda")[:, None, None]
pool = (300.0 + 2600.0 * torch.exp(-60.0 * (xx * xx + yy * yy) + 4.0 * zz) + 0 * zz).reshape(-1).contiguous()
for label, hot in (("40% of the nodes molten, scattered", 2000.0), ("0.8% molten, scattered", 1320.0), ("one melt pool (clustered)", None), ("nothing molten", 1000.0)):
    T0 = (300 + hot * torch.rand(nn, device="cuda", generator=g)) if hot else pool.clone()
    S1 = (torch.rand(nn, device="cuda", generator=g) > 0.5).float()
    print(label, float((T0 >= 1609).float().mean()))
    Tout = torch.empty_like(T0); S2 = torch.zeros(nn, device="cuda", dtype=torch.uint8); acc = torch.zeros(nn, device="cuda"); mx = torch.zeros(nn, device="cuda")
    tx = torch.rand(nx, device="cuda"); ty = torch.rand(ny, device="cuda"); tz = torch.rand(nz, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    queue = torch.zeros(2 + 2 * (nn // 120 + 1024), device="cuda", dtype=torch.int32)
    for name, extra, q in (("fast", 0, None), ("fast + hot-plane queue", 0, queue), ("general", ops.STEP_GENERAL_KERNEL, None)):
        ts = []
        for it in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.level_step(props, grid, T0, S1, Tout, 1e-5, src=(tx, ty, tz, 1e-3), n_substrate=2*nx*ny, S1_out=S1, S2_out=S2, S2_prev=S2, accum=acc, max_accum=mx,
                           flags=ops.STEP_SKIP_FACES | ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_WRITE_S2 | ops.STEP_ACCUM | ops.STEP_FUSED_FLUX | extra, bk_queue=q)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = np.median(ts[3:]) * 1e-3
        print(f"  Part-2 shape {name}: {t*1e6:.1f} us  ({nn*33/t/1e9:.0f} GB/s at the 33 B/DOF of SURVEY 8d)")
