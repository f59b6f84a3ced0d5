"""Scratch timing of K1 at the 10 M-node Level-3 size (not the bench contract; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gomelt_b200 as gm

P = gm.schema.SetupProperties({"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15,
                               "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3,
                               "latent_heat_evap": 6457000.0})
ops = gm.ops
props = gm._lib.make_props(P)
nx, ny, nz = 513, 513, 39
grid = gm._lib.make_grid((nx, ny, nz), (0.02, 0.02, 0.02))
nn = nx * ny * nz
g = torch.Generator(device="cuda").manual_seed(0)
T0 = 300 + 2000 * torch.rand(nn, device="cuda", generator=g)
S1 = (torch.rand(nn, device="cuda", generator=g) > 0.5).float()
Tout = torch.empty_like(T0); S1o = torch.empty_like(T0)
flush = torch.empty(64 * 1024 * 1024, device="cuda")
tx = torch.rand(nx, device="cuda", generator=g); ty = torch.rand(ny, device="cuda", generator=g); tz = torch.rand(nz, device="cuda", generator=g)
top = torch.zeros(nx * ny, device="cuda")
print("variant", os.environ.get("GOMELT_K1_VARIANT", "2"), "generic", os.environ.get("GOMELT_K1_GENERIC", "0"))
for zc in [int(a) for a in sys.argv[1:]] or (0, 20, 13, 10):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.level_step(props, grid, T0, S1, Tout, 1e-5, src=(tx, ty, tz, 1e-3), n_substrate=0,
                       flags=ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX, S1_out=S1o, z_chunk=zc)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts[3:]) * 1e-3
    print(f"z_chunk={zc:3d}: {t*1e6:8.1f} us  {nn/t/1e9:7.2f} G DOF/s  {nn*16/t/1e9:7.1f} GB/s algorithmic ({nn*16/t/6542.1e9*100:.1f}% of measured HBM peak)")
import hashlib
torch.cuda.synchronize()
print("out sha1", hashlib.sha1(Tout.cpu().numpy().tobytes()).hexdigest()[:16], hashlib.sha1(S1o.cpu().numpy().tobytes()).hexdigest()[:16])
if os.environ.get("GOMELT_K1_DUMP"):
    np.save(os.environ["GOMELT_K1_DUMP"], Tout.cpu().numpy())
