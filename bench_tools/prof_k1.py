"""Launch K1 a few times at the 10 M-node Level-3 size (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gomelt_b200 as gm
P = gm.schema.SetupProperties({"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15,
                               "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3,
                               "latent_heat_evap": 6457000.0})
ops = gm.ops
props = gm._lib.make_props(P)
nx, ny, nz = 513, 513, 39
zc = int(sys.argv[1]) if len(sys.argv) > 1 else 0
grid = gm._lib.make_grid((nx, ny, nz), (0.02, 0.02, 0.02))
nn = nx * ny * nz
g = torch.Generator(device="cuda").manual_seed(0)
T0 = 300 + 2000 * torch.rand(nn, device="cuda", generator=g)
S1 = (torch.rand(nn, device="cuda", generator=g) > 0.5).float()
Tout = torch.empty_like(T0); S1o = torch.empty_like(T0)
tx = torch.rand(nx, device="cuda", generator=g); ty = torch.rand(ny, device="cuda", generator=g); tz = torch.rand(nz, device="cuda", generator=g)
top = torch.zeros(nx * ny, device="cuda")
for it in range(4):
    ops.level_step(props, grid, T0, S1, Tout, 1e-5, src=(tx, ty, tz, 1e-3), n_substrate=0,
                   flags=ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX, S1_out=S1o, z_chunk=zc)
torch.cuda.synchronize()
