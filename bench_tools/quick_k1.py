"""Scratch timing of K1 at the 10 M-node Level-3 size (not the bench contract; see bench.py).

    python bench_tools/quick_k1.py [shape ...]     shapes: nat (natural-boundary bench shape, general kernel),
        sub (subcycle L3 substep shape: SKIP_FACES, fast kernel), subg (same call on the general kernel),
        l2 (rhs shape), dwell (Level-1 dwell shape on a 1001 x 1001 x 25 slab)
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gomelt_b200 as gm

P = gm.schema.SetupProperties({"laser_radius": 0.1, "laser_depth": 0.1, "laser_absorptivity": 0.45, "T_amb": 298.15,
                               "T_solidus": 1533, "T_liquidus": 1609, "h_conv": 1.5e-05, "emissivity": 0.3,
                               "latent_heat_evap": 6457000.0})
ops = gm.ops
props = gm._lib.make_props(P)
PEAK = 6542.1e9
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def run(shape, zc=0, mods=()):
    nx, ny, nz = (1001, 1001, 25) if shape.startswith("dwell") else (513, 513, 39)
    grid = gm._lib.make_grid((nx, ny, nz), (0.02, 0.02, 0.02))
    nn = nx * ny * nz
    g = torch.Generator(device="cuda").manual_seed(0)
    hot = os.environ.get("GOMELT_QUICK_HOT", "1") == "1" and "cold" not in mods
    extra = ops.STEP_NO_COLD_PLANES if "nocold" in mods else 0
    T0 = (300 + 2000 * torch.rand(nn, device="cuda", generator=g)) if hot else (300 + 900 * torch.rand(nn, device="cuda", generator=g))
    S1 = (torch.rand(nn, device="cuda", generator=g) > 0.5).float()
    Tout = torch.empty_like(T0); S1o = torch.empty_like(T0)
    tx = torch.rand(nx, device="cuda", generator=g); ty = torch.rand(ny, device="cuda", generator=g); tz = torch.rand(nz, device="cuda", generator=g)
    rhs = 1e-4 * torch.randn(nn, device="cuda", generator=g)
    kw, bpd = {}, 16
    if shape == "nat":
        kw = dict(src=(tx, ty, tz, 1e-3), flags=ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX, S1_out=S1o)
    elif shape in ("sub", "subg", "subi"):  # subi: S1 updated in place (what the steppers do)
        kw = dict(src=(tx, ty, tz, 1e-3), S1_out=(S1 if shape == "subi" else S1o), n_substrate=2 * nx * ny,
                  flags=ops.STEP_CLAMP | ops.STEP_WRITE_S1 | ops.STEP_FUSED_FLUX | ops.STEP_SKIP_FACES |
                  (ops.STEP_GENERAL_KERNEL if shape == "subg" else 0))
    elif shape in ("l2", "l2g"):
        kw = dict(rhs=rhs, n_substrate=2 * nx * ny, flags=ops.STEP_CLAMP | ops.STEP_FUSED_FLUX | ops.STEP_SKIP_FACES |
                  (ops.STEP_GENERAL_KERNEL if shape == "l2g" else 0))
        bpd = 16
    elif shape in ("dwell", "dwellg"):
        kw = dict(flags=ops.STEP_BC_CONST | ops.STEP_FUSED_FLUX | (ops.STEP_GENERAL_KERNEL if shape == "dwellg" else 0),
                  bc5=[298.15] * 5)
        bpd = 12
    ts = []
    for it in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.level_step(props, grid, T0, S1, Tout, 1e-5, z_chunk=zc, **dict(kw, flags=kw["flags"] | extra))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts[3:]) * 1e-3
    print(f"{shape:6s} {','.join(mods):12s} z_chunk={zc:3d}: {t*1e6:8.1f} us  {nn/t/1e9:7.2f} G DOF/s  {nn*bpd/t/1e9:7.1f} GB/s algorithmic "
          f"({nn*bpd/t/PEAK*100:.1f}% of measured HBM peak)", flush=True)


for a in sys.argv[1:] or ["nat", "sub", "subg"]:
    a, _, mods = a.partition("@")   # e.g. subi@cold  subi@cold,nocold  (cold: field below the solidus)
    shape, _, zc = a.partition(":")
    run(shape, int(zc or 0), tuple(m for m in mods.split(",") if m))
