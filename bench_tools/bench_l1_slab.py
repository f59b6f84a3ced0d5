"""Multi-GPU arm of bench.py: Level-1 z-slab dwell sweeps with NCCL halo exchange (BASELINE.json
configs[4]).  Weak scaling: every rank owns SLAB_PLANES planes of a 1001 x 1001 grid
(25.05 M nodes per GPU; 8 GPUs = 200 planes = 200.4 M nodes).  One *step* = one explicit sweep
of the whole Level-1 grid (stepGOMELTDwellTime cF:2617-2664): surface flux on the top plane,
fused level step, one-plane halo exchange of T with both z-neighbours."""
import json
import os
import time

L1_NX, L1_NY = 1001, 1001
SLAB_PLANES = 25
L1_H = (0.2, 0.2, 0.2)
DT_DWELL = 2e-3
B_ALG_L1 = 12


def make_slab(gm, props, rank, world, device, P):
    import numpy as np
    import torch

    slab_mod = gm.slab
    nz = SLAB_PLANES * world
    bc5 = [P["T_amb"]] * 5
    sl = slab_mod.Level1Slab(gm, props, (L1_NX, L1_NY, nz), L1_H, rank, world, bc5, device=device,
                             symmetric=os.environ.get("GOMELT_SLAB_NCCL", "0") != "1")
    # synthetic state: T_amb + smooth warm region decaying with depth (global z), bulk everywhere
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    nown = (sl.k1 - sl.k0) * sl.plane
    zg = torch.arange(sl.k0, sl.k1, device=device, dtype=torch.float32).repeat_interleave(sl.plane)
    T = P["T_amb"] + 600.0 * torch.exp(-(nz - 1 - zg) / 40.0) + 5.0 * torch.rand(nown, device=device, generator=g)
    S1 = torch.ones(nown, device=device)
    sl.set_owned(T, S1)
    return sl


def run_gomelt_multi(args, read_peaks, ClockSampler, host_properties, single_gpu=False):
    import torch
    import torch.distributed as dist

    import gomelt_b200 as gm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        # keep stdout to the ONE JSON line: the NCCL communicator set-up prints "NCCL version ..." on fd 1, so
        # fd 1 points at stderr while the process group (and its first collective) is created
        import sys

        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    gm.load()
    P = host_properties()
    props = gm._lib.make_props(P)
    K, W = args.steps, max(args.warmup, 3)
    sl = make_slab(gm, props, rank, world, device, P)
    nn_total = L1_NX * L1_NY * SLAB_PLANES * world
    for _ in range(W):
        sl.dwell_sweep(DT_DWELL)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = gm.ops.LAUNCHES
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sl.dwell_sweep(DT_DWELL)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_s = float(ms.item()) * 1e-3
    launches = gm.ops.LAUNCHES - l0
    # end to end: host buffers of the owned planes up, K sweeps, owned planes down (per step = per sweep)
    nown = (sl.k1 - sl.k0) * sl.plane
    hT = torch.empty(nown, dtype=torch.float32).pin_memory()
    hT.copy_(sl.owned(sl.T))
    oT = torch.empty(nown, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te0 = time.perf_counter()
    Ke = max(1, min(K, 10))
    for _ in range(Ke):
        sl.owned(sl.T).copy_(hT, non_blocking=True)
        if world > 1:
            for w in gm.slab.exchange_planes(sl.T, sl.plane, sl.zb, sl.ze, rank, world):
                w.wait()
        if sl.symmetric:
            sl._hdl.barrier(channel=0)  # neighbours' ghosts in place before the sweep reads them
        T = sl.dwell_sweep(DT_DWELL)
        oT.copy_(sl.owned(T), non_blocking=True)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te = torch.tensor([time.perf_counter() - te0], device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    # same-workload single-GPU reference (one slab of the same size, no neighbours), timed on rank 0 outside
    # the timed region, so that a weak-scaling efficiency can be read off this one line
    one_gpu = None
    if world > 1:
        if rank == 0:
            s1 = gm.slab.Level1Slab(gm, props, (L1_NX, L1_NY, SLAB_PLANES), L1_H, 0, 1, [P["T_amb"]] * 5, device=device)
            s1.set_owned(sl.owned(sl.T)[: s1.plane * SLAB_PLANES].clone(), torch.ones(s1.plane * SLAB_PLANES, device=device))
            for _ in range(W):
                s1.dwell_sweep(DT_DWELL)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(K):
                s1.dwell_sweep(DT_DWELL)
            f1.record()
            torch.cuda.synchronize()
            one_gpu = f0.elapsed_time(f1) * 1e-3 / K
            del s1
        dist.barrier()
    if rank == 0:
        peaks = read_peaks()
        value = K * nn_total / total_s
        achieved = B_ALG_L1 * (nn_total / world) * K / total_s / 1e9  # per GPU
        line = {
            "metric": "Level-1 DOF-updates/s", "value": value, "unit": "DOF-updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * total_s / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"L1-slab: {L1_NX}x{L1_NY}x{SLAB_PLANES * world}-node Level-1 grid "
                                   f"({nn_total} nodes), z-slabs of {SLAB_PLANES} planes per GPU, dwell sweeps "
                                   "(fused level step incl. surface flux; the boundary planes are stored into the "
                                   "neighbours' ghost planes by the same kernel over NVLink peer memory, one "
                                   "device-side barrier per sweep)" if sl.symmetric else
                                   "(fused level step incl. surface flux + one-plane T halo exchange by NCCL "
                                   "send/recv per sweep)",
                       "nodes": nn_total, "parallelism": f"z-slab x{world}",
                       "l2": "working set per GPU 300 MB > 126 MB L2 (inputs larger than L2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "level_step_v3",
                         "bytes_per_dof": B_ALG_L1, "peak_source": peaks["source"],
                         "note": "per GPU, whole sweep (halo exchange included)"},
            "e2e": {"value": Ke * nn_total / e2e_s, "unit": "DOF-updates/s",
                    "h2d_bytes_per_step": 4 * nown * world, "d2h_bytes_per_step": 4 * nown * world,
                    "api": "host-buffer sweep: upload owned T planes, halo fill, one dwell sweep, download"},
            "gpu_launches": launches, "clocks": clocks,
            "same_workload_1gpu": None if one_gpu is None else {
                "ms_per_step": 1e3 * one_gpu, "value": L1_NX * L1_NY * SLAB_PLANES / one_gpu,
                "weak_scaling_efficiency": one_gpu / (total_s / K),
                "note": "one 25-plane slab on rank 0, no neighbours, timed right after the N-GPU region"},
            "halo_bytes_per_step_per_gpu": 4 * sl.plane * ((1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)),
            "halo": ("in-kernel peer stores (symmetric memory)" if sl.symmetric else "NCCL send/recv") if world > 1
                    else "none",
        }
        if single_gpu:
            return line
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return None
