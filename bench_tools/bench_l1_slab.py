"""Multi-GPU arm of bench.py: Level-1 z-slab dwell sweeps with the fused halo exchange (BASELINE.json configs[4];
SURVEY.md 8e).  Weak scaling: every rank owns SLAB_PLANES = 100 planes of a 1001 x 1001 grid (100.2 M nodes and
1.6 GB of fields per GPU - shards sized for the 180 GB of a B200, so that the two ghost planes and the exchange are
2-3 % of a sweep; 2 GPUs = 200.4 M, 8 GPUs = 801.6 M nodes).  The line also carries the 50- and 25-plane figures.  One *step* = one explicit sweep of the whole Level-1 grid
(stepGOMELTDwellTime cF:2617-2664): surface flux on the top plane, fused level step, Dirichlet faces, one-plane halo
exchange of T with both z-neighbours - ONE C-ABI call = two launches per rank and sweep (gomelt_abi.h, halo_sync).

After the timed region the line gets a ``parity_check``: the same K sweeps are repeated from the deterministic
initial state, and rank 0 recomputes, alone and without any halo exchange, a cut of planes around every slab
boundary (2 (K + 3) planes: the stencil has radius 1, so after K sweeps the planes K + 1 away from the cut's own
edges are exact); the two planes on either side of every boundary must agree with the N-GPU result bit for bit.
A mismatch makes the run exit non-zero.
"""
import json
import os
import sys
import time

L1_NX, L1_NY = 1001, 1001
SLAB_PLANES = int(os.environ.get("GOMELT_SLAB_PLANES", "100"))
L1_H = (0.2, 0.2, 0.2)
DT_DWELL = 2e-3
B_ALG_L1 = 12
PARITY_SWEEPS = 5


def initial_T(torch, P, nz, g0, g1, device):
    """Deterministic synthetic temperature of global planes [g0, g1): T_amb + a warm region decaying with depth +
    a +-2.5 K pattern that is a pure function of the global node index (so any rank can rebuild any plane)."""
    plane = L1_NX * L1_NY
    zg = torch.arange(g0, g1, device=device, dtype=torch.float32).repeat_interleave(plane)
    idx = torch.arange(g0 * plane, g1 * plane, device=device, dtype=torch.int64)
    noise = ((idx * 2654435761) % 1000003).to(torch.float32) * (5.0 / 1000003.0)
    return P["T_amb"] + 600.0 * torch.exp(-(nz - 1 - zg) / 40.0) + noise


def make_slab(gm, props, rank, world, device, P, planes=None, fused=True):
    import torch

    planes = SLAB_PLANES if planes is None else planes
    nz = planes * world
    bc5 = [P["T_amb"]] * 5
    sl = gm.slab.Level1Slab(gm, props, (L1_NX, L1_NY, nz), L1_H, rank, world, bc5, device=device,
                            symmetric=os.environ.get("GOMELT_SLAB_NCCL", "0") != "1", fused=fused)
    reset_slab(sl, P, device)
    return sl


def reset_slab(sl, P, device):
    import torch

    nz = sl.nodes_global[2]
    sl.set_owned(initial_T(torch, P, nz, sl.k0, sl.k1, device), torch.ones((sl.k1 - sl.k0) * sl.plane, device=device))


def timed_sweeps(torch, dist, sl, K, W, world, device):
    for _ in range(W):
        sl.dwell_sweep(DT_DWELL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        sl.dwell_sweep(DT_DWELL)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) * 1e-3


def one_gpu_reference(torch, gm, props, P, planes, K, W, device):
    """One slab of the same size alone on this GPU (no neighbours): the weak-scaling denominator."""
    s1 = gm.slab.Level1Slab(gm, props, (L1_NX, L1_NY, planes), L1_H, 0, 1, [P["T_amb"]] * 5, device=device)
    s1.set_owned(initial_T(torch, P, planes, 0, planes, device), torch.ones(s1.plane * planes, device=device))
    for _ in range(W):
        s1.dwell_sweep(DT_DWELL)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        s1.dwell_sweep(DT_DWELL)
    f1.record()
    torch.cuda.synchronize()
    t = f0.elapsed_time(f1) * 1e-3 / K
    del s1
    torch.cuda.empty_cache()
    return t


def parity_check(torch, dist, gm, props, P, sl, rank, world, device):
    """See the module docstring.  Returns the dict on rank 0, None elsewhere."""
    K = PARITY_SWEEPS
    nz = sl.nodes_global[2]
    plane = sl.plane
    reset_slab(sl, P, device)
    for _ in range(K):
        sl.dwell_sweep(DT_DWELL)
    torch.cuda.synchronize()
    dist.barrier()
    # the two planes on either side of every boundary travel to rank 0
    bounds = [gm.slab.partition_planes(nz, world)[r][1] for r in range(world - 1)]  # first plane of rank r + 1
    pieces = {}
    for b, kb in enumerate(bounds):
        lo_rank, hi_rank = b, b + 1
        if rank == lo_rank:
            below = sl.owned(sl.T)[-2 * plane:].clone()          # planes kb - 2, kb - 1
            if rank == 0:
                pieces[(b, 0)] = below
            else:
                dist.send(below, 0)
        if rank == hi_rank:
            above = sl.owned(sl.T)[:2 * plane].clone()           # planes kb, kb + 1
            dist.send(above, 0)
        if rank == 0:
            if lo_rank != 0:
                t = torch.empty(2 * plane, device=device)
                dist.recv(t, lo_rank)
                pieces[(b, 0)] = t
            t = torch.empty(2 * plane, device=device)
            dist.recv(t, hi_rank)
            pieces[(b, 1)] = t
    out = None
    if rank == 0:
        worst, bad, checked = 0.0, 0, 0
        m = K + 3
        for b, kb in enumerate(bounds):
            c0, c1 = max(kb - m, 0), min(kb + m, nz)
            cut = gm.slab.Level1Slab(gm, props, (L1_NX, L1_NY, c1 - c0), L1_H, 0, 1, [P["T_amb"]] * 5, device=device)
            cut.set_owned(initial_T(torch, P, nz, c0, c1, device), torch.ones((c1 - c0) * plane, device=device))
            for _ in range(K):
                cut.dwell_sweep(DT_DWELL)
            want = cut.T[(kb - 2 - c0) * plane:(kb + 2 - c0) * plane]
            got = torch.cat([pieces[(b, 0)], pieces[(b, 1)]])
            d = (got - want).abs()
            worst = max(worst, float(d.max()))
            bad += int((got != want).sum())
            checked += 4
            del cut
        moved = float((pieces[(0, 0)] - initial_T(torch, P, nz, bounds[0] - 2, bounds[0], device)).abs().max())
        out = {"ok": bad == 0 and moved > 0.0, "max_abs_diff": worst, "nodes_differing": bad, "planes_checked": checked,
               "boundaries": len(bounds), "sweeps": K, "max_change_of_a_checked_plane_K": moved,
               "how": "rank 0 recomputes a cut of 2(K+3) planes around every slab boundary alone (no halo exchange) and "
                      "compares the two planes on either side with the N-GPU result, bit for bit"}
    dist.barrier()
    return out


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
    l0 = gm.ops.LAUNCHES
    t0 = time.time()
    total_s = timed_sweeps(torch, dist, sl, K, 0, world, device)
    t1 = time.time()
    launches = gm.ops.LAUNCHES - l0
    # end to end: host buffers of the owned planes up, one sweep, owned planes down (per step = per sweep)
    nown = (sl.k1 - sl.k0) * sl.plane
    hT = torch.empty(nown, dtype=torch.float32).pin_memory()
    hT.copy_(sl.owned(sl.T))
    hS = torch.ones(nown, dtype=torch.float32).pin_memory()
    oT = torch.empty(nown, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te0 = time.perf_counter()
    Ke = max(1, min(K, 5))
    for _ in range(Ke):
        sl.owned(sl.T).copy_(hT, non_blocking=True)
        sl.fill_ghosts()                      # halo fill of the uploaded field (restarts the fused protocol)
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
    del hT, hS
    # same-workload single-GPU reference (one slab of the same size, no neighbours), timed on rank 0 outside
    # the timed region, so that a weak-scaling efficiency can be read off this one line; plus the round-1
    # configuration (25 planes per GPU) and the round-1 mechanism (push kernel + barrier launch) for comparison
    one_gpu, alt, par, drop_in = None, {}, None, None
    if world > 1:
        if rank == 0:
            one_gpu = one_gpu_reference(torch, gm, props, P, SLAB_PLANES, K, W, device)
        dist.barrier()
        par = parity_check(torch, dist, gm, props, P, sl, rank, world, device)
        try:
            from bench_tools.dist_check import run as drop_in_check

            drop_in = drop_in_check(torch, dist, gm, rank, world)
        except Exception as exc:  # reported, not fatal for the timing line
            drop_in = {"ok": False, "error": repr(exc)}
        del sl
        torch.cuda.empty_cache()
        for name, planes, fused in (("slab_50_planes", 50, True), ("slab_25_planes", 25, True),
                                    ("slab_25_planes_round1_push_barrier", 25, False)):
            if planes == SLAB_PLANES and fused:
                continue
            s2 = make_slab(gm, props, rank, world, device, P, planes=planes, fused=fused)
            t_n = timed_sweeps(torch, dist, s2, K, W, world, device) / K
            del s2
            torch.cuda.empty_cache()
            if rank == 0:
                t_1 = one_gpu_reference(torch, gm, props, P, planes, K, W, device)
                alt[name] = {"ms_per_step": 1e3 * t_n, "one_gpu_ms_per_step": 1e3 * t_1,
                             "weak_scaling_efficiency": t_1 / t_n, "nodes": L1_NX * L1_NY * planes * world}
            dist.barrier()
    symmetric, fused = (world > 1 and os.environ.get("GOMELT_SLAB_NCCL", "0") != "1"), True
    if rank == 0:
        peaks = read_peaks()
        value = K * nn_total / total_s
        achieved = B_ALG_L1 * (nn_total / world) * K / total_s / 1e9  # per GPU
        how = ("fused level step incl. surface flux and Dirichlet faces, then one exchange kernel: the two boundary planes "
               "go into the neighbours' ghost planes over NVLink peer memory, release / acquire counters order the "
               "sweeps: two launches per sweep, no barrier launch, no NCCL call" if symmetric else
               "fused level step incl. surface flux + one-plane T halo exchange by NCCL send/recv per sweep"
               if world > 1 else "fused level step incl. surface flux and Dirichlet faces, one launch per sweep")
        line = {
            "metric": "Level-1 DOF-updates/s", "value": value, "unit": "DOF-updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * total_s / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"L1-slab: {L1_NX}x{L1_NY}x{SLAB_PLANES * world}-node Level-1 grid "
                                   f"({nn_total} nodes), z-slabs of {SLAB_PLANES} planes per GPU, dwell sweeps ({how})",
                       "nodes": nn_total, "parallelism": f"z-slab x{world}",
                       "l2": f"working set per GPU {12 * nn_total // world // 1000000} MB > 126 MB L2 (inputs larger than L2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "level_step_v3",
                         "bytes_per_dof": B_ALG_L1, "peak_source": peaks["source"],
                         "note": "per GPU, whole sweep (halo exchange included)"},
            "e2e": {"value": Ke * nn_total / e2e_s, "unit": "DOF-updates/s",
                    "h2d_bytes_per_step": 4 * nown * world, "d2h_bytes_per_step": 4 * nown * world,
                    "api": "host-buffer sweep: upload owned T planes, halo fill, one dwell sweep, download"},
            "gpu_launches": launches, "clocks": clocks,
            "weak_scaling_efficiency": None if one_gpu is None else one_gpu / (total_s / K),
            "same_workload_1gpu": None if one_gpu is None else {
                "ms_per_step": 1e3 * one_gpu, "value": L1_NX * L1_NY * SLAB_PLANES / one_gpu,
                "weak_scaling_efficiency": one_gpu / (total_s / K),
                "note": f"one {SLAB_PLANES}-plane slab on rank 0, no neighbours, timed right after the N-GPU region; the "
                        "driver's own `efficiency` divides by the N=1 line, which is a different metric (Level-3 window)"},
            "parity_check": par, "drop_in_parity": drop_in, "other_configurations": alt,
            "halo_bytes_per_step_per_gpu": 4 * L1_NX * L1_NY * (2 if world > 2 else (1 if world == 2 else 0)),
            "halo": ("exchange kernel after the step: peer stores + release / acquire counters (symmetric memory)"
                     if symmetric else "NCCL send/recv") if world > 1 else "none",
        }
        if single_gpu:
            return line
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if par is not None and not par["ok"]:
        sys.exit(3)
    return None
