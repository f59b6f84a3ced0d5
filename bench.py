#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 GO-MELT step (contract: task prompt / DESIGN.md section 7).

N = 1  workload "L3-10M": BASELINE.json configs[1] — single laser track, Level-3 melt-pool window
       512 x 512 x 38 elements (513*513*39 = 10 263 591 nodes), h = 0.02 mm, T-dependent
       properties (examples/example.json property block), dt = 1e-5 s.  One *step* = one Level-3
       subcycle block of N3 = 5 explicit substeps (the inner scan of subcycleGOMELT, cF:3367-3412):
       per substep  source tables (computeSourcesL3) + top-surface flux (computeConvRadBC) +
       fused level step (computeStateProperties + solveMatrixFreeFE + clamp) + face prolongation
       from the Level-2 parent (assignBCsFine), laser advancing along +x.  metric = Level-3 DOF-updates/s (1 DOF-update = one node advanced one sweep).
N > 1  workload "L1-slab": BASELINE.json configs[4] — part-scale Level-1 mesh z-slab-decomposed, one
       rank per GPU, dwell sweeps (stepGOMELTDwellTime cF:2617-2664) with a one-plane halo exchange
       per sweep (one C-ABI call = the fused level step + an exchange kernel: boundary planes into the neighbours'
       ghost planes over NVLink peer memory, release / acquire counters, no NCCL and no barrier launch in the sweep;
       GOMELT_SLAB_NCCL=1: NCCL send/recv); weak scaling (100 planes of 1001 x 1001 per GPU).  metric = Level-1 DOF-updates/s.

`--impl reference` times the reference algorithm on the host cores: element gather, 8 x 8 apply, scatter-add as
dense tensor operations on all host threads (oracle/torch_cpu.py, pinned to the NumPy oracle; "restated
reference, not JAX/XLA": JAX is not installable here or on the GPU box) on the SAME 10 M-node window, a bounded
number of blocks.

Beyond the contract keys the N = 1 line carries: `roofline` with K1's duration from CUDA events recorded inside the
native call; `e2e` (host buffers, state as bytes on the wire, three steps in flight) with `f32_state_value`,
`serial_value` and `resident_state` (state on the device as in the driver, toolpath rows in, monitor out);
`l1_slab_1gpu` (the N > 1 workload on one GPU); `whole_step` (a whole subcycleGOMELT / stepGOMELT + moveEverything at
C2 scale with a per-kernel table); `example_json` (config 0 through the driver: wall-s per sim-s, graph replays against
row by row); `config4_full_size` (config 3 at full size).  Every N > 1 line carries `weak_scaling_efficiency`
(against the same slab alone), `parity_check` (slab boundaries bit for bit) and `drop_in_parity` (the whole driver on
N ranks against the plain run, bit for bit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EXAMPLE_PROPS = {
    "thermal_conductivity_powder": 0.4, "thermal_conductivity_bulk_a0": 4.23,
    "thermal_conductivity_bulk_a1": 0.016, "thermal_conductivity_fluid_a0": 29.0,
    "heat_capacity_solid_a0": 383.1, "heat_capacity_solid_a1": 0.174, "heat_capacity_mushy": 3235.0,
    "heat_capacity_fluid": 769.0, "density": 8e-06, "laser_radius": 0.1, "laser_depth": 0.1,
    "laser_power": 285.0, "laser_absorptivity": 0.45, "T_amb": 298.15, "T_solidus": 1533,
    "T_liquidus": 1609, "T_boiling": 3038.0, "h_conv": 1.5e-05, "emissivity": 0.3,
    "evaporation_coefficient": 0.82, "latent_heat_evap": 6457000.0, "molar_mass": 58.69,
    "layer_height": 0.04,
}
L3_ELEMENTS = (512, 512, 38)
L3_H = 0.02
N3 = 5
DT = 1e-5
LASER_V = 1000.0  # mm/s
B_ALG_L3 = 16     # bytes per Level-3 DOF-update: T0 r4 + S1 r4 + T w4 + S1 w4
B_ALG_L1 = 12     # Level-1 dwell sweep: T r/w + S1 r


def host_properties():
    """SetupProperties semantics (cF:267-345) without importing the oracle: product-side schema."""
    import gomelt_b200 as gm

    return gm.schema.SetupProperties(EXAMPLE_PROPS)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                c = float(f[1])
                smax = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(c)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than one sample: fall back to every sample taken
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except (ValueError, IndexError):
                    pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# Level-3 block on the device (N = 1)
# --------------------------------------------------------------------------------------------
class L3Block:
    """Device-resident Level-3 window advanced by subcycle blocks through the product API."""

    def __init__(self, elements=L3_ELEMENTS, h=L3_H, seed=0):
        import numpy as np
        import torch

        import gomelt_b200 as gm

        self.gm, self.torch, self.np = gm, torch, np
        self.P = host_properties()
        self.props = gm._lib.make_props(self.P)
        ex, ey, ez = elements
        self.nodes = (ex + 1, ey + 1, ez + 1)
        self.grid = gm._lib.make_grid(self.nodes, (h, h, h))
        self.nn = self.nodes[0] * self.nodes[1] * self.nodes[2]
        nx, ny, nz = self.nodes
        x = np.linspace(0.0, ex * h, nx, dtype=np.float32)
        y = np.linspace(0.0, ey * h, ny, dtype=np.float32)
        z = np.linspace(-ez * h, 0.0, nz, dtype=np.float32)
        self.coords_host = (x, y, z)
        self.coords = [torch.as_tensor(c).cuda() for c in (x, y, z)]
        # initial state: T_amb + smooth +-50 K perturbation (rng seed 0),
        # bulk below z = -0.04 (S1 = 1), one powder layer on top (S1 = 0)
        rng = np.random.default_rng(seed)
        pert = 50.0 * np.sin(np.linspace(0, 9, nx))[None, None, :] * np.cos(np.linspace(0, 7, ny))[None, :, None] \
            * np.ones(nz)[:, None, None]
        pert = pert + rng.uniform(-1.0, 1.0, size=(nz, ny, nx))
        T0 = (self.P["T_amb"] + 51.0 + pert).astype(np.float32).reshape(-1)
        S1 = np.repeat((z <= -0.04 + 1e-6).astype(np.float32), nx * ny)
        self.T0_host, self.S1_host = T0, S1
        self.n_sub = int((z < 1e-5 - 0.04).sum()) * nx * ny * 0  # no substrate override in the window
        self.Ta = torch.as_tensor(T0).cuda()
        self.Tb = torch.empty_like(self.Ta)
        self.S1 = torch.as_tensor(S1).cuda()
        self.cur = self.Ta
        self.tables = torch.empty(N3 * (nx + ny + nz), device="cuda")
        self.tx, self.ty, self.tz = self.tables[:nx], self.tables[nx:nx + ny], self.tables[nx + ny:nx + ny + nz]
        # Level-2 parent (same element counts at h2 = 2 h3 enclosing the window, top planes
        # aligned): its new / old temperature fields feed the per-substep face prolongation of the Level-3
        # window (assignBCsFine cF:1598-1620 with the time blend of cF:3386-3389), as in subcycleGOMELT
        h2 = 2.0 * h
        x2 = np.linspace(-0.5 * ex * h, -0.5 * ex * h + ex * h2, nx, dtype=np.float32)
        y2 = np.linspace(-0.5 * ey * h, -0.5 * ey * h + ey * h2, ny, dtype=np.float32)
        z2 = np.linspace(-ez * h2, 0.0, nz, dtype=np.float32)
        self.coords2 = [torch.as_tensor(c).cuda() for c in (x2, y2, z2)]
        par = (self.P["T_amb"] + 51.0 + 50.0 * np.sin(np.linspace(0, 5, nx))[None, None, :]
               * np.cos(np.linspace(0, 4, ny))[None, :, None] * np.ones(nz)[:, None, None]).astype(np.float32).reshape(-1)
        self.L2new = torch.as_tensor(par).cuda()
        self.L2old = torch.as_tensor((par - 0.5).astype(np.float32)).cuda()
        self.faces = (self.coords2, self.L2new, self.L2old, float(N3), float(self.P["T_amb"]))
        self.step_flags = gm.ops.STEP_CLAMP | gm.ops.STEP_SKIP_FACES
        self.laser = np.array([0.25 * ex * h, 0.5 * ey * h, 0.0], np.float32)
        self.k1_events = []

    def _rows(self):
        """The next N3 toolpath rows (x, y, z, Ljump, Ldwell, dt, P - cP:71-74): laser advancing along +x."""
        np = self.np
        rows = np.zeros((N3, 7), np.float32)
        for i in range(N3):
            self.laser[0] += LASER_V * DT
            rows[i] = (self.laser[0], self.laser[1], self.laser[2], 1, 1, DT, self.P["laser_power"])
        return rows

    def block(self):
        """One Level-3 subcycle block through the product's one-call inner scan (gomelt_l3_substeps_f32), the call
        subcycleGOMELT makes (cF:3367-3412): 1 source-table launch + N3 x (fused level step that leaves the five
        Dirichlet faces + face prolongation from the Level-2 parent).  Returns the tensor holding the newest T."""
        ops = self.gm.ops
        other = self.Tb if self.cur is self.Ta else self.Ta
        self.cur = ops.l3_substeps(self.props, self.grid, self.coords, self._rows(), self.cur, other, self.cur,
                                   self.S1, self.tables, n_substrate=self.n_sub, flags=self.step_flags,
                                   faces=self.faces)
        return self.cur

    def block_k1_events(self):
        """One more block through the same call, with a pair of CUDA events recorded INSIDE the native call right before
        and after every fused level step (gomelt_substeps_args_t.step_events): the kernel's duration as it runs in the
        timed step, on the launching stream, without a Python issue path between the two records (events around a
        ``ops.level_step`` call from Python read 2 us more on an idle GPU: bench_tools/event_overhead.py)."""
        ops, torch = self.gm.ops, self.torch
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * N3)]
        other = self.Tb if self.cur is self.Ta else self.Ta
        self.cur = ops.l3_substeps(self.props, self.grid, self.coords, self._rows(), self.cur, other, self.cur,
                                   self.S1, self.tables, n_substrate=self.n_sub, flags=self.step_flags,
                                   faces=self.faces, step_events=evs)
        self.k1_events += [(evs[2 * i], evs[2 * i + 1]) for i in range(N3)]
        return self.cur


def run_gomelt_single(args):
    import numpy as np
    import torch

    import gomelt_b200 as gm

    torch.cuda.set_device(0)
    gm.load()
    ops = gm.ops
    K, W = args.steps, max(args.warmup, 3)
    blk = L3Block()
    nn = blk.nn
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")  # 256 MiB > 126 MB L2
    flush_rd = torch.zeros(256 * 1024 * 1024 // 4, device="cuda")

    def l2_flush():
        """Write a buffer larger than L2, then read a second one: the inputs are evicted AND the dirty lines of
        the flush itself are written back before the timed region starts (cold, clean L2)."""
        flush.zero_()
        flush_rd.max()
    for _ in range(W):
        blk.block()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps -------------
    evs = []
    launches0 = ops.LAUNCHES
    t_wall0 = time.time()
    torch.cuda.synchronize()
    for _ in range(K):
        l2_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blk.block()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = ops.LAUNCHES - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_s = sum(step_ms) * 1e-3
    value = K * N3 * nn / total_s
    # ---- roofline leg: the same blocks once more (same flush, same one-call path) with CUDA events recorded inside the
    # native call around every fused level step ----
    for _ in range(min(K, 10)):
        l2_flush()
        blk.block_k1_events()
    torch.cuda.synchronize()
    k1_ms = [a.elapsed_time(b) for a, b in blk.k1_events]
    k1_avg_s = (sum(k1_ms) / len(k1_ms)) * 1e-3
    k1_first = [k1_ms[i] for i in range(0, len(k1_ms), N3)]   # first substep of a block: cold L2 (flushed)
    # ---- end-to-end through the host-buffer API ----------------------------------------------
    e2e = run_e2e_hostbuffers(blk, K)
    clocks = sampler.stop(t_wall0, time.time())
    peaks = read_peaks()
    achieved = B_ALG_L3 * nn / k1_avg_s / 1e9
    traffic = read_traffic("level_step_v3")
    line = {
        "metric": "Level-3 DOF-updates/s", "value": value, "unit": "DOF-updates/s", "n_gpus": 1,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * total_s / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "L3-10M: 512x512x38-element Level-3 window (10263591 nodes), one step = "
                               "N3=5 substeps through gomelt_l3_substeps_f32, the call subcycleGOMELT makes (1 launch "
                               "for all source tables + 5 x [fused level step incl. state/properties, surface flux, "
                               "source, clamp, Dirichlet faces left + face prolongation from the Level-2 parent]), "
                               "T-dependent properties, dt=1e-5, moving laser",
                   "nodes": nn, "substeps_per_step": N3,
                   "l2": "flushed between steps (256 MiB write, then 256 MiB read so the flush's own dirty lines are "
                         "written back, outside the timed events); per-step CUDA events summed", "state": "device-resident (value) / host buffers (e2e)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                     "kernel": "level_step_v3", "bytes_per_dof": B_ALG_L3,
                     "kernel_us": k1_avg_s * 1e6, "peak_source": peaks["source"],
                     "kernel_us_first_substep_cold_l2": 1e3 * sum(k1_first) / len(k1_first),
                     "how": "CUDA events recorded on the launching stream INSIDE gomelt_l3_substeps_f32 right before / after "
                            "every level_step_v3 launch (gomelt_substeps_args_t.step_events), over the same blocks issued "
                            "once more right after the timed region with the same L2 flush: the kernel as it runs in the "
                            "timed step (substeps 2..5 of a block find part of their input in L2, as they do there); "
                            "events around a Python-issued launch read ~2 us more (bench_tools/event_overhead.py)"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    # the path that shards (Level-1 z-slabs, the N>1 workload) measured on this one GPU, so that the
    # N>1 lines have a same-workload denominator
    try:
        from bench_tools.bench_l1_slab import run_gomelt_multi

        del blk, flush, flush_rd
        torch.cuda.empty_cache()
        ref = run_gomelt_multi(args, read_peaks, lambda i: None, host_properties, single_gpu=True)
        line["l1_slab_1gpu"] = {k: ref[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "e2e")}
        line["l1_slab_1gpu"]["workload"] = ref["config"]["workload"]
    except Exception as exc:  # the headline line must still print
        line["l1_slab_1gpu"] = {"error": repr(exc)}
    # the reference's default run end to end (before the whole-step block: that one traces its kernels through CUPTI,
    # which stays attached to the process and taxes every later launch of this host-issue-bound run)
    try:
        torch.cuda.empty_cache()
        line["example_json"] = example_json_run()
    except Exception as exc:
        line["example_json"] = {"error": repr(exc)}
    # whole steps through the drop-in entry points at the same scale (10 M-node Level 3 WITH its Level 2 / Level 1:
    # subcycleGOMELT, stepGOMELT, moveEverything, the T' projections ...) and the reference's default run end to end
    try:
        from bench_tools.bench_whole_step import run as whole_step

        torch.cuda.empty_cache()
        line["whole_step"] = whole_step(EXAMPLE_PROPS, peaks["hbm_gbs"])
    except Exception as exc:
        line["whole_step"] = {"error": repr(exc)}
    try:  # BASELINE.json configs[3] at full size (50 M-node Level 1, three layers): wall-s per sim-s through the driver
        import torch

        torch.cuda.empty_cache()
        from bench_tools.run_config4 import run as config4_run

        line["config4_full_size"] = config4_run()
        torch.cuda.empty_cache()
    except Exception as exc:
        line["config4_full_size"] = {"error": repr(exc)}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(line))


def example_json_run():
    """BASELINE.json configs[0]: examples/example.json verbatim (1299 toolpath rows = 50 stepGOMELT + 30 subcycleGOMELT
    + 499 dwell steps, 579 moveEverything; 1.006 simulated seconds) through the drop-in driver -> wall-s per sim-s."""
    import tempfile

    import torch

    import gomelt_b200 as gm
    from bench_tools.run_example import load_input

    res = eager = None
    for _ in range(2):  # the first pass warms up module loading / allocator / kernel images
        eager = gm.driver.go_melt(load_input(tempfile.mkdtemp()), write_final=False, graphs=False)
    walls, best = [], None
    cf = gm.computeFunctions
    for _ in range(4):   # host-issue-bound: the fastest of the repeats is reported, all are listed
        # every repeat starts like a first run: no coordinate array on the device, no per-position pair descriptor,
        # no captured graph (those live in the run's own workspace) - only the loaded library and the warm allocator
        cf._PAIR_CACHE.clear()
        cf._CACHE.store.clear()
        l0, g0 = gm.ops.LAUNCHES, gm.ops.GRAPH_LAUNCHES
        res = gm.driver.go_melt(load_input(tempfile.mkdtemp()), write_final=False)
        walls.append(round(res["wall_seconds"], 4))
        if best is None or res["wall_seconds"] < best["wall_seconds"]:
            best = res
    res = best
    torch.cuda.synchronize()
    same = all(torch.equal(res["Levels"][i]["T0"], eager["Levels"][i]["T0"]) for i in (1, 2, 3))
    return {"workload": "examples/example.json + example.gcode, whole run, device-resident state, no file output; the "
                        "499 rows of the pause replay a CUDA graph of two rows (computeFunctions.dwellRows); the host-side caches "
                        "(device copies of coordinate arrays, per-position pair descriptors) are emptied before every repeat",
            "wall_s": res["wall_seconds"], "wall_s_of_every_repeat": walls, "sim_s": res["sim_seconds"],
            "wall_s_per_sim_s": res["wall_seconds"] / res["sim_seconds"], "toolpath_rows": res["time_inc"],
            "counts": res["counts"], "lib_launches": gm.ops.LAUNCHES - l0,
            "kernels_replayed_from_graphs": gm.ops.GRAPH_LAUNCHES - g0,
            "row_by_row": {"wall_s": eager["wall_seconds"], "wall_s_per_sim_s": eager["wall_seconds"] / eager["sim_seconds"],
                           "fields_bit_equal_to_the_graph_run": bool(same)}}


def run_e2e_hostbuffers(blk, K):
    """Same block through the host-buffer API (gomelt_b200/hostpipe.py): pinned host T0,S1 -> device -> N3 substeps
    through gomelt_l3_substeps_f32 -> pinned host T,S1, every step, copies inside the timed region.  Consecutive
    steps are independent batches, so the pipeline keeps two in flight (upload of step i+1 | substeps of step i |
    download of step i-1 on three streams); `serial` is the same loop with one step in flight."""
    import torch

    gm = blk.gm
    nn = blk.nn
    pin = lambda src=None: (torch.empty(nn, dtype=torch.float32).pin_memory() if src is None
                            else torch.as_tensor(src).clone().pin_memory())
    hT, hS = [pin(blk.T0_host), pin(blk.T0_host)], [pin(blk.S1_host), pin(blk.S1_host)]
    oT, oS = [pin(), pin()], [pin(), pin()]
    pin8 = lambda src=None: (torch.empty(nn, dtype=torch.uint8).pin_memory() if src is None
                             else torch.as_tensor(src).to(torch.uint8).pin_memory())
    hS8, oS8 = [pin8(blk.S1_host), pin8(blk.S1_host)], [pin8(), pin8()]
    out = {}
    while len(hT) < 3:   # three steps in flight keep both DMA directions busy (2: 3.9e10, 3: 4.2e10, 4: 4.1e10 DOF-updates/s)
        hT.append(pin(blk.T0_host)); oT.append(pin()); hS8.append(pin8(blk.S1_host)); oS8.append(pin8())
    for name, depth, sin, sout in (("serial", 1, hS, oS), ("pipelined", 2, hS, oS), ("pipelined_u8_state", 3, hS8, oS8)):
        pipe = gm.hostpipe.HostBlockPipeline(gm.ops, blk.props, blk.grid, blk.coords, depth=depth, n_rows=N3,
                                             n_substrate=blk.n_sub, flags=blk.step_flags, faces=blk.faces)
        m = len(sin)
        for i in range(m):  # warm-up (allocator, first-touch of the pinned buffers)
            pipe.submit(hT[i % m], sin[i % m], blk._rows(), oT[i % m], sout[i % m])
        pipe.drain()
        t0 = time.perf_counter()
        for i in range(K):
            pipe.submit(hT[i % m], sin[i % m], blk._rows(), oT[i % m], sout[i % m])
        pipe.drain()
        out[name] = K * N3 * nn / (time.perf_counter() - t0)
        del pipe
    # What a user of the drop-in does instead (the reference keeps Levels on the device between jitted calls): the state
    # stays resident, a step's inputs are its toolpath rows (host -> kernel arguments), its result the monitor the driver
    # reads back (printLevelMaxMin gm:473-474: min / max / non-finite count of the field, one fused reduction)
    mm = torch.empty(3, device="cuda")
    host_mm = torch.empty(3).pin_memory()
    for _ in range(2):
        blk.block()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        cur = blk.block()
        gm.ops.minmax(cur, out=mm)
        host_mm.copy_(mm, non_blocking=True)
        torch.cuda.synchronize()   # the monitor value is needed before the next block is issued
    resident = K * N3 * nn / (time.perf_counter() - t0)
    return {"value": out["pipelined_u8_state"], "unit": "DOF-updates/s", "h2d_bytes_per_step": 5 * nn,
            "d2h_bytes_per_step": 5 * nn, "serial_value": out["serial"], "f32_state_value": out["pipelined"],
            "f32_state_bytes_per_step_each_way": 8 * nn,
            "resident_state": {"value": resident, "unit": "DOF-updates/s", "h2d_bytes_per_step": 4 * 7 * N3,
                               "d2h_bytes_per_step": 12,
                               "api": "state device-resident as in the drop-in driver: per step the N3 toolpath rows from "
                                      "the host, the block through gomelt_l3_substeps_f32, and the min / max monitor "
                                      "(gomelt_minmax_f32) read back to pinned host memory before the next step is issued",
                               "monitor_min_max_nonfinite": [float(v) for v in host_mm]},
            "api": "gomelt_b200.hostpipe.HostBlockPipeline.submit: upload T0 (f32) and S1 (uint8 on the host and on the "
                   "wire: the state of a window level is 0 / 1; widened to the kernels' float32 on the device) from pinned "
                   "host memory; N3 substeps through gomelt_l3_substeps_f32; download T, S1 - three steps in flight on three "
                   "streams (f32_state_value: S1 as float32 on the wire, two in flight; serial_value: that with one)"}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


def read_traffic(kernel):
    p = os.path.join(ROOT, "profiles", "roofline_inputs.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if kernel in d:
            return d[kernel].get("dram_bytes_per_launch")
    return None


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy f32 restatement of the reference) on a bounded sample
# --------------------------------------------------------------------------------------------
CPU_SAMPLE_ELEMENTS = (128, 128, 38)


def _oracle_block(elements, nblocks, seed=0):
    """One worker: nblocks x N3 Level-3 substeps of the reference algorithm on its own window."""
    import numpy as np

    from oracle import computeFunctions as cF
    from oracle.util import make_level

    P = cF.SetupProperties(EXAMPLE_PROPS)
    ex, ey, ez = elements
    lv = make_level(elements, ((0.0, ex * L3_H), (0.0, ey * L3_H), (-ez * L3_H, 0.0)))
    nn = lv["nn"]
    T = np.full(nn, np.float32(P["T_amb"] + 51.0), np.float32)
    S1 = np.repeat((lv["node_coords"][2] <= -0.04 + 1e-6).astype(np.float32), lv["nodes"][0] * lv["nodes"][1])
    laser = np.array([0.25 * ex * L3_H, 0.5 * ey * L3_H, 0.0], np.float32)
    ne_nn = (0, lv["ne"], 0, 0, nn)
    t0 = time.perf_counter()
    for _ in range(nblocks * N3):
        laser[0] += LASER_V * DT
        S1, _, k, rc = cF.computeStateProperties(T, S1, P, 0)
        F = cF.computeSourcesL3(lv, laser, ne_nn, P, P["laser_power"])
        F = cF.computeConvRadBC(lv, T, lv["ne"], nn, P, F)
        T = np.maximum(np.float32(P["T_amb"]), cF.solveMatrixFreeFE(lv, nn, lv["ne"], k, rc, DT, T, F, 0))
    return nblocks * N3 * nn, time.perf_counter() - t0


def _host_threads():
    """The host cores this process may use.  Under torchrun OMP_NUM_THREADS is preset to 1; the CPU arms set their thread
    count explicitly (torch.set_num_threads) so that the baseline is the same with and without the launcher."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _torch_cpu_block(elements, nblocks, threads=None):
    """nblocks x N3 Level-3 substeps of the reference algorithm as dense tensor operations on all host threads
    (oracle/torch_cpu.py: what a multi-threaded CPU array runtime makes of the reference; pinned to the NumPy oracle by
    tests/test_oracle_kats.py)."""
    import numpy as np
    import torch

    from oracle import computeFunctions as cF
    from oracle.torch_cpu import L3SubstepCPU
    from oracle.util import make_level

    P = cF.SetupProperties(EXAMPLE_PROPS)
    ex, ey, ez = elements
    lv = make_level(elements, ((0.0, ex * L3_H), (0.0, ey * L3_H), (-ez * L3_H, 0.0)))
    port = L3SubstepCPU(lv, P, threads=threads or _host_threads())
    nn = lv["nn"]
    T = torch.full((nn,), float(P["T_amb"]) + 51.0)
    S1 = torch.from_numpy(np.repeat((lv["node_coords"][2] <= -0.04 + 1e-6).astype(np.float32), lv["nodes"][0] * lv["nodes"][1]))
    laser = np.array([0.25 * ex * L3_H, 0.5 * ey * L3_H, 0.0], np.float32)
    t0 = time.perf_counter()
    for _ in range(nblocks * N3):
        laser[0] += LASER_V * DT
        T, S1 = port.substep(T, S1, laser, P["laser_power"], DT)
    return nblocks * N3 * nn, time.perf_counter() - t0, torch.get_num_threads()


def cpu_baseline_sample():
    """The reference algorithm on this box's host cores, bounded: one block (N3 substeps) of the SAME 10.26 M-node window
    with the multi-threaded tensor port, and - for scale - one block of a 128 x 128 x 38-element window with the
    single-core NumPy oracle (the parity oracle itself)."""
    dofs, secs, threads = _torch_cpu_block(L3_ELEMENTS, 1)
    d1, s1 = _oracle_block(CPU_SAMPLE_ELEMENTS, 1)
    return {"value": dofs / secs, "unit": "DOF-updates/s", "cores": threads, "kind": "port",
            "sample": f"1 block (N3={N3} substeps) of the same Level-3 step on the SAME {'x'.join(map(str, L3_ELEMENTS))}-element "
                      f"window ({dofs // N3} nodes): the reference algorithm (element gather, 8x8 apply, scatter-add) as dense "
                      f"torch tensor operations on {threads} host threads (oracle/torch_cpu.py; not JAX/XLA, which is not "
                      f"installable here); the face prolongation of the substeps (<1 % of the work) is left out",
            "seconds": secs,
            "numpy_single_core": {"value": d1 / s1, "cores": 1, "seconds": s1,
                                  "sample": f"1 block on a {'x'.join(map(str, CPU_SAMPLE_ELEMENTS))}-element window, NumPy float32 "
                                            f"oracle (np.add.at scatter)"}}


def _torch_cpu_dwell(nodes, sweeps, h=0.2, dt=2e-3):
    """``sweeps`` Level-1 dwell sweeps (stepGOMELTDwellTime cF:2617-2664) of the reference algorithm on the host threads,
    on an (nx, ny, nz)-node slab of the N > 1 workload."""
    import numpy as np
    import torch

    from oracle import computeFunctions as cF
    from oracle.torch_cpu import L3SubstepCPU
    from oracle.util import make_level

    P = cF.SetupProperties(EXAMPLE_PROPS)
    nx, ny, nz = nodes
    lv = make_level((nx - 1, ny - 1, nz - 1), ((0.0, (nx - 1) * h), (0.0, (ny - 1) * h), (-(nz - 1) * h, 0.0)))
    port = L3SubstepCPU(lv, P, threads=_host_threads())
    nn = lv["nn"]
    T = torch.full((nn,), float(P["T_amb"]) + 200.0)
    S1 = torch.ones(nn)
    port.dwell_step(T[:0].new_full((nn,), float(P["T_amb"]) + 200.0), S1, dt, [float(P["T_amb"])] * 5)   # warm-up
    t0 = time.perf_counter()
    for _ in range(sweeps):
        T = port.dwell_step(T, S1, dt, [float(P["T_amb"])] * 5)
    return sweeps * nn, time.perf_counter() - t0, torch.get_num_threads()


def run_reference_multi(args):
    """--impl reference under the N > 1 launch: the arm's metric there is Level-1 DOF-updates/s of dwell sweeps, so the
    CPU arm times dwell sweeps too - a bounded slab of the same grid (1001 x 1001 x 10 nodes) on all host threads."""
    nodes = (1001, 1001, 10)
    Kb = max(1, min(args.steps, 3))
    dofs, secs, threads = _torch_cpu_dwell(nodes, Kb)
    value = dofs / secs
    sample = (f"{Kb} dwell sweep(s) of a {nodes[0]}x{nodes[1]}x{nodes[2]}-node slab of the part-scale grid with the multi-threaded "
              f"tensor port of the reference algorithm ({threads} threads); the GPU arm sweeps {args.gpus} x 100 planes of the same "
              "grid.  JAX is not installable on this box.")
    line = {
        "impl": "reference", "metric": "Level-1 DOF-updates/s", "value": value, "unit": "DOF-updates/s", "n_gpus": args.gpus,
        "steps": Kb, "warmup": args.warmup, "ms_per_step": 1e3 * secs / Kb, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "L1-slab dwell sweeps (bounded sample: one 10-plane slab on the host cores)", "sample": sample,
                   "nodes": nodes[0] * nodes[1] * nodes[2]},
        "cpu_baseline": {"value": value, "unit": "DOF-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_reference_multi(args)
        return
    import multiprocessing as mp

    ncores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # bounded: every "step" = one N3-substep block; K capped so that the run ends within a few minutes
    Kb = max(1, min(K, 2))
    # (a) the multi-threaded tensor port on the SAME window as the GPU arm (all host threads)
    if W > 0:
        _torch_cpu_block((64, 64, 16), 1)
    dofs_t, secs_t, threads = _torch_cpu_block(L3_ELEMENTS, Kb)
    # (b) the NumPy oracle, one process per core on independent smaller windows (what round 1 reported)
    nproc = max(1, min(ncores, 32))
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    elements = (96, 96, 38)
    with mp.get_context("fork").Pool(nproc) as pool:
        t0 = time.perf_counter()
        res = pool.starmap(_oracle_block, [(elements, 1, i) for i in range(nproc)])
        wall_n = time.perf_counter() - t0
    v_torch, v_numpy = dofs_t / secs_t, sum(r[0] for r in res) / wall_n
    value = max(v_torch, v_numpy)
    best = "tensor port" if v_torch >= v_numpy else "NumPy oracle, one process per core"
    sample = (f"{Kb} block(s) x N3={N3} substeps of the SAME {'x'.join(map(str, L3_ELEMENTS))}-element Level-3 window with the "
              f"multi-threaded tensor port of the reference algorithm ({threads} threads): {v_torch:.3e} DOF-updates/s; and 1 block "
              f"on {nproc} independent {'x'.join(map(str, elements))}-element windows with the NumPy oracle, one process per "
              f"core: {v_numpy:.3e}; value = the faster of the two ({best}).  JAX is not installable on this box.")
    line = {
        "impl": "reference", "metric": "Level-3 DOF-updates/s", "value": value, "unit": "DOF-updates/s",
        "n_gpus": args.gpus, "steps": Kb, "warmup": W, "ms_per_step": 1e3 * secs_t / Kb,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "L3-10M: 512x512x38-element Level-3 window (10263591 nodes), one step = N3=5 substeps "
                               "(bounded: fewer steps than the GPU arm, same window)", "sample": sample,
                   "nodes": (L3_ELEMENTS[0] + 1) * (L3_ELEMENTS[1] + 1) * (L3_ELEMENTS[2] + 1), "substeps_per_step": N3},
        "cpu_baseline": {"value": value, "unit": "DOF-updates/s", "cores": threads if v_torch >= v_numpy else nproc,
                         "kind": "port", "sample": sample,
                         "tensor_port": {"value": v_torch, "threads": threads}, "numpy_multiprocess": {"value": v_numpy, "procs": nproc}},
        "e2e": {"value": value, "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gomelt", choices=["gomelt", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from bench_tools.bench_l1_slab import run_gomelt_multi

        run_gomelt_multi(args, read_peaks, ClockSampler, host_properties)
        return
    run_gomelt_single(args)


if __name__ == "__main__":
    main()
