"""Level-3 subcycle blocks on HOST buffers, software-pipelined over three CUDA streams.

A caller that keeps its fields on the host (bench.py's ``e2e`` leg; a driver that snapshots every block)
pays two PCIe transfers of the window per block - 82 MB each way at the 10 M-node window - against 0.25 ms
of arithmetic.  The transfers of consecutive blocks are independent of each other's arithmetic, so the
pipeline keeps ``depth`` blocks in flight: block i+1 uploads (copy-in stream) while block i runs its N3
substeps (compute stream) and block i-1 downloads (copy-out stream); the two DMA directions run concurrently.
Every block still makes the one C-ABI call subcycleGOMELT makes (``gomelt_l3_substeps_f32``, cF:3367-3412).

All host tensors must be pinned (``torch.Tensor.pin_memory``); results are valid after ``drain()`` or after the
event returned by ``submit`` has completed.

The state S1 of a window level is 0 / 1 by construction (it is re-thresholded by every computeStateProperties, cF:2589;
fractional values exist on Level 1 only), so the host side may hold it as ``uint8``: a quarter of the bytes on the
wire, widened / narrowed on the device (the kernels keep the reference's float32 state).  ``submit`` takes either
dtype for ``hS1`` / ``oS1``.
"""
import torch


class HostBlockPipeline:
    def __init__(self, ops, props, grid, coords, *, depth=2, n_rows=5, n_substrate=0, flags=0, faces=None, device=None):
        self.ops, self.props, self.grid, self.coords = ops, props, grid, coords
        self.n_substrate, self.flags, self.faces = n_substrate, flags, faces
        self.nn = grid.nx * grid.ny * grid.nz
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        mk = lambda: torch.empty(self.nn, dtype=torch.float32, device=dev)
        mk8 = lambda: torch.empty(self.nn, dtype=torch.uint8, device=dev)
        self.slots = [{"Ta": mk(), "Tb": mk(), "S1": mk(), "S8": None, "mk8": mk8, "free": None} for _ in range(depth)]
        self.tables = torch.empty(n_rows * (grid.nx + grid.ny + grid.nz), dtype=torch.float32, device=dev)
        self.k = 0

    def submit(self, hT, hS1, rows, oT, oS1):
        """Queue one block: (hT, hS1) -> device -> len(rows) substeps -> (oT, oS1).  Returns the CUDA event that
        marks its results complete on the host side."""
        slot = self.slots[self.k % len(self.slots)]
        self.k += 1
        if slot["free"] is not None:          # the slot's previous block has been downloaded
            self.s_in.wait_event(slot["free"])
        with torch.cuda.stream(self.s_in):
            slot["Ta"].copy_(hT, non_blocking=True)
            if hS1.dtype == torch.uint8:   # bytes on the wire, float32 state on the device
                if slot["S8"] is None:
                    slot["S8"] = slot["mk8"]()
                slot["S8"].copy_(hS1, non_blocking=True)
                slot["S1"].copy_(slot["S8"])
            else:
                slot["S1"].copy_(hS1, non_blocking=True)
            up = torch.cuda.Event()
            up.record()
        self.s_run.wait_event(up)
        with torch.cuda.stream(self.s_run):
            T = self.ops.l3_substeps(self.props, self.grid, self.coords, rows, slot["Ta"], slot["Tb"], slot["Ta"],
                                     slot["S1"], self.tables, n_substrate=self.n_substrate, flags=self.flags,
                                     faces=self.faces)
            ran = torch.cuda.Event()
            ran.record()
        self.s_out.wait_event(ran)
        with torch.cuda.stream(self.s_out):
            oT.copy_(T, non_blocking=True)
            if oS1.dtype == torch.uint8:
                if slot["S8"] is None:
                    slot["S8"] = slot["mk8"]()
                slot["S8"].copy_(slot["S1"])   # exact: the state is 0.0 / 1.0
                oS1.copy_(slot["S8"], non_blocking=True)
            else:
                oS1.copy_(slot["S1"], non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        slot["free"] = done
        return done

    def drain(self):
        for s in (self.s_in, self.s_run, self.s_out):
            s.synchronize()
