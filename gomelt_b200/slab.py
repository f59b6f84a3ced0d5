"""Level-1 z-slab decomposition across the GPUs of one box (DESIGN.md "Multi-GPU").

The reference is single-device (gm:557).  Here the part-scale Level-1 grid is cut into contiguous
z-slabs, one rank per GPU.  Node numbering is x-fastest (cF:646-689), so a z-plane is ``nx*ny``
contiguous floats and a halo is a zero-copy slice.  Each rank stores its owned planes ``[k0, k1)``
plus one ghost plane per neighbour; an explicit sweep (stepGOMELTDwellTime cF:2617-2664) needs the
neighbour's boundary plane of T (27-point stencil, radius 1) and - once - of S1.

Per sweep, device-side path (``symmetric=True``, the default on GPUs): the two temperature buffers and a block of
uint32 counters live in peer-mapped symmetric memory (torch.distributed._symmetric_memory: CUDA VMM handles exchanged
once), and a sweep is ONE C-ABI call (gomelt_level_step_f32 with ``halo_sync``, see gomelt_abi.h) = two launches: the
fused level step (Dirichlet faces included), then ``halo_exchange_kernel``, which copies the two boundary planes into
the neighbours' ghost planes with 16-byte stores over NVLink, signals the neighbours' arrival counters (system-scope
release) and waits (acquire) for theirs.  No NCCL call, no barrier launch, no host involvement between sweeps.
(``fused=False`` keeps the round-1 form for A/B: ``halo_push_kernel`` after the step, a symmetric-memory barrier
launch per sweep.)

Fallback path (``symmetric=False``; CPU tensors in the gloo tests, or GOMELT_SLAB_NCCL=1 for A/B):
  1. K1 on the two boundary planes of the owned range,
  2. those planes go to the neighbours' ghost planes (NCCL send/recv over NVLink, side stream),
  3. K1 on the interior planes overlaps the transfer,
  4. the main stream joins the transfer; buffers swap.

Host logic only: the arithmetic is ``ops.level_step`` / ``ops.surface_flux`` (CUDA, no CPU path).
``partition_planes`` and ``exchange_planes`` are device-agnostic so that the world_size-2 gloo
tests can exercise the N>1 plumbing on CPU tensors.
"""
import torch
import torch.distributed as dist


def partition_planes(nz, world):
    """Contiguous, balanced plane ranges [(k0, k1)] of ``nz`` planes over ``world`` ranks."""
    if world < 1 or nz < world:
        raise ValueError(f"cannot cut {nz} planes into {world} slabs")
    base, extra = divmod(nz, world)
    out, k = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((k, k + n))
        k += n
    return out


def partition_active_planes(nz, nz_active, world):
    """Plane ranges that balance the ACTIVE planes [0, nz_active): the planes above the powder bed cost a store per
    node and no stencil, so they all go to the last rank on top of its share (layer activation, cF:495-517: a z-slab
    decomposition of all nz planes would leave the upper ranks idle until the build reaches them)."""
    nz_active = max(min(int(nz_active), nz), world)
    parts = partition_planes(nz_active, world)
    parts[-1] = (parts[-1][0], nz)
    return parts


def local_extent(rank, world, k0, k1):
    """(first stored global plane, stored plane count, z_begin, z_end) of a rank's local array."""
    lo = 1 if rank > 0 else 0
    hi = 1 if rank < world - 1 else 0
    return k0 - lo, (k1 - k0) + lo + hi, lo, lo + (k1 - k0)


def boundary_transfers(field, plane, z_begin, z_end, rank, world):
    """(sends, recvs) = [(tensor, peer)] of one ghost exchange of ``field`` (flat, x-fastest, local planes): the first /
    last owned plane goes to the lower / upper neighbour's ghost plane, theirs come back."""
    sends, recvs = [], []
    if rank > 0:
        sends.append((field[z_begin * plane:(z_begin + 1) * plane], rank - 1))
        recvs.append((field[(z_begin - 1) * plane:z_begin * plane], rank - 1))
    if rank < world - 1:
        sends.append((field[(z_end - 1) * plane:z_end * plane], rank + 1))
        recvs.append((field[z_end * plane:(z_end + 1) * plane], rank + 1))
    return sends, recvs


def exchange_planes(field, plane, z_begin, z_end, rank, world, group=None):
    """Send the first / last owned plane of ``field`` (flat, x-fastest, local planes) to the lower /
    upper neighbour's ghost plane and receive theirs.  Returns the list of Work handles."""
    sends, recvs = boundary_transfers(field, plane, z_begin, z_end, rank, world)
    ops = []
    for (t, q), (u, _) in zip(sends, recvs):  # per neighbour: send, then receive
        ops.append(dist.P2POp(dist.isend, t, q, group))
        ops.append(dist.P2POp(dist.irecv, u, q, group))
    return dist.batch_isend_irecv(ops) if ops else []


class Level1Slab:
    """One rank's slab of Level 1: explicit sweeps in dwell mode (stepGOMELTDwellTime cF:2617-2664) or with a load
    vector and the clamp (the Level-1 sweeps of stepGOMELT / subcycleGOMELT, cF:2172-2185, 2813-2854)."""

    def __init__(self, gm, props, nodes, h, rank, world, bc5, nz_active=None, n_substrate=0, device=None,
                 symmetric=False, fused=True, parts=None):
        self.gm, self.ops, self.props = gm, gm.ops, props
        self.rank, self.world = rank, world
        nx, ny, nz = (int(v) for v in nodes)
        self.nodes_global = (nx, ny, nz)
        self.plane = nx * ny
        parts = list(parts) if parts is not None else partition_planes(nz, world)
        self.k0, self.k1 = parts[rank]
        self.g0, self.nzl, self.zb, self.ze = local_extent(rank, world, self.k0, self.k1)
        self.grid = gm._lib.make_grid((nx, ny, self.nzl), h)
        self.bc5 = list(bc5)
        self.device = device
        self.set_active(nz if nz_active is None else int(nz_active), n_substrate)
        n = self.plane * self.nzl
        self.symmetric = bool(symmetric) and world > 1 and device is not None
        self.fused = self.symmetric and bool(fused)
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem

            # one symmetric allocation of two buffers, sized for the largest slab so that every rank's layout is
            # the same (buffer b of rank q starts at ptrs[q] + 4 * b * nmax), followed by the counter block
            # (rounded up to 128 bytes: K1's TMA plane ring wants 16-byte aligned field pointers for every buffer)
            self._nmax = -(-self.plane * (max(b - a for a, b in parts) + 2) // 32) * 32
            self._nbuf = 2
            self._nsync = int(gm._lib.load().gomelt_halo_sync_words())
            self._buf = symm_mem.empty(self._nbuf * self._nmax + self._nsync, dtype=torch.float32, device=device)
            self._hdl = symm_mem.rendezvous(self._buf, dist.group.WORLD.group_name)
            self._ptrs = [int(q) for q in self._hdl.buffer_ptrs]
            self._halves = [self._buf[b * self._nmax:b * self._nmax + n] for b in range(self._nbuf)]
            self._sync = self._buf[self._nbuf * self._nmax:].view(torch.int32)
            self._sync.zero_()
            self._seq = 0
            self._cur = 0
            self.T, self.Tn = self._halves[0], self._halves[1]
            # local index of the plane of each neighbour that mirrors my boundary plane
            self._ghost_lo = None if rank == 0 else local_extent(rank - 1, world, *parts[rank - 1])[3]  # its z_end
            self._ghost_hi = None if rank == world - 1 else 0                                            # its plane 0
        else:
            self.T = torch.empty(n, dtype=torch.float32, device=device)
            self.Tn = torch.empty(n, dtype=torch.float32, device=device)
        self.S1 = torch.empty(n, dtype=torch.float32, device=device)
        self.comm = torch.cuda.Stream(device=device) if (device is not None and world > 1 and not self.symmetric) else None
        self.sweeps = 0
        self.transport = None  # dist.Transport: ghost exchanges staged through the host (gloo with device tensors)

    def _exchange(self, field):
        """One blocking ghost exchange of ``field`` (stream-ordered on device tensors)."""
        if self.world == 1:
            return
        if self.transport is not None:
            self.transport.exchange(*boundary_transfers(field, self.plane, self.zb, self.ze, self.rank, self.world))
        else:
            for w in exchange_planes(field, self.plane, self.zb, self.ze, self.rank, self.world):
                w.wait()

    def swap(self):
        """Make the spare buffer the current field (its owned planes were written by the caller)."""
        if self.symmetric:
            self._cur = (self._cur + 1) % self._nbuf
            self.T, self.Tn = self._halves[self._cur], self._halves[(self._cur + 1) % self._nbuf]
        else:
            self.T, self.Tn = self.Tn, self.T

    def unswap(self):
        """Undo the buffer swap of the last sweep: the next sweep starts from the same field again (predictor /
        corrector pairs) and overwrites the last result.  The halo protocol keeps counting sweeps."""
        if self.symmetric:
            self._cur = (self._cur - 1) % self._nbuf
            self.T, self.Tn = self._halves[self._cur], self._halves[(self._cur + 1) % self._nbuf]
        else:
            self.T, self.Tn = self.Tn, self.T

    def refresh_ghosts_T(self):
        """Ghost planes of the current temperature from the neighbours, ordered on the stream (the owned boundary planes
        were modified outside a sweep: clamp, injected child solution)."""
        self._exchange(self.T)

    def set_active(self, nz_active_global, n_substrate_global):
        """tmp_ne_nn / substrate of the whole grid (cF:495-517, 562-579) -> this slab's local planes / node ids."""
        self.nz_active_global = int(nz_active_global)
        self.nz_active = min(max(self.nz_active_global - self.g0, 0), self.nzl)
        self.n_substrate = min(max(int(n_substrate_global) - self.g0 * self.plane, 0), self.nzl * self.plane)
        self.owns_top = self.k0 <= self.nz_active_global - 1 < self.k1

    # ---- state ---------------------------------------------------------------------------
    def owned(self, field):
        return field[self.zb * self.plane:self.ze * self.plane]

    def set_owned(self, T_owned, S1_owned):
        self.owned(self.T).copy_(T_owned)
        self.owned(self.S1).copy_(S1_owned)
        self.fill_ghosts()

    def fill_ghosts(self):
        """Ghost planes of T and S1 from the neighbours (collective); restarts the counter protocol (counters zeroed,
        sequence number 0)."""
        if self.world == 1:
            return
        for f in (self.T, self.S1):
            self._exchange(f)
        if self.device is not None:
            torch.cuda.synchronize(self.device)
        if self.symmetric:
            self._sync.zero_()
            self._seq = 0
            torch.cuda.synchronize(self.device)
            self._hdl.barrier(channel=0)
            torch.cuda.synchronize(self.device)

    # ---- one sweep -----------------------------------------------------------------------
    def _flags(self, rhs, clamp, topflux=None):
        return (self.ops.STEP_BC_CONST | (self.ops.STEP_FUSED_FLUX if topflux is None else 0)
                | (self.ops.STEP_CLAMP if clamp else 0))

    def _k1(self, dt, z0, z1, topflux, rhs=None, clamp=False):
        if z1 <= z0:
            return
        # the surface load of the top active plane (computeConvRadBC) is evaluated inside K1 by the launch
        # that finalises that plane; `topflux` is kept for callers that pre-computed it
        self.ops.level_step(self.props, self.grid, self.T, self.S1, self.Tn, dt, topflux=topflux, rhs=rhs,
                            nz_active=self.nz_active, n_substrate=self.n_substrate,
                            flags=self._flags(rhs, clamp, topflux), bc5=self.bc5, z_range=(z0, z1))

    def _peer(self, q, half, ghost_plane):
        return self._ptrs[q] + 4 * (half * self._nmax + ghost_plane * self.plane)

    def _peer_sync(self, q):
        return self._ptrs[q] + 4 * self._nbuf * self._nmax

    def sweep_fused(self, dt, rhs=None, clamp=False):
        """ONE C-ABI call: the fused level step over the owned planes + the device-side halo exchange (module docstring)."""
        cur = self._cur
        nxt = (cur + 1) % self._nbuf
        lo_rank, hi_rank = self.rank - 1, self.rank + 1
        lo = self._peer(lo_rank, nxt, self._ghost_lo) if self.rank > 0 else None
        hi = self._peer(hi_rank, nxt, self._ghost_hi) if self.rank < self.world - 1 else None
        halo = None
        if self.fused:
            halo = (self._peer_sync(self.rank), self._peer_sync(lo_rank) if lo else None,
                    self._peer_sync(hi_rank) if hi else None, self._seq)
        self.ops.level_step(self.props, self.grid, self._halves[cur], self.S1, self._halves[nxt], dt, rhs=rhs,
                            nz_active=self.nz_active, n_substrate=self.n_substrate,
                            flags=self._flags(rhs, clamp), bc5=self.bc5,
                            z_range=(self.zb, self.ze), peer_lo=lo, peer_hi=hi, halo=halo)
        if self.fused:
            self._seq += 1
        else:
            self._hdl.barrier(channel=0)  # round-1 form: pushed after the step, sweeps ordered by a barrier launch
        self._cur = nxt
        self.T, self.Tn = self._halves[nxt], self._halves[(nxt + 1) % self._nbuf]
        self.sweeps += 1
        return self.T

    def dwell_sweep(self, dt):
        return self.sweep(dt)

    def sweep(self, dt, rhs=None, clamp=False):
        if self.symmetric:
            return self.sweep_fused(dt, rhs, clamp)
        top = None
        zb, ze = self.zb, self.ze
        if self.world == 1:
            self._k1(dt, zb, ze, top, rhs, clamp)
        elif self.transport is not None:   # staged transport: the whole slab, then the ghost planes of the new field
            self._k1(dt, zb, ze, top, rhs, clamp)
            self._exchange(self.Tn)
        else:
            lo = zb + 1 if self.rank > 0 else zb
            hi = ze - 1 if self.rank < self.world - 1 else ze
            hi = max(hi, lo)
            self._k1(dt, zb, lo, top, rhs, clamp)            # first owned plane (needed by the rank below)
            self._k1(dt, hi, ze, top, rhs, clamp)            # last owned plane (needed by the rank above)
            main = torch.cuda.current_stream(self.device)
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                works = exchange_planes(self.Tn, self.plane, zb, ze, self.rank, self.world)
            self._k1(dt, lo, hi, top, rhs, clamp)            # interior overlaps the halo transfer
            with torch.cuda.stream(self.comm):
                for w in works:
                    w.wait()
            main.wait_stream(self.comm)
        self.T, self.Tn = self.Tn, self.T
        self.sweeps += 1
        return self.T
