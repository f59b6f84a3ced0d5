"""Slab-decomposed Level 1 inside the drop-in (BASELINE.json north_star: "Level 1 ... slab-decomposed
across the GPUs of one box ..., the small fine-level windows stay on the rank that owns the laser").

Every rank runs the same driver loop on the same toolpath (the control flow is host arithmetic on the toolpath rows, so
it is identical everywhere).  What differs is who holds what:

  * every rank owns a z-slab of Level 1 (``slab.Level1Slab``: owned planes + one ghost plane per neighbour, halo
    exchange fused with the sweep over NVLink peer memory);
  * ONE rank - the laser owner - additionally holds Levels 2 / 3 / 0 and runs the native steppers
    (csrc/k_steppers.cu).  Its ``Levels[1]["T0"]`` / ``["S1"]`` and the Level-1 load vector are full-size MIRRORS that
    are kept valid only on the Level-1 box under the Level-2 window (every transfer kernel the steppers run - face
    prolongation cF:1598-1620, getNewTprime cF:2060-2099, the correction projections cF:1396-1565, the window shift
    cF:2439-2443, the S1 push cF:2546-2556 - touches Level 1 only there).  Using the mirror rather than a cut-out grid
    keeps every kernel on the very coordinates and indices of the single-GPU run: results are bit-identical.

The steppers reach the slabs through ``gomelt_hier_t.l1_solve`` (include/gomelt_abi.h): every Level-1 solve becomes

      S1 box down (first solve of a call)   owner mirror -> slab owners       updateStateProperties cF:2546-2556
      load box down                         owner mirror -> slab owners       computeSources / T' corrections
      sweep of every slab                   (halo exchange inside)            computeL1Temperature cF:2813-2854
      T box up                              slab owners -> owner mirror       what the children interpolate from

and a stepper call ends with ``finish``: the clamp of stepGOMELT (cF:2360-2362, after the unclamped prolongation), the
box of the final field WITH the injected child solution (cF:2086-2097) down into the owned planes, and a refresh of the
ghost planes.  ``moveEverything`` starts with the T box of the NEW window position going up.  A box intersects one or
two slabs; transfers are point-to-point between the owner and those ranks (``gomelt_patch_copy_f32`` packs / unpacks;
NCCL send / recv, or staged through the host under gloo so that the N>1 protocol can be tested on CPU-only and
single-GPU machines), never a collective over all ranks.  Non-owner ranks run the same sequence of calls with no
windows: they only serve the Level-1 side of it.

Not done: migrating the windows to another rank when the build grows into the next slab (the owner is fixed at set-up:
the rank that owns the top active plane of the first layer; transfers stay correct, only their locality changes).
"""
import numpy as np

from . import _lib, ops
from .slab import Level1Slab, local_extent, partition_active_planes

F32 = np.float32


class Box:
    """Node box [lo, lo + n) of Level 1 (global indices, x, y, z)."""

    def __init__(self, lo, n):
        self.lo = [int(v) for v in lo]
        self.n = [int(v) for v in n]

    def zcut(self, z0, z1):
        """Intersection with the planes [z0, z1) -> Box or None."""
        a, b = max(self.lo[2], z0), min(self.lo[2] + self.n[2], z1)
        if b <= a or self.n[0] <= 0 or self.n[1] <= 0:
            return None
        return Box([self.lo[0], self.lo[1], a], [self.n[0], self.n[1], b - a])

    @property
    def size(self):
        return self.n[0] * self.n[1] * self.n[2]


def footprint_box(child, L1, margin=2):
    """Level-1 nodes that any transfer between ``child`` (a window level) and Level 1 can touch: the parent cells under
    the window's extent, ``margin`` nodes wider on every side, clipped to the grid."""
    lo, n = [], []
    for d in range(3):
        xc = np.asarray(L1["node_coords"][d], np.float64)
        xf = np.asarray(child["node_coords"][d], np.float64)
        h = (xc[-1] - xc[0]) / (xc.size - 1)
        a = int(np.floor((xf.min() - xc[0]) / h)) - margin
        b = int(np.ceil((xf.max() - xc[0]) / h)) + margin
        a, b = max(a, 0), min(b, xc.size - 1)
        if b < a:
            a, b = 0, -1
        lo.append(a)
        n.append(b - a + 1)
    return Box(lo, n)


def index_box(index_vectors):
    """Box of a tensor-product index set of consecutive indices (overlapNodes)."""
    lo = [int(np.asarray(v)[0]) for v in index_vectors]
    n = [int(np.asarray(v).size) for v in index_vectors]
    for v, a, m in zip(index_vectors, lo, n):
        v = np.asarray(v)
        if not np.array_equal(v, np.arange(a, a + m)):
            raise _lib.GomeltError("overlap index set is not a box of consecutive nodes")
    return Box(lo, n)


class Transport:
    """Point-to-point transfers of float32 tensors between ranks of the default process group: NCCL send / recv on
    device tensors; under gloo, device tensors are staged through the host."""

    def __init__(self, dist_module):
        self.dist = dist_module
        self.backend = dist_module.get_backend() if dist_module.is_initialized() else None

    def exchange(self, sends, recvs):
        """sends = [(tensor, dst_rank)], recvs = [(tensor, src_rank)] (contiguous).  Returns when the received data is
        usable on the current stream."""
        if not sends and not recvs:
            return
        d = self.dist
        stage = self.backend == "gloo"
        ops_, post = [], []
        for t, dst in sends:
            buf = t.detach().cpu() if (stage and t.is_cuda) else t
            ops_.append(d.P2POp(d.isend, buf, dst))
        for t, src in recvs:
            if stage and t.is_cuda:
                buf = t.new_empty(t.shape, device="cpu")
                post.append((t, buf))
            else:
                buf = t
            ops_.append(d.P2POp(d.irecv, buf, src))
        for w in d.batch_isend_irecv(ops_):
            w.wait()
        for t, buf in post:
            t.copy_(buf)


class Level1Dist:
    """The Level-1 side of one rank (module docstring)."""

    def __init__(self, gm, Levels, properties, rank, world, owner=None, device=None, symmetric=None, dist_module=None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.gm, self.rank, self.world = gm, int(rank), int(world)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        L1 = Levels[1]
        nx, ny, nz = (int(v) for v in L1["nodes"])
        self.nodes = (nx, ny, nz)
        self.plane = nx * ny
        # slabs balance the planes that are active when the build starts (those up to the first layer); the planes above
        # the powder bed go to the last rank on top of its share (they cost a store per node until the build reaches them)
        z1 = np.asarray(L1["node_coords"][2], F32)
        active0 = int((z1 <= F32(properties.get("layer_height", 0.0)) + F32(1e-5)).sum())
        self.parts = partition_active_planes(nz, active0, self.world)
        self.extents = [local_extent(q, self.world, *self.parts[q]) for q in range(self.world)]  # (g0, nzl, zb, ze)
        self.props = _lib.make_props(properties)
        self.T_amb = float(F32(properties["T_amb"]))
        c = L1["conditions"]
        self.bc5 = [c["y"][0], c["y"][1], c["x"][0], c["x"][1], c["z"][0]]
        if self.world > 1:
            if dist_module is None:
                import torch.distributed as dist_module
            self.tr = Transport(dist_module)
            if symmetric is None:
                symmetric = self.tr.backend == "nccl"
        else:
            self.tr, symmetric = None, False
        self.slab = Level1Slab(gm, self.props, self.nodes, L1["h"], self.rank, self.world, self.bc5, device=self.device,
                               symmetric=bool(symmetric), parts=self.parts)
        if self.world > 1 and not self.slab.symmetric and self.tr.backend == "gloo":
            self.slab.transport = self.tr  # ghost planes through the staged transport as well
        self.slab.T.fill_(self.T_amb)
        self.slab.Tn.fill_(self.T_amb)
        self.slab.S1.zero_()
        self.rhs = None
        n_store = int(L1.get("n_S1_storage", 0))
        self.S1_storage = torch.zeros((n_store, self.slab.S1.numel()), device=self.device, dtype=torch.float32)
        if owner is None:  # the rank that owns the top active plane of the first layer
            top = int((np.asarray(L1["node_coords"][2]) <= F32(0.0) + F32(1e-5)).sum()) - 1
            owner = next(q for q, (a, b) in enumerate(self.parts) if a <= max(top, 0) < b)
        self.owner = int(owner)
        self.is_owner = self.rank == self.owner
        self._s1_box = None   # S1 box to ship at the next solve
        self._cb = None
        self.stats = {"solves": 0, "boxes_down": 0, "boxes_up": 0, "bytes": 0}

    # ---- geometry ---------------------------------------------------------------------------------------------
    def _stored(self, q):
        g0, nzl, _, _ = self.extents[q]
        return g0, g0 + nzl

    def _owned(self, q):
        return self.parts[q]

    def set_active(self, tmp_ne_nn, n_substrate_L1):
        nz_active = int(tmp_ne_nn[1]) // self.plane
        if nz_active != self.slab.nz_active_global or int(n_substrate_L1) != getattr(self, "_nsub", None):
            self._nsub = int(n_substrate_L1)
            self.slab.set_active(nz_active, self._nsub)

    # ---- box transfers ----------------------------------------------------------------------------------------
    def _local(self, field):
        """(tensor or address, dims) of a slab-local array."""
        return field, (self.nodes[0], self.nodes[1], self.slab.nzl)

    def _count(self, box, up):
        self.stats["boxes_up" if up else "boxes_down"] += 1
        self.stats["bytes"] += 4 * box.size

    def down(self, box, mirror, local, owned_only=False):
        """Owner mirror -> every rank's slab-local array ``local`` on ``box`` (its stored planes: owned + ghosts, or the
        owned planes only).  ``mirror`` = tensor / raw address of a full-size Level-1 array (owner), ignored elsewhere."""
        torch = self.torch
        rng = self._owned if owned_only else self._stored
        g0 = self.extents[self.rank][0]
        sends, recvs, unpack = [], [], None
        if self.is_owner:
            for q in range(self.world):
                cut = box.zcut(*rng(q))
                if cut is None:
                    continue
                self._count(cut, False)
                if q == self.rank:
                    ops.patch_copy(mirror, self.nodes, cut.lo, local, self._local(local)[1],
                                   [cut.lo[0], cut.lo[1], cut.lo[2] - g0], cut.n)
                else:
                    buf = torch.empty(cut.size, device=self.device, dtype=torch.float32)
                    ops.patch_copy(mirror, self.nodes, cut.lo, buf, cut.n, [0, 0, 0], cut.n)
                    sends.append((buf, q))
        else:
            cut = box.zcut(*rng(self.rank))
            if cut is not None:
                buf = torch.empty(cut.size, device=self.device, dtype=torch.float32)
                recvs.append((buf, self.owner))
                unpack = (buf, cut)
        if self.tr is not None:
            self.tr.exchange(sends, recvs)
        if unpack is not None:
            buf, cut = unpack
            ops.patch_copy(buf, cut.n, [0, 0, 0], local, self._local(local)[1], [cut.lo[0], cut.lo[1], cut.lo[2] - g0], cut.n)

    def up(self, box, local, mirror):
        """Every rank's OWNED planes of ``local`` on ``box`` -> the owner's mirror."""
        torch = self.torch
        g0 = self.extents[self.rank][0]
        sends, recvs, unpack = [], [], []
        cut = box.zcut(*self._owned(self.rank))
        if self.is_owner:
            if cut is not None:
                self._count(cut, True)
                ops.patch_copy(local, self._local(local)[1], [cut.lo[0], cut.lo[1], cut.lo[2] - g0], mirror, self.nodes, cut.lo,
                               cut.n)
            for q in range(self.world):
                if q == self.rank:
                    continue
                cq = box.zcut(*self._owned(q))
                if cq is None:
                    continue
                self._count(cq, True)
                buf = torch.empty(cq.size, device=self.device, dtype=torch.float32)
                recvs.append((buf, q))
                unpack.append((buf, cq))
        elif cut is not None:
            buf = torch.empty(cut.size, device=self.device, dtype=torch.float32)
            ops.patch_copy(local, self._local(local)[1], [cut.lo[0], cut.lo[1], cut.lo[2] - g0], buf, cut.n, [0, 0, 0], cut.n)
            sends.append((buf, self.owner))
        if self.tr is not None:
            self.tr.exchange(sends, recvs)
        for buf, cq in unpack:
            ops.patch_copy(buf, cq.n, [0, 0, 0], mirror, self.nodes, cq.lo, cq.n)

    # ---- the protocol of one stepper call ------------------------------------------------------------------------
    def begin_call(self, Levels, tmp_ne_nn, substrate, push_S1=True):
        """Start of stepGOMELT / subcycleGOMELT on every rank: active planes, and the box of Level-1 state that the
        stepper is about to overwrite from Level 2 (shipped with the first solve)."""
        self.set_active(tmp_ne_nn, substrate[1])
        self.t_box = footprint_box(Levels[2], Levels[1])
        self.ov_box = index_box(Levels[2]["overlapNodes"])
        self._s1_box = self.ov_box if push_S1 else None
        self._resweep = False

    def solve(self, dt, flags, mirrors=None):
        """One Level-1 solve on every rank.  mirrors = (T0, S1, T_out, rhs) raw addresses on the owner (rhs may be 0).
        Non-owners pass has-rhs through ``flags`` bit 0x20000 (they derive it from the stepper they mirror)."""
        sl = self.slab
        T0m, S1m, Toutm, rhsm = mirrors if mirrors is not None else (None, None, None, None)
        has_rhs = bool(rhsm) if self.is_owner else bool(flags & 0x20000)
        if self._s1_box is not None:
            self.down(self._s1_box, S1m, sl.S1)
            if sl.n_substrate > 0:
                sl.S1[: sl.n_substrate] = 1.0   # updateStateProperties cF:2553-2556 (whole planes)
            self._s1_box = None
        rhs = None
        if has_rhs:
            if self.rhs is None:
                self.rhs = self.torch.zeros_like(sl.S1)
            else:
                self.rhs.zero_()
            self.down(self.t_box, rhsm, self.rhs, owned_only=True)
            rhs = self.rhs
        if self._resweep:   # predictor and corrector both start from the old field (cF:2355 / 2375, 3306 / 3456)
            sl.unswap()
        sl.sweep(float(dt), rhs=rhs, clamp=bool(flags & _lib.STEP_CLAMP))
        self._resweep = True
        self.up(self.t_box, sl.T, Toutm)
        self.stats["solves"] += 1

    def finish(self, clamp_after, final_mirror=None):
        """End of a stepper call: the clamp that stepGOMELT applies after the prolongation, the injected child solution
        (getNewTprime cF:2086-2097) down into the owned planes, ghost planes refreshed."""
        sl = self.slab
        if clamp_after:
            ops.clamp_min(sl.owned(sl.T), self.T_amb)
        self.down(self.ov_box, final_mirror, sl.T, owned_only=True)
        sl.refresh_ghosts_T()
        self._resweep = False

    def dwell(self, dt, tmp_ne_nn, substrate):
        """stepGOMELTDwellTime cF:2617-2664 on every rank: one sweep, no boxes."""
        self.set_active(tmp_ne_nn, substrate[1])
        self.slab.sweep(float(dt))
        self.stats["solves"] += 1

    # ---- hook for the native steppers (owner) ------------------------------------------------------------------------
    def hook(self):
        """(function pointer, keep-alive) for gomelt_hier_t.l1_solve."""
        if self._cb is None:
            def cb(user, T0, S1, Tout, dt, rhs, flags):
                try:
                    self.solve(dt, flags, mirrors=(T0, S1, Tout, rhs))
                    return 0
                except Exception as exc:  # surfaced by the stepper as a non-zero return
                    self._hook_error = exc
                    return -99

            self._cb = _lib.L1_SOLVE_FN(cb)
        return self._cb

    # ---- whole-field helpers (set-up, layer change, output) ---------------------------------------------------------
    def gather(self, field=None):
        """Full Level-1 field on the owner (None elsewhere): output / tests."""
        torch = self.torch
        sl = self.slab
        field = sl.T if field is None else field
        full = torch.empty(self.plane * self.nodes[2], device=self.device, dtype=torch.float32) if self.is_owner else None
        self.up(Box([0, 0, 0], self.nodes), field, full)
        return full

    def scatter(self, full, field):
        """Owner's full-size array -> every rank's stored planes of ``field``."""
        self.down(Box([0, 0, 0], self.nodes), full, field)

    def layer_shift(self, L1, new_coords, state_idx):
        """gm:215-224 on the slabs: T0 <- max(I(T0) at the raised planes, T_amb); S1 <-> S1_storage rotation.  The
        interpolation runs on the GLOBAL coordinate arrays with the field pointer moved back by the slab's first stored
        plane, so that cell indices and weights are those of the single-GPU run; a slab reads at most one plane above
        its owned range (the planes rise by one layer < h_z), which is its ghost plane."""
        torch = self.torch
        sl = self.slab
        cf = self.gm.computeFunctions
        g0, nzl, zb, ze = self.extents[self.rank]
        # (host check of that claim: the parent cell of every owned target plane lies within the stored planes)
        zc = np.asarray(L1["node_coords"][2], np.float64)
        zt = np.asarray(new_coords[2], np.float64)[sl.k0:sl.k1]
        cell = np.clip(np.floor((zt - zc[0]) / (zc[1] - zc[0])), 0, zc.size - 2).astype(int)
        if cell.size and (cell.min() < g0 or cell.max() + 1 > g0 + nzl - 1):
            raise _lib.GomeltError("layer shift reaches beyond the ghost planes of this slab")
        src = cf._coords(L1["node_coords"])
        tgt = cf._coords([new_coords[0], new_coords[1], np.asarray(new_coords[2], F32)[sl.k0:sl.k1]])
        out = sl.owned(sl.Tn)
        ops.interp(src, sl.T, tgt, out, u_offset=g0 * self.plane)
        ops.clamp_min(out, self.T_amb)
        sl.swap()
        sl.fill_ghosts()
        if self.S1_storage.shape[0] > 0:
            self.S1_storage[state_idx - 1, :] = sl.S1
            sl.S1.copy_(self.S1_storage[state_idx, :])
