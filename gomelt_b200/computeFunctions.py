"""Drop-in ``computeFunctions`` namespace of the B200 build: the names the reference driver reaches
through ``from computeFunctions import *`` (gm:10), with the same positional arguments,
running on ``cuda`` through the C ABI of ``libgomelt_sm100.so``.

State lives on the device: ``Levels[i]["T0" | "Tprime0" | "S1" | "S2"]`` are torch CUDA tensors
(float32, S2 bool); geometry (``node_coords``, overlap index sets, bounds) is host NumPy.  Any field
may also be handed in as a NumPy array (it is uploaded), which is how the parity tests drive it.

What differs from the reference *by construction* (results agree to float32 rounding):
  * ``Shapes`` / ``LInterp`` are small descriptors, not (n,8) / (nef,8,8) operator arrays: every
    transfer weight is recomputed in registers (csrc/k_transfer.cu);
  * k and rho*cp are evaluated inside the fused level step from (T0, S1) - the arrays the reference
    passes around as ``Lk`` / ``Lrhocp`` are only materialised where a correction term integrates them;
  * the corrector sweep of Level 3 in ``stepGOMELT`` re-uses the predictor interior (identical inputs,
    cF:2199-2201 vs 2355/2375) and only re-applies the Dirichlet faces;
  * ``stepGOMELT`` / ``subcycleGOMELT`` / ``stepGOMELTDwellTime`` are ONE native call each (csrc/k_steppers.cu) and
    update the state tensors of ``Levels`` IN PLACE (the reference returns new arrays): keep a ``.clone()`` if you need
    the field of an earlier step;
  * the subcycle histories ``L2all, L3all, L3pall`` (returned and dropped by the driver, gm:437) are
    returned as ``None``.
There is no CPU path: every entry point raises ``GomeltError`` without the library or a GPU.
"""
import copy  # noqa: F401  (re-exported: gm:200 uses ``copy`` through the star import)
import math
import os
import sys

import numpy as np

from . import _lib, levels, ops, schema
from .schema import SetupNonmesh, SetupProperties, getStaticSubcycle  # noqa: F401
from .toolpath import count_lines, parsingGcode  # noqa: F401
from .output import saveFinalResult, saveResult, saveResults, saveResultsFinal, saveState  # noqa: F401

F32 = np.float32


def _torch():
    return _lib.require_cuda()


def _dev(x, dtype=None):
    """Field -> contiguous CUDA tensor (uploads NumPy input)."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda()
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x))).cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _f(x):
    return _dev(x, _torch().float32)


def _host(x):
    torch = _torch()
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _CoordCache:
    """Device copies of 1-D coordinate / index arrays, keyed by content (they are tiny).  A new array goes up through a
    pinned staging ring with an asynchronous copy on the current stream: a plain ``.cuda()`` of pageable memory is a
    synchronous copy, i.e. it waits for everything queued on the stream - a window move right after a stepper call would
    stall the host for the whole block and leave the GPU idle while the next block is being issued (~40 small uploads per
    move to a new position: 7.9 -> 6.x ms per C2 subcycle block)."""
    RING = 4 << 20

    def __init__(self):
        self.store = {}
        self.pin = None
        self.off = 0

    def _upload(self, a):
        torch = _torch()
        t = torch.from_numpy(a)
        if a.nbytes == 0 or a.nbytes > self.RING // 8:
            return t.cuda()
        if self.pin is None:
            self.pin = torch.empty(self.RING, dtype=torch.uint8).pin_memory()
        off = (self.off + 63) // 64 * 64
        if off + a.nbytes > self.RING:   # wrap: every copy issued out of the ring must have completed (rare)
            torch.cuda.current_stream().synchronize()
            off = 0
        stage = self.pin[off:off + a.nbytes].view(t.dtype)
        stage.copy_(t.reshape(-1))
        self.off = off + a.nbytes
        dev = torch.empty(t.shape, dtype=t.dtype, device="cuda")
        dev.reshape(-1).copy_(stage, non_blocking=True)
        return dev

    def get(self, arr, dtype):
        a = np.ascontiguousarray(np.asarray(_host(arr)).astype(dtype))
        key = (a.dtype.str, a.tobytes())
        t = self.store.get(key)
        if t is None:
            if len(self.store) > 4096:
                self.store.clear()
            t = self._upload(a)
            self.store[key] = t
        return t


_CACHE = _CoordCache()


def _coords(c3):
    return [_CACHE.get(c, np.float32) for c in c3]


def _index3(i3):
    return [_CACHE.get(i, np.int32) for i in i3]


def _props(properties):
    p = properties.get("_gomelt_props") if isinstance(properties, dict) else None
    if p is None:
        p = _lib.make_props(properties)
        if isinstance(properties, dict):
            properties["_gomelt_props"] = p
    return p


def _grid(L):
    return _lib.make_grid(L["nodes"], L["h"])


# ----------------------------------------------------------------------------------------------
# setup (host geometry + device fields)
# ----------------------------------------------------------------------------------------------
_DIST_CONFIG = None


def enable_distributed(rank, world, owner=None, symmetric=None):
    """Slab-decompose Level 1 over the ranks of the default torch.distributed process group for every ``Levels`` built
    from now on (dist.py; one process per GPU, every rank runs the same driver).  ``enable_distributed(0, 1)`` runs the
    same machinery on one rank; ``disable_distributed()`` turns it off."""
    global _DIST_CONFIG
    _DIST_CONFIG = {"rank": int(rank), "world": int(world), "owner": owner, "symmetric": symmetric}


def disable_distributed():
    global _DIST_CONFIG
    _DIST_CONFIG = None


def distOf(Levels):
    """The Level1Dist of a distributed ``Levels`` list, else None."""
    return Levels[0].get("_gomelt_dist") if isinstance(Levels[0], dict) else None


def isWorker(Levels):
    """True on the ranks that hold only a Level-1 slab (no windows, no Level 0)."""
    d = distOf(Levels)
    return d is not None and not d.is_owner


def gatherL1(Levels, what="T0"):
    """Full Level-1 temperature / state on the laser owner (None on the other ranks); collective."""
    d = distOf(Levels)
    if d is None:
        return Levels[1][what]
    return d.gather(d.slab.T if what == "T0" else d.slab.S1)


def outputView(Levels, with_storage=False):
    """What the file-output hooks of a distributed run see (collective; None on the ranks without windows): a shallow
    copy of ``Levels`` whose Level-1 temperature / state - and, for checkpoints, ``S1_storage`` - are assembled from the
    slabs on the laser owner instead of its mirrors (which are valid under the windows only)."""
    d = distOf(Levels)
    if d is None:
        return Levels
    T = d.gather(d.slab.T)
    S = d.gather(d.slab.S1)
    rows = [d.gather(d.S1_storage[i].contiguous()) for i in range(d.S1_storage.shape[0])] if with_storage else None
    if not d.is_owner:
        return None
    view = list(Levels)
    view[1] = dict(Levels[1], T0=T, S1=S)
    if rows is not None:
        view[1]["S1_storage"] = _torch().stack(rows) if rows else _torch().zeros((0, T.numel()), device="cuda")
    view[0] = {k: v for k, v in Levels[0].items() if not k.startswith("_gomelt")}
    return view


def adoptCheckpoint(Levels, loaded):
    """Restart of a distributed run (gm:108-129): ``loaded`` = the Levels list of a checkpoint (single-GPU format: whole
    Level-1 fields; every rank has read it), ``Levels`` = the freshly set-up distributed list.  Every rank cuts its slab
    (ghost planes included) out of the whole fields, the laser owner keeps the rest, the other ranks drop it."""
    torch = _torch()
    d = distOf(Levels)
    sl = d.slab
    g0, nzl, _, _ = d.extents[d.rank]
    lo, hi = g0 * d.plane, (g0 + nzl) * d.plane
    sl.T.copy_(_f(loaded[1]["T0"])[lo:hi])
    sl.S1.copy_(_f(loaded[1]["S1"])[lo:hi])
    st = loaded[1].get("S1_storage")
    if st is not None and d.S1_storage.shape[0] > 0:
        d.S1_storage.copy_(_f(st).reshape(d.S1_storage.shape[0], -1)[:, lo:hi])
    sl.fill_ghosts()   # (collective: restarts the halo protocol on the new field)
    loaded[0]["_gomelt_dist"] = d
    loaded[1]["S1_storage"] = None
    if d.is_owner:
        loaded[1]["T0"], loaded[1]["S1"] = _f(loaded[1]["T0"]), _f(loaded[1]["S1"])   # the mirrors: valid everywhere now
    else:
        for i in (1, 2, 3):
            loaded[i]["T0"] = loaded[i]["S1"] = loaded[i]["S2"] = None
        for i in (2, 3):
            loaded[i]["Tprime0"] = None
        loaded[0]["S1"] = loaded[0]["S2"] = None
    torch.cuda.synchronize()
    return loaded


def layerShiftL1(Levels, tmp_coords, state_idx, T_amb):
    """gm:215-224 for a slab-decomposed Level 1 (dist.Level1Dist.layer_shift)."""
    distOf(Levels).layer_shift(Levels[1], tmp_coords, state_idx)
    Levels[1]["node_coords"] = copy.deepcopy(tmp_coords)
    return Levels


def _setup_distributed(L, properties):
    import gomelt_b200 as gm

    from . import dist as _dist

    torch = _torch()
    cfg = _DIST_CONFIG
    d = _dist.Level1Dist(gm, L, properties, cfg["rank"], cfg["world"], owner=cfg["owner"], symmetric=cfg["symmetric"])
    L[0]["_gomelt_dist"] = d
    L[1]["S1_storage"] = None  # slab-local: d.S1_storage
    if d.is_owner:
        nn1 = L[1]["nn"]
        # full-size mirrors, valid on the box under the Level-2 window (uniform at the start: valid everywhere)
        L[1]["T0"] = torch.full((nn1,), float(F32(properties["T_amb"])), device="cuda", dtype=torch.float32)
        L[1]["S1"] = torch.zeros(nn1, device="cuda", dtype=torch.float32)
        for i in (2, 3):
            nn = L[i]["nn"]
            L[i]["T0"] = torch.full((nn,), float(F32(properties["T_amb"])), device="cuda", dtype=torch.float32)
            L[i]["S1"] = torch.zeros(nn, device="cuda", dtype=torch.float32)
            L[i]["S2"] = torch.zeros(nn, device="cuda", dtype=torch.bool)
            L[i]["Tprime0"] = torch.zeros(nn, device="cuda", dtype=torch.float32)
        L[0]["S1"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.float32)
        L[0]["S2"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.bool)
    else:
        for i in (1, 2, 3):
            L[i]["T0"] = L[i]["S1"] = L[i]["S2"] = None
        for i in (2, 3):
            L[i]["Tprime0"] = None
        L[0]["S1"] = L[0]["S2"] = None
    return L


def SetupLevels(solver_input, properties):
    """cF:115-264.  Unused reference fields (T, k, rhocp, Tprime: cF:150-157, 172) are not allocated."""
    torch = _torch()
    L = levels.build_levels(solver_input, properties)
    if _DIST_CONFIG is not None:
        return _setup_distributed(L, properties)
    for i in (1, 2, 3):
        nn = L[i]["nn"]
        L[i]["T0"] = torch.full((nn,), float(F32(properties["T_amb"])), device="cuda", dtype=torch.float32)
        L[i]["S1"] = torch.zeros(nn, device="cuda", dtype=torch.float32)
        L[i]["S2"] = torch.zeros(nn, device="cuda", dtype=torch.bool)
    L[1]["S1_storage"] = torch.zeros((L[1]["n_S1_storage"], L[1]["nn"]), device="cuda", dtype=torch.float32)
    for i in (2, 3):
        L[i]["Tprime0"] = torch.zeros(L[i]["nn"], device="cuda", dtype=torch.float32)
    L[0]["S1"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.float32)
    L[0]["S2"] = torch.zeros(L[0]["nn"], device="cuda", dtype=torch.bool)
    return L


getStaticNodesAndElements = levels.static_sizes
calcStaticTmpNodesAndElements = levels.active_sizes
getSubstrateNodes = levels.substrate_counts


def getOverlapRegion(node_coords, nx, ny):
    return levels.overlap_ids(node_coords, nx, ny)


# ----------------------------------------------------------------------------------------------
# interpolation entry points
# ----------------------------------------------------------------------------------------------
def interpolatePointsMatrix(Level, node_coords_new):
    """cF:1028-1107.  The reference returns (n,8) weights + indices; here a descriptor: the weights are
    recomputed in registers whenever the pair is used (``LInterp`` is opaque to the driver, gm:74-75)."""
    return {"src_coords": [np.array(_host(c), F32) for c in Level["node_coords"]],
            "tgt_coords": [np.array(_host(c), F32) for c in node_coords_new]}


def interpolatePoints(Level, u, node_coords_new):
    """cF:1131-1210 -> CUDA tensor of length nx'*ny'*nz'."""
    torch = _torch()
    tgt = _coords(node_coords_new)
    out = torch.empty(tgt[0].numel() * tgt[1].numel() * tgt[2].numel(), device="cuda", dtype=torch.float32)
    return ops.interp(_coords(Level["node_coords"]), _f(u), tgt, out)


def interpolate_w_matrix(C2F, T):
    """cF:1110-1128 on a descriptor from interpolatePointsMatrix."""
    torch = _torch()
    tgt = _coords(C2F["tgt_coords"])
    out = torch.empty(tgt[0].numel() * tgt[1].numel() * tgt[2].numel(), device="cuda", dtype=torch.float32)
    return ops.interp(_coords(C2F["src_coords"]), _f(T), tgt, out)


# ----------------------------------------------------------------------------------------------
# per-level pieces
# ----------------------------------------------------------------------------------------------
def computeStateProperties(T, S1, properties, n_substrate):
    """cF:2567-2614 -> (S1 f32, S2 bool, k, rhocp) as CUDA tensors."""
    torch = _torch()
    T, S1 = _f(T), _f(S1)
    S1o, k, rc = torch.empty_like(T), torch.empty_like(T), torch.empty_like(T)
    S2o = torch.empty(T.numel(), device="cuda", dtype=torch.bool)
    ops.state_props(_props(properties), T, S1, int(n_substrate), S1_out=S1o, S2_out=S2o, k_out=k, rhocp_out=rc)
    return S1o, S2o, k, rc


def computeConvRadBC(Level, LevelT0, ne, nn, properties, F):
    """cF:2207-2301 with the reference's signature: convection + radiation + evaporation load of the top face of the
    elements [ne - ne_x*ne_y, ne) added to ``F`` (exact-libm stand-alone kernel; the steppers use K1's fused
    epilogue).  Returns a new [nn] CUDA tensor."""
    nx, ny = int(Level["nodes"][0]), int(Level["nodes"][1])
    nz_active = int(ne) // ((nx - 1) * (ny - 1)) + 1
    out = _f(F).clone()
    top = out[(nz_active - 1) * nx * ny: nz_active * nx * ny]
    ops.surface_flux(_props(properties), _grid(Level), _f(LevelT0), top, nz_active=nz_active, add=True)
    return out


def _nz_active(L, tmp_nn):
    return int(tmp_nn) // (L["nodes"][0] * L["nodes"][1])


def _projected_source(L3, parent, rows, powers, properties, F=None):
    """computeSources cF:928-988 (one row) / computeLevelSource cF:2667-2730 (mean over rows): the laser
    source integrated at Level-3 Gauss points, projected on ``parent`` -> [nn_parent] load vector (two launches for any
    number of rows: gomelt_projected_source_f32)."""
    torch = _torch()
    nx, ny, nz = (int(v) for v in parent["nodes"])
    accumulate = F is not None
    if F is None:
        F = torch.empty(nx * ny * nz, device="cuda", dtype=torch.float32)
    n = len(powers)
    r = np.zeros((n, 7), F32)
    r[:, :3] = np.asarray([np.asarray(_host(q), F32)[:3] for q in rows], F32)
    r[:, 6] = np.asarray(_host(powers), F32)
    h3 = L3["h"]
    wq = F32(F32(F32(h3[0]) * F32(h3[1])) * F32(h3[2])) * F32(0.125)
    tables = torch.empty(n * (nx + ny + nz), device="cuda", dtype=torch.float32)
    return ops.projected_source(_props(properties), _coords(L3["node_coords"]), _coords(parent["node_coords"]), float(wq), r,
                                tables, F, accumulate=accumulate)


_PAIR_CACHE = {}


def _pair_cells(fine, parent):
    """Group the fine elements by parent cell (the role of Shapes[.][2] / the node set of Gauss point 0,
    cF:1351): per axis, parent cell of each fine element by the reference's floor rule.  Memoised by the
    content of the six coordinate arrays: the windows revisit the same integer shifts row after row."""
    key = tuple(np.asarray(L["node_coords"][d], F32).tobytes() for L in (fine, parent) for d in range(3))
    hit = _PAIR_CACHE.get(key)
    if hit is None:
        if len(_PAIR_CACHE) > 512:
            _PAIR_CACHE.clear()
        hit = _PAIR_CACHE[key] = _pair_cells_build(fine, parent)
    return hit


class _CellSum:
    """Lazily allocated, size-shared scratch [ncell * 8] of gomelt_project_f32 (``.data_ptr()`` like a tensor)."""
    _pool = {}

    def __init__(self, n):
        self.n = int(n)

    def tensor(self):
        t = _CellSum._pool.get(self.n)
        if t is None:
            if len(_CellSum._pool) > 8:
                _CellSum._pool.clear()
            t = _CellSum._pool[self.n] = _torch().empty(self.n, device="cuda", dtype=_torch().float32)
        return t

    def data_ptr(self):
        return self.tensor().data_ptr()


def _pair_cells_build(fine, parent):
    torch = _torch()
    g = F32(0.57735026918962576)
    lo, hi = F32(0.5) * (F32(1) + g), F32(0.5) * (F32(1) - g)
    cell0, ncell, first, hint, wtab, rmax, off = [], [], [], 1, [], [], []
    for d in range(3):
        xf = np.asarray(fine["node_coords"][d], F32)
        xc = np.asarray(parent["node_coords"][d], F32)
        xq0 = lo * xf[:-1] + hi * xf[1:]
        xq1 = hi * xf[:-1] + lo * xf[1:]
        hc = xc[1] - xc[0]
        ec = np.clip(np.floor((xq0 - xc[0]) / hc).astype(np.int64), 0, xc.size - 2)
        c0, c1 = int(ec[0]), int(ec[-1])
        cell0.append(c0)
        ncell.append(c1 - c0 + 1)
        fs = np.searchsorted(ec, np.arange(c0, c1 + 2), side="left").astype(np.int32)
        first.append(_CACHE.get(fs, np.int32))
        hint *= int(np.diff(fs).max())
        rmax.append(int(np.diff(fs).max()))
        # nested grouping (gomelt_project_args_t.uniform_off): cell i holds the elements [i r - off, (i + 1) r - off)
        r_d, n_el = rmax[-1], xf.size - 1
        o_d = r_d - int(fs[1]) if fs.size > 2 else 0
        want = np.clip(np.arange(fs.size, dtype=np.int64) * r_d - o_d, 0, n_el)
        off.append(o_d if (0 <= o_d < r_d and np.array_equal(want, fs)) else -1)
        # parent shape-function factors at both Gauss points of every fine element, each in the parent cell that holds
        # that Gauss point (cF:1283-1335): (x1 - xq0, xq0 - x0, x1 - xq1, xq1 - x0) - see gomelt_project_args_t.wtab_*
        e1 = np.clip(np.floor((xq1 - xc[0]) / hc).astype(np.int64), 0, xc.size - 2)
        tab = np.stack([xc[ec + 1] - xq0, xq0 - xc[ec], xc[e1 + 1] - xq1, xq1 - xc[e1]], axis=1).astype(F32)
        wtab.append(_CACHE.get(tab.reshape(-1), np.float32))
    # (the per-cell scratch of a stand-alone projection is shared by size: the steppers carve theirs from the work block,
    # and a cached grouping per window position must not pin 41 MB each at C2 size)
    return {"cell0": cell0, "ncell": ncell, "first": first, "hint": hint, "cellsum": _CellSum(ncell[0] * ncell[1] * ncell[2] * 8),
            "wtab": wtab, "rmax": rmax,
            "off": off,
            "hf": [float(F32(v)) for v in fine["h"]], "hc": [float(F32(v)) for v in parent["h"]],
            "fine": _coords(fine["node_coords"]), "parent": _coords(parent["node_coords"])}


def _project(cells, A, coef, V, mode, scale=1.0, A2=None, **kw):
    return ops.project(cells["fine"], cells["parent"], A, coef, V, cells, mode=mode, scale=scale, A2=A2,
                       accumulate=True, **kw)


def _bc5(L1):
    c = L1["conditions"]
    return [c["y"][0], c["y"][1], c["x"][0], c["x"][1], c["z"][0]]


def getNewTprime(Fine, FineT0, CoarseT, Coarse, C2F=None):
    """cF:2060-2099: inject the fine solution into the parent's overlap nodes (in place on ``CoarseT``),
    then T' = T_fine - I(parent).  Returns (Tprime, CoarseT)."""
    torch = _torch()
    FineT0, CoarseT = _f(FineT0), _f(CoarseT)
    fc, cc = _coords(Fine["node_coords"]), _coords(Coarse["node_coords"])
    ops.interp(fc, FineT0, _coords(Fine["overlapCoords"]), CoarseT,
               index_map=(*_index3(Fine["overlapNodes"]), Coarse["nodes"][0], Coarse["nodes"][1]))
    Tprime = torch.empty_like(FineT0)
    ops.interp(cc, CoarseT, fc, Tprime, mode=_lib.INTERP_RSUB, base=FineT0)
    return Tprime, CoarseT


def getBothNewTprimes(Levels, FineT, MesoT, M2F, CoarseT, C2M):
    """cF:2102-2132."""
    lTp, mT0 = getNewTprime(Levels[3], FineT, MesoT, Levels[2])
    mTp, uT0 = getNewTprime(Levels[2], mT0, CoarseT, Levels[1])
    return lTp, mTp, mT0, uT0


class BoxIndex:
    """``Levels[0]["idx"]`` / ``["idx_L2"]``: the flat Level-0 ids of a window's nodes (getOverlapRegion cF:1642-1669)
    as the three index vectors they are the tensor product of.  The ids themselves (one int64 per window node, rebuilt
    on every moveEverything in the reference) are materialised only when somebody asks for an array (``np.asarray``);
    the device side gathers / scatters through the vectors (gomelt_box_copy)."""

    def __init__(self, vectors, nx, ny):
        self.vectors = [np.ascontiguousarray(np.asarray(v), dtype=np.int32) for v in vectors]
        self.nx, self.ny = int(nx), int(ny)
        self._ids = None

    @property
    def shape(self):
        return (self.vectors[0].size * self.vectors[1].size * self.vectors[2].size,)

    @property
    def size(self):
        return self.shape[0]

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        if self._ids is None:
            self._ids = levels.overlap_ids(self.vectors, self.nx, self.ny)
        return self._ids if dtype is None else self._ids.astype(dtype)

    def device(self):
        return _index3(self.vectors)


def take_box(a, idx):
    """a[idx] for a BoxIndex (float32 / bool / uint8 CUDA tensor) -> new tensor."""
    torch = _torch()
    out = torch.empty(idx.size, device="cuda", dtype=a.dtype)
    return ops.box_copy(a, out, idx.device(), idx.nx, idx.ny, scatter=False)


def put_box(a, idx, v):
    """a[idx] = v in place."""
    return ops.box_copy(v.to(a.dtype).contiguous(), a, idx.device(), idx.nx, idx.ny, scatter=True)


def _ensure_fields(Levels):
    torch = _torch()
    for i in (1, 2, 3):
        Levels[i]["T0"], Levels[i]["S1"] = _f(Levels[i]["T0"]), _f(Levels[i]["S1"])
    for i in (2, 3):
        Levels[i]["Tprime0"] = _f(Levels[i]["Tprime0"])
    Levels[3]["S2"] = _dev(Levels[3]["S2"], torch.bool)
    Levels[0]["S1"], Levels[0]["S2"] = _f(Levels[0]["S1"]), _dev(Levels[0]["S2"], torch.bool)


class _Workspace:
    """Device scratch that belongs to one ``Levels`` list: the work block of the native steppers, the second Level-1
    temperature buffer (the new part-scale field is left there and the handles are swapped: no copy), and the second
    T0 / T'0 buffers of the windows that moveEverything writes into."""

    def __init__(self):
        self.work = None
        self.l1_spare = None
        self.alt = {}

    def buffer(self, key, like):
        t = self.alt.get(key)
        if t is None or t.numel() != like.numel() or t.dtype != like.dtype or t.data_ptr() == like.data_ptr():
            t = self.alt[key] = _torch().empty_like(like)
        return t

    def spare_for(self, T1):
        t = self.l1_spare
        if t is None or t.numel() != T1.numel() or t.data_ptr() == T1.data_ptr():
            t = self.l1_spare = _torch().empty_like(T1)
        return t

    def work_for(self, nfloats):
        if self.work is None or self.work.numel() < nfloats:
            self.work = _torch().empty(int(nfloats), device="cuda", dtype=_torch().float32)
        return self.work


def _workspace(Levels):
    ws = Levels[0].get("_gomelt_ws")
    if ws is None:
        ws = Levels[0]["_gomelt_ws"] = _Workspace()
    return ws


def _fill_level(dst, L, n_substrate, with_state=True):
    nx, ny, nz = (int(v) for v in L["nodes"])
    dst.grid = _lib.make_grid((nx, ny, nz), L["h"])
    c = _coords(L["node_coords"])
    dst.x, dst.y, dst.z = (t.data_ptr() for t in c)
    dst.T0, dst.S1 = L["T0"].data_ptr(), L["S1"].data_ptr()
    if with_state:
        dst.Tprime0 = L["Tprime0"].data_ptr()
    dst.n_substrate = int(n_substrate)
    return c


def _fill_pair(dst, cells):
    for d in range(3):
        dst.cell0[d], dst.ncell[d] = int(cells["cell0"][d]), int(cells["ncell"][d])
    dst.first_x, dst.first_y, dst.first_z = (t.data_ptr() for t in cells["first"])
    dst.elems_per_cell_hint = int(cells["hint"])
    dst.wtab_x, dst.wtab_y, dst.wtab_z = (t.data_ptr() for t in cells["wtab"])
    for d in range(3):
        dst.rmax[d] = int(cells["rmax"][d])
        dst.uniform_off[d] = int(cells["off"][d])


def _fill_overlap(dst, L):
    ix = _index3(L["overlapNodes"])
    cx = _coords(L["overlapCoords"])
    dst.ix, dst.iy, dst.iz = (t.data_ptr() for t in ix)
    dst.cx, dst.cy, dst.cz = (t.data_ptr() for t in cx)
    for d in range(3):
        dst.n[d] = int(ix[d].numel())
    return ix, cx


def _hier(Levels, Shapes, tmp_ne_nn, substrate, properties, N2=1, N3=1, windows=True):
    """gomelt_hier_t of the current ``Levels`` (+ the tensors that must stay alive through the call)."""
    ws = _workspace(Levels)
    L0, L1, L2, L3 = Levels
    h = _lib.Hier()
    keep = [_fill_level(h.L1, L1, substrate[1], with_state=False)]
    h.bc5 = (_lib.C.c_float * 5)(*[float(v) for v in _bc5(L1)])
    h.nz_active_L1 = _nz_active(L1, tmp_ne_nn[1])
    spare = ws.spare_for(L1["T0"])
    h.L1_spare = spare.data_ptr()
    if windows:
        keep.append(_fill_level(h.L2, L2, substrate[2]))
        keep.append(_fill_level(h.L3, L3, substrate[3]))
        h.L3.S2 = L3["S2"].data_ptr()
        _fill_pair(h.L2L1, Shapes["L2L1"])
        _fill_pair(h.L3L1, Shapes["L3L1"])
        _fill_pair(h.L3L2, Shapes["L3L2"])
        keep.append(_fill_overlap(h.ov2, L2))
        keep.append(_fill_overlap(h.ov3, L3))
        h.L0_S1, h.L0_S2 = L0["S1"].data_ptr(), L0["S2"].data_ptr()
        h.L0_nx, h.L0_ny, h.L0_nz = (int(v) for v in L0["nodes"])
        i3 = _index3(L0["overlapNodes"])
        h.l0_ix, h.l0_iy, h.l0_iz = (t.data_ptr() for t in i3)
        keep.append(i3)
        # Level-0 S2 is zero outside the index set of the last scatter into THIS tensor: clear that set, not the grid
        prev = getattr(ws, "l0_prev", None)
        if prev is not None and prev[0] == L0["S2"].data_ptr():
            h.l0p_ix, h.l0p_iy, h.l0p_iz = (t.data_ptr() for t in prev[1])
            for dd in range(3):
                h.l0p_n[dd] = int(prev[1][dd].numel())
            keep.append(prev[1])
        ws.l0_prev = (L0["S2"].data_ptr(), i3)
        need = ops.hier_work_floats(h, N2, N3)
    else:
        need = 0 if spare is not None else int(L1["nn"])
    work = ws.work_for(max(need, 64))
    h.work, h.work_floats = work.data_ptr(), int(work.numel())
    d = distOf(Levels)
    if d is not None:  # slab-decomposed Level 1: the steppers' Level-1 solves go through the hook (dist.py)
        cb = d.hook()
        h.l1_solve = _lib.C.cast(cb, _lib.C.c_void_p)
        keep.append(cb)
    return h, keep, ws


def _native(d, fn, *args):
    """A native stepper call; an exception raised inside the Level-1 hook (dist.py) is re-raised here."""
    if d is not None:
        d._hook_error = None
    try:
        return fn(*args)
    except _lib.GomeltError:
        if d is not None and getattr(d, "_hook_error", None) is not None:
            raise d._hook_error
        raise


def _swap_l1(Levels, ws, in_spare):
    if in_spare:
        Levels[1]["T0"], ws.l1_spare = ws.l1_spare, Levels[1]["T0"]


# ----------------------------------------------------------------------------------------------
# step orchestrators: one native call each (csrc/k_steppers.cu); state is updated IN PLACE in the Levels' tensors
# ----------------------------------------------------------------------------------------------
def stepGOMELT(Levels, ne_nn, tmp_ne_nn, Shapes, LInterp, v, properties, dt, laserP, substrate):
    """cF:2304-2397: one single-step predictor / corrector update of Levels 1-3 (gomelt_step_f32)."""
    torch = _torch()
    d = distOf(Levels)
    if d is not None:
        d.begin_call(Levels, tmp_ne_nn, substrate)
        if not d.is_owner:  # the Level-1 side of the two solves of cF:2355 / 2375
            flags = _lib.STEP_BC_CONST | _lib.STEP_FUSED_FLUX | _lib.L1_CLAMP_AFTER | 0x20000
            d.solve(float(F32(_host(dt))), flags)
            d.solve(float(F32(_host(dt))), flags)
            d.finish(clamp_after=True)
            return Levels, None
    _ensure_fields(Levels)
    row = np.zeros(7, F32)
    row[:3] = np.asarray(_host(v), F32)[:3]
    row[5], row[6] = F32(_host(dt)), F32(_host(laserP))
    h, keep, ws = _hier(Levels, Shapes, tmp_ne_nn, substrate, properties)
    resetmask = torch.empty(int(Levels[3]["nn"]), device="cuda", dtype=torch.bool)
    _swap_l1(Levels, ws, _native(d, ops.step, _props(properties), h, row, resetmask))
    if d is not None:
        d.finish(clamp_after=True, final_mirror=Levels[1]["T0"])
    return Levels, resetmask


def stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, properties, dt, substrate):
    """cF:2617-2664: Level 1 only, no clamp (gomelt_dwell_step_f32)."""
    d = distOf(Levels)
    if d is not None:  # every rank sweeps its slab; the owner's mirror is refreshed by the next moveEverything
        d.dwell(float(F32(_host(dt))), tmp_ne_nn, substrate)
        return Levels
    L1 = Levels[1]
    L1["T0"], L1["S1"] = _f(L1["T0"]), _f(L1["S1"])
    h, keep, ws = _hier(Levels, None, tmp_ne_nn, substrate, properties, windows=False)
    _swap_l1(Levels, ws, ops.dwell_step(_props(properties), h, float(_host(dt))))
    return Levels


def subcycleGOMELT(Levels, ne_nn, Shapes, substrate, LInterp, tmp_ne_nn, laser_position, properties, laserP,
                   subcycle, max_accum_L3, accum_L3):
    """cF:3224-3632: Level 1 once, Level 2 x N2, Level 3 x N2*N3, predictor pass then corrector pass
    (gomelt_subcycle_f32).  ``max_accum_L3`` / ``accum_L3`` (the melt-time windows, gm:448-449) are updated in place when
    they are CUDA tensors and returned."""
    rows = np.array(_host(laser_position), F32, copy=True).reshape(-1, 7)
    rows[:, 6] = np.asarray(_host(laserP), F32)
    N2, N3 = int(subcycle[0]), int(subcycle[1])
    d = distOf(Levels)
    if d is not None:
        d.begin_call(Levels, tmp_ne_nn, substrate)
        if not d.is_owner:  # the Level-1 side of cF:3306 / 3456: dt_all summed in the order of the native call
            dt_all = F32(0)
            for r in range(N2 * N3):
                dt_all = F32(dt_all + rows[r, 5])
            flags = _lib.STEP_BC_CONST | _lib.STEP_FUSED_FLUX | _lib.STEP_CLAMP | 0x20000
            d.solve(float(dt_all), flags)
            d.solve(float(dt_all), flags)
            d.finish(clamp_after=False)
            return Levels, None, None, None, max_accum_L3, accum_L3
    _ensure_fields(Levels)
    mx, ac = _f(max_accum_L3), _f(accum_L3)
    h, keep, ws = _hier(Levels, Shapes, tmp_ne_nn, substrate, properties, N2, N3)
    _swap_l1(Levels, ws, _native(d, ops.subcycle, _props(properties), h, rows, N2, N3, mx, ac))
    if d is not None:
        d.finish(clamp_after=False, final_mirror=Levels[1]["T0"])
    return Levels, None, None, None, mx, ac


# ----------------------------------------------------------------------------------------------
# window shift
# ----------------------------------------------------------------------------------------------
def _trunc_shift(v, h):
    """(v / h + 1e-2).astype(int): truncation toward zero of a float32 quotient (cF:1716-1718)."""
    return int(np.asarray(F32(v) / F32(h) + F32(1e-2)).astype(int))


def _constrain(vtot, L):
    b = L["bounds"]
    return [np.clip(F32(vtot[i]), F32(b[k][0]), F32(b[k][1])) for i, k in enumerate(("ix", "iy", "iz"))]


def moveEverything(v, vstart, Levels, move_v, LInterp, L1L2Eratio, L2L3Eratio, height):
    """cF:2400-2510: integer-cell shift of the Level-3 / Level-2 windows; T0 / T'0 re-interpolated at the new
    window nodes (one launch per window: gomelt_shift_window_f32); overlap index sets updated; S1 / S2 regathered from
    Level 0.  ``Shapes`` = per-pair fine-element -> parent-cell grouping (three small int arrays each).
    With a slab-decomposed Level 1 (dist.py) the geometry below is evaluated on every rank, the device part on the
    laser owner only, after the Level-1 box under the old and the new window positions has come up from the slabs."""
    torch = _torch()
    d = distOf(Levels)
    worker = d is not None and not d.is_owner
    if not worker:
        _ensure_fields(Levels)
    L0, L1, L2, L3 = Levels
    vtot = np.asarray(_host(v), F32) - np.asarray(_host(vstart), F32)
    old2, old3 = L2["node_coords"], L3["node_coords"]
    # ---- Level 3 (shifts in Level-2 cells) ----
    v3 = _constrain(vtot, L3)
    h2 = L2["h"]
    s3 = [_trunc_shift(v3[i], h2[i]) for i in range(3)]
    new3 = [(np.asarray(L3["init_node_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32) for i in range(3)]
    ov3_nodes = [np.asarray(L3["orig_overlap_nodes"][i]) + s3[i] for i in range(3)]
    ov3_coords = [(np.asarray(L3["orig_overlap_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32) for i in range(3)]
    # ---- Level 2 (shifts in Level-1 cells in x, y; whole layers in z) ----
    v2 = _constrain(vtot, L2)
    h1 = [L1["h"][0], L1["h"][1], F32(height)]
    s2 = [_trunc_shift(v2[i], h1[i]) for i in range(3)]
    new2 = [(np.asarray(L2["init_node_coors"][i], F32) + F32(h1[i]) * s2[i]).astype(F32) for i in range(3)]
    move_v = [s2[i] * int(L1L2Eratio[i]) for i in range(3)]
    sz1 = _trunc_shift(v2[2], L1["h"][2])
    o, c = L2["orig_overlap_nodes"], L2["orig_overlap_coors"]
    ov2_nodes = [np.asarray(o[0]) + s2[0], np.asarray(o[1]) + s2[1], np.asarray(o[2]) + sz1]
    ov2_coords = [(np.asarray(c[0], F32) + F32(L1["h"][0]) * s2[0]).astype(F32),
                  (np.asarray(c[1], F32) + F32(L1["h"][1]) * s2[1]).astype(F32),
                  (np.asarray(c[2], F32) + F32(height) * _trunc_shift(v2[2], height)).astype(F32)]
    if d is not None:
        # the Level-1 temperature under the new window positions (and the old Level-2 one): slabs -> the owner's mirror
        from .dist import Box, footprint_box

        boxes = [footprint_box({"node_coords": cc}, L1) for cc in (new2, new3, old2)]
        lo = [min(b.lo[k] for b in boxes) for k in range(3)]
        hi = [max(b.lo[k] + b.n[k] for b in boxes) for k in range(3)]
        d.up(Box(lo, [hi[k] - lo[k] for k in range(3)]), d.slab.T, None if worker else L1["T0"])
    if not worker:
        ws = _workspace(Levels)
        c1, c2_old, c3_old = _coords(L1["node_coords"]), _coords(old2), _coords(old3)
        # T'3 <- I_3(T'3), T3 <- I_1(T1) + (I_2(T'2) + T'3) at the new nodes (cF:2439-2443), with the OLD Level-2 window
        Tp3n, T3n = ops.shift_window(c1, L1["T0"], c3_old, L3["Tprime0"], _coords(new3), ws.buffer("Tp3", L3["Tprime0"]),
                                     ws.buffer("T3", L3["T0"]), mid_coords=c2_old, Tp_mid=L2["Tprime0"])
        Tp2n, T2n = ops.shift_window(c1, L1["T0"], c2_old, L2["Tprime0"], _coords(new2), ws.buffer("Tp2", L2["Tprime0"]),
                                     ws.buffer("T2", L2["T0"]))
        # the new fields become the Levels' fields, the old tensors become the spare buffers of the next move
        ws.alt["Tp3"], L3["Tprime0"] = L3["Tprime0"], Tp3n
        ws.alt["T3"], L3["T0"] = L3["T0"], T3n
        ws.alt["Tp2"], L2["Tprime0"] = L2["Tprime0"], Tp2n
        ws.alt["T2"], L2["T0"] = L2["T0"], T2n
    L3["overlapCoords"], L2["overlapNodes"], L2["overlapCoords"] = ov3_coords, ov2_nodes, ov2_coords
    L3["node_coords"], L2["node_coords"] = new3, new2
    LInterp = [interpolatePointsMatrix(L1, new2), None]
    # Level-3 overlap indices are relative to Level 2, which has itself moved
    L3["overlapNodes"] = [ov3_nodes[i] - move_v[i] for i in range(3)]
    # ---- Level 0 index sets of the two windows ----
    r3 = [int(x) for x in L2L3Eratio]
    L0["overlapNodes"] = [np.asarray(L0["orig_overlap_nodes"][i]) + r3[i] * s3[i] for i in range(3)]
    L0["overlapCoords"] = [(np.asarray(L0["orig_overlap_coors"][i], F32) + F32(h2[i]) * s3[i]).astype(F32)
                           for i in range(3)]
    hz = [L1["h"][0], L1["h"][1], L2["h"][2]]
    rz = [int(L1L2Eratio[0]) * r3[0], int(L1L2Eratio[1]) * r3[1], r3[2]]
    s0 = [_trunc_shift(v2[i], hz[i]) for i in range(3)]
    L0["overlapNodes_L2"] = [np.asarray(L0["orig_overlap_nodes_L2"][i]) + rz[i] * s0[i] for i in range(3)]
    L0["overlapCoords_L2"] = [(np.asarray(L0["orig_overlap_coors_L2"][i], F32) + F32(hz[i]) * s0[i]).astype(F32)
                              for i in range(3)]
    L0["overlapNodes"][2] = L0["overlapNodes"][2] - move_v[2] * r3[2]
    L0["overlapNodes_L2"][2] = L0["overlapNodes_L2"][2] - move_v[2] * r3[2]
    L0["idx"] = BoxIndex(L0["overlapNodes"], L0["nodes"][0], L0["nodes"][1])
    L0["idx_L2"] = BoxIndex(L0["overlapNodes_L2"], L0["nodes"][0], L0["nodes"][1])
    LInterp[1] = interpolatePointsMatrix(L2, new3)
    if worker:
        return Levels, None, LInterp, move_v
    # ---- state regather from Level 0 (cF:2500-2502), in place: every node of the windows is overwritten ----
    i3, i2 = L0["idx"].device(), L0["idx_L2"].device()
    ops.box_copy(L0["S1"], L2["S1"], i2, L0["nodes"][0], L0["nodes"][1], scatter=False)
    ops.box_copy(L0["S1"], L3["S1"], i3, L0["nodes"][0], L0["nodes"][1], scatter=False)
    ops.box_copy(L0["S2"], L3["S2"], i3, L0["nodes"][0], L0["nodes"][1], scatter=False)
    Shapes = {"L2L1": _pair_cells(L2, L1), "L3L1": _pair_cells(L3, L1), "L3L2": _pair_cells(L3, L2)}
    return Levels, Shapes, LInterp, move_v


# ----------------------------------------------------------------------------------------------
# runs of identical dwell rows as CUDA-graph replays
# ----------------------------------------------------------------------------------------------
_DWELL_TRACE = bool(os.environ.get("GOMELT_DWELL_TRACE"))   # print the capture decisions of dwellRows (read once, at import)


class _DwellGraph:
    def __init__(self):
        self.key = self.graph = self.p0 = None
        self.keep = []
        self.launches = 0
        self.replays = 0


def _pointer_state(Levels, ws):
    """Addresses of every device buffer a dwell row reads or writes, in their current roles."""
    t = [Levels[1]["T0"], Levels[1]["S1"], ws.l1_spare, Levels[0]["S1"], Levels[0]["S2"]]
    for i in (2, 3):
        t += [Levels[i]["T0"], Levels[i]["Tprime0"], Levels[i]["S1"]]
    t += [Levels[3]["S2"]] + [ws.alt.get(k) for k in ("Tp3", "T3", "Tp2", "T2")]
    return tuple(0 if x is None else x.data_ptr() for x in t)


def dwellRows(Levels, n, v, vstart, move_v, LInterp, L1L2Eratio, L2L3Eratio, height, tmp_ne_nn, ne_nn, properties, dt,
              substrate, graphs=True):
    """``n`` consecutive IDENTICAL Level-1-only rows of the driver (gm:292-370 in dwell mode): n x [moveEverything
    cF:2400-2510 at the same laser position + stepGOMELTDwellTime cF:2617-2664 with the same dt].  Such a row is ~10
    small launches whose host issue costs several times their run time, and the windows do not move, so after two rows
    every buffer is back in its role (the window fields and the Level-1 field ping-pong): the first two rows run
    eagerly, the next two are captured as ONE CUDA graph, and the rest of the run - also in later calls with the same
    arguments - replays it.  Returns what the last moveEverything returned: (Levels, Shapes, LInterp, move_v)."""
    torch = _torch()
    Shapes = None

    def one():
        nonlocal Levels, Shapes, LInterp, move_v
        Levels, Shapes, LInterp, move_v = moveEverything(v, vstart, Levels, move_v, LInterp, L1L2Eratio, L2L3Eratio, height)
        Levels = stepGOMELTDwellTime(Levels, tmp_ne_nn, ne_nn, properties, dt, substrate)

    ws = _workspace(Levels)
    done = 0
    if not graphs or distOf(Levels) is not None or n < 3:
        for _ in range(n):
            one()
        return Levels, Shapes, LInterp, move_v
    key = (np.asarray(_host(v), F32).tobytes(), np.asarray(_host(vstart), F32).tobytes(), float(F32(_host(dt))),
           tuple(int(q) for q in tmp_ne_nn), tuple(int(q) for q in substrate), float(F32(height)))
    g = getattr(ws, "dwell_graph", None)
    if g is None or g.key != key:
        g = ws.dwell_graph = _DwellGraph()
        g.key = key
        before = _pointer_state(Levels, ws)
        one()
        one()
        done = 2
        if _pointer_state(Levels, ws) != before:   # first rows after a window move: the roles settle one pair later
            before = _pointer_state(Levels, ws)
            if n - done >= 2:
                one()
                one()
                done += 2
        if _DWELL_TRACE:
            now = _pointer_state(Levels, ws)
            print("dwellRows: n=%d done=%d roles_back=%s differing_slots=%s" % (
                n, done, now == before, [i for i, (a, b) in enumerate(zip(now, before)) if a != b]), file=sys.stderr)
        if n - done >= 2 and _pointer_state(Levels, ws) == before:
            l0 = ops.LAUNCHES
            keep_from = len(_CACHE.store)
            # (captured by hand on a side stream: torch.cuda.graph() would also collect garbage and empty the allocator's
            # cache - tens of milliseconds, and every later allocation back to cudaMalloc; nothing is allocated in here)
            graph = torch.cuda.CUDAGraph()
            side = getattr(ws, "capture_stream", None)
            if side is None:
                side = ws.capture_stream = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph.capture_begin()
                try:
                    one()
                    one()
                finally:
                    graph.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            if _pointer_state(Levels, ws) == before and len(_CACHE.store) == keep_from:
                g.graph, g.p0, g.launches = graph, before, ops.LAUNCHES - l0
                g.keep = [list(_CACHE.store.values()), Shapes, LInterp]   # everything the graph's kernels point at
                graph.replay()   # (capturing does not execute; the library counted these launches at capture)
                g.replays += 1
                done += 2
            else:   # a cache filled during the capture: nothing ran, run the pair now
                if _DWELL_TRACE:
                    print("dwellRows: capture dropped: roles_back=%s cache %d -> %d" % (
                        _pointer_state(Levels, ws) == before, keep_from, len(_CACHE.store)), file=sys.stderr)
                one()
                one()
                done += 2
    if g.graph is not None:
        if n - done >= 1 and _pointer_state(Levels, ws) != g.p0:
            one()   # an odd row count left the buffers in their other roles
            done += 1
        while n - done >= 2 and _pointer_state(Levels, ws) == g.p0:
            g.graph.replay()
            ops.GRAPH_LAUNCHES += g.launches   # kernels executed by replays (not seen by the library's own counter)
            g.replays += 1
            done += 2
    for _ in range(n - done):
        one()
    if Shapes is None:   # only replays ran in this call: the descriptors are those of the captured rows
        Shapes, LInterp = g.keep[1], g.keep[2]
    return Levels, Shapes, LInterp, move_v


# ----------------------------------------------------------------------------------------------
# melt-time bookkeeping and monitors
# ----------------------------------------------------------------------------------------------
def accumSingleStepFused(Levels, all_reset, accum_time, max_accum_time, dt, T_liquidus):
    """gm:339-357 + melting_temp cF:3696-3712 on the Level-0 melt-time arrays, in place, one launch."""
    torch = _torch()
    L0 = Levels[0]
    idx = L0["idx"] if isinstance(L0["idx"], BoxIndex) else BoxIndex(L0["overlapNodes"], L0["nodes"][0], L0["nodes"][1])
    ops.accum_single_step(_f(Levels[3]["T0"]), _dev(all_reset, torch.bool), float(F32(_host(dt))), float(F32(T_liquidus)),
                          accum_time, max_accum_time, idx.device(), idx.nx, idx.ny)
    return accum_time, max_accum_time


def melting_temp(temps, delt_T, T_melt, accum_time, idx):
    """cF:3696-3712: accum_time[idx] += (temps > T_melt) * dt."""
    torch = _torch()
    acc = _f(accum_time).clone()
    idx_t = _dev(np.asarray(idx if isinstance(idx, BoxIndex) else _host(idx)).astype(np.int64)) \
        if not isinstance(idx, torch.Tensor) else idx.long()
    above = (_f(temps) > float(F32(T_melt))).to(torch.float32) * float(F32(_host(delt_T)))
    acc.index_add_(0, idx_t, above)
    return acc


def printLevelMaxMin(Ls, Lnames):
    """cF:3635-3665: print min / max of every level's T0 and stop on invalid physics (non-finite, <= 0 or
    > 1e5 K) like the reference (``sys.exit(1)``); one fused reduction per level instead of two host syncs."""
    import sys

    from . import output

    for name, (lo, hi, bad) in zip(Lnames, output.level_minmax([None] + list(Ls))):
        print(f"{name}: min T = {lo:.2f} K, max T = {hi:.2f} K")
        if bad or not (0 < lo <= 1e5) or not (0 < hi <= 1e5):
            print("Terminating program: temperature out of range")
            sys.exit(1)


def save_object(obj, filename):
    """gm:398 / cF save_object: ``obj`` = [Levels, accum_time, max_accum_time, time_inc, record_inc]; written as a
    raw-dump checkpoint directory next to where the reference writes its ``.pkl`` (output.save_checkpoint)."""
    from . import output

    path = str(filename)
    return output.save_checkpoint(path[:-4] if path.endswith(".pkl") else path, *obj)


def levelMaxMin(Ls):
    """printLevelMaxMin cF:3635-3665 without the prints / exit: [(min, max, ok)] for levels 1.."""
    res = []
    for i in range(1, len(Ls)):
        T = _f(Ls[i]["T0"])
        lo, hi = float(T.min()), float(T.max())
        res.append((lo, hi, all(math.isfinite(x) and 0 < x <= 1e5 for x in (lo, hi))))
    return res
