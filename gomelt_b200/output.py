"""Output path of the drop-in, dependency-free.

* ``write_vtr`` / ``read_vtr``: VTK XML RectilinearGrid files with raw appended binary data - what
  ``pyevtk.hl.gridToVTK`` writes for the reference's ``saveResult`` / ``saveFinalResult`` / ``saveState``
  (cF:1945-2057) - from NumPy alone (pyevtk is not a dependency of this package).
* ``saveResult`` / ``saveFinalResult`` / ``saveState`` / ``saveResults`` / ``saveResultsFinal``: the reference's
  names, arguments, file names (``Level1_00000001.vtr`` ...) and record-step rule (cF:3668-3693).  Device fields
  are copied to pinned host memory on a side stream (``Snapshot``), so a record does not serialise the stepping
  stream; the file is written when the copy has landed.
* ``save_checkpoint`` / ``load_checkpoint``: the reference pickles the whole ``Levels`` pytree with dill
  (gm:390-402, 113-121); here a checkpoint is a directory with one raw little-endian dump per array and a JSON
  header, which scales to 50-200 M-node levels and needs no unpickling of device arrays.  The toolpath seek
  contract of the restart (fixed-width rows, gm:123-126) is ``toolpath_seek``.
* ``level_minmax``: ``printLevelMaxMin`` (cF:3635-3665) without the per-level host round trips: one fused
  reduction launch per level (gomelt_minmax_f32: min, max and the number of non-finite values).

Nothing here is on the timed path; the host code is NumPy, the device code is the C-ABI library.
"""
import json
import os
import struct

import numpy as np

F32 = np.float32
_VTK_TYPES = {"float32": "Float32", "float64": "Float64", "int32": "Int32", "uint8": "UInt8", "int64": "Int64"}
_NP_TYPES = {v: k for k, v in _VTK_TYPES.items()}


# ----------------------------------------------------------------------------------------------
# .vtr (VTK XML RectilinearGrid, raw appended data, UInt64 headers)
# ----------------------------------------------------------------------------------------------
def write_vtr(path, x, y, z, point_data):
    """Write ``path`` (``.vtr`` appended when missing).  ``x, y, z``: 1-D coordinate arrays; ``point_data``:
    name -> array of shape (nx, ny, nz) (the layout pyevtk's gridToVTK takes) or flat x-fastest (nx*ny*nz)."""
    if not path.endswith(".vtr"):
        path += ".vtr"
    x, y, z = (np.ascontiguousarray(np.asarray(c)) for c in (x, y, z))
    nx, ny, nz = x.size, y.size, z.size
    blocks, entries, offset = [], [], 0

    def add(name, arr, ncomp=1):
        nonlocal offset
        a = np.ascontiguousarray(arr)
        if a.dtype.name not in _VTK_TYPES:
            a = a.astype(F32)
        a = a.astype(a.dtype.newbyteorder("<"), copy=False)
        entries.append((name, _VTK_TYPES[a.dtype.name], ncomp, offset))
        blocks.append(a)
        offset += 8 + a.nbytes

    fields = []
    for name, arr in point_data.items():
        a = np.asarray(arr)
        if a.ndim == 3:  # (nx, ny, nz) -> x-fastest flat order (Fortran order of that shape)
            if a.shape != (nx, ny, nz):
                raise ValueError(f"write_vtr: field {name!r} has shape {a.shape}, grid is {(nx, ny, nz)}")
            a = a.ravel(order="F")
        elif a.size != nx * ny * nz:
            raise ValueError(f"write_vtr: field {name!r} has {a.size} values, grid has {nx * ny * nz} nodes")
        fields.append((name, a.ravel()))
    for name, a in fields:
        add(name, a)
    first_coord = len(entries)
    for name, c in (("x_coordinates", x), ("y_coordinates", y), ("z_coordinates", z)):
        add(name, c)
    ext = f"0 {nx - 1} 0 {ny - 1} 0 {nz - 1}"
    head = ['<?xml version="1.0"?>',
            '<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
            f'<RectilinearGrid WholeExtent="{ext}">', f'<Piece Extent="{ext}">', "<PointData>"]
    for name, typ, ncomp, off in entries[:first_coord]:
        head.append(f'<DataArray Name="{name}" NumberOfComponents="{ncomp}" type="{typ}" format="appended" offset="{off}"/>')
    head += ["</PointData>", "<CellData>", "</CellData>", "<Coordinates>"]
    for name, typ, ncomp, off in entries[first_coord:]:
        head.append(f'<DataArray Name="{name}" NumberOfComponents="{ncomp}" type="{typ}" format="appended" offset="{off}"/>')
    head += ["</Coordinates>", "</Piece>", "</RectilinearGrid>", '<AppendedData encoding="raw">']
    with open(path, "wb") as fh:
        fh.write(("\n".join(head) + "\n_").encode())
        for a in blocks:
            fh.write(struct.pack("<Q", a.nbytes))
            fh.write(a.tobytes())
        fh.write(b"\n</AppendedData>\n</VTKFile>\n")
    return path


def read_vtr(path):
    """Read back a file written by ``write_vtr``: ((x, y, z), {name: flat x-fastest array})."""
    import re

    raw = open(path, "rb").read()
    mark = raw.index(b'<AppendedData encoding="raw">')
    data0 = raw.index(b"_", mark) + 1
    head = raw[:mark].decode()
    out = {}
    for m in re.finditer(r'<DataArray Name="([^"]+)" NumberOfComponents="\d+" type="(\w+)" format="appended" offset="(\d+)"/>', head):
        name, typ, off = m.group(1), m.group(2), int(m.group(3))
        (nbytes,) = struct.unpack_from("<Q", raw, data0 + off)
        out[name] = np.frombuffer(raw, dtype=np.dtype(_NP_TYPES[typ]).newbyteorder("<"), count=nbytes // np.dtype(_NP_TYPES[typ]).itemsize,
                                  offset=data0 + off + 8).copy()
    coords = tuple(out.pop(k) for k in ("x_coordinates", "y_coordinates", "z_coordinates"))
    return coords, out


# ----------------------------------------------------------------------------------------------
# device -> pinned host snapshots
# ----------------------------------------------------------------------------------------------
def _is_tensor(a):
    return hasattr(a, "is_cuda")


class Snapshot:
    """Asynchronous device -> pinned-host copies of a few fields (a record must not stall the stepping stream:
    the copy runs on a side stream after the fields' producers; ``arrays()`` waits for it)."""

    def __init__(self, fields):
        self.host, self.event = {}, None
        tensors = {k: v for k, v in fields.items() if _is_tensor(v) and v.is_cuda}
        for k, v in fields.items():
            if k not in tensors:
                self.host[k] = np.asarray(v.detach().cpu().numpy() if _is_tensor(v) else v)
        if tensors:
            import torch

            dev = next(iter(tensors.values())).device
            side = _side_stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            self._pinned = {}
            with torch.cuda.stream(side):
                for k, v in tensors.items():
                    buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    buf.copy_(v, non_blocking=True)
                    v.record_stream(side)
                    self._pinned[k] = buf
                self.event = torch.cuda.Event()
                self.event.record(side)

    def arrays(self):
        if self.event is not None:
            self.event.synchronize()
            for k, buf in self._pinned.items():
                self.host[k] = buf.numpy()
            self.event = None
        return self.host


_SIDE = {}


def _side_stream(dev):
    import torch

    key = str(dev)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


# ----------------------------------------------------------------------------------------------
# the reference's save functions (cF:1945-2057, 3668-3693)
# ----------------------------------------------------------------------------------------------
def _level_grid(Level, zoffset):
    c = [np.asarray(Level["node_coords"][d]) for d in range(3)]
    return c[0], c[1], c[2] - zoffset


def _save(Level, names, path, zoffset):
    snap = Snapshot({k: Level[k] for k in names.values()})
    x, y, z = _level_grid(Level, zoffset)
    h = snap.arrays()
    return write_vtr(path, x, y, z, {label: np.asarray(h[key], dtype=F32).ravel() for label, key in names.items()})


_T_AND_S = {"Temperature (K)": "T0", "State (Powder/Solid)": "S1"}


def saveResult(Level, save_str, record_lab, save_path, zoffset):
    """cF:1945-1983: ``{save_path}{save_str}{record_lab:08}.vtr`` with the temperature and state fields."""
    return _save(Level, _T_AND_S, f"{save_path}{save_str}{record_lab:08}", zoffset)


def saveFinalResult(Level, save_str, save_path, zoffset):
    """cF:1986-2023: ``{save_path}{save_str}Final.vtr``."""
    return _save(Level, _T_AND_S, f"{save_path}{save_str}Final", zoffset)


def saveState(Level, save_str, record_lab, save_path, zoffset):
    """cF:2026-2057: the state field only."""
    return _save(Level, {"State (Powder/Solid)": "S1"}, f"{save_path}{save_str}{record_lab:08}", zoffset)


def saveResults(Levels, Nonmesh, savenum):
    """cF:3668-3682: Levels 2 and 3 every record, Level 1 on its own record step."""
    out = []
    if Nonmesh["output_files"] == 1:
        if savenum == 1 or (np.mod(savenum, Nonmesh["Level1_record_step"]) == 1 or Nonmesh["Level1_record_step"] == 1):
            out.append(saveResult(Levels[1], "Level1_", savenum, Nonmesh["save_path"], 2e-3))
        out.append(saveResult(Levels[2], "Level2_", savenum, Nonmesh["save_path"], 1e-3))
        out.append(saveResult(Levels[3], "Level3_", savenum, Nonmesh["save_path"], 0))
    return out


def saveResultsFinal(Levels, Nonmesh):
    """cF:3685-3693."""
    out = []
    if Nonmesh["output_files"] == 1:
        out.append(saveFinalResult(Levels[1], "Level1_", Nonmesh["save_path"], 2e-3))
        out.append(saveFinalResult(Levels[2], "Level2_", Nonmesh["save_path"], 1e-3))
        out.append(saveFinalResult(Levels[3], "Level3_", Nonmesh["save_path"], 0))
    return out


def driver_hooks():
    """Hooks for ``driver.go_melt(..., hooks=driver_hooks())`` that restore the reference's file output:
    saveResults at every record (gm:467-470), saveState(Level 0) at a layer change (gm:245), checkpoints at
    the end of a layer (gm:390-402), the min / max monitor (gm:473-474)."""
    def on_checkpoint(Levels, accum_time, max_accum_time, time_inc, record_inc, Nonmesh):
        d = os.path.join(Nonmesh["save_path"] + "checkpoint", f"Checkpoint{str(Nonmesh['layer_num']).zfill(4)}")
        save_checkpoint(d, Levels, accum_time, max_accum_time, time_inc, record_inc)

    def on_info(Levels):
        """printLevelMaxMin cF:3635-3665: the monitor stops the run on invalid physics (sys.exit(1) there)."""
        stop = False
        for i, (lo, hi, bad) in enumerate(level_minmax(Levels), start=1):
            print(f"Level {i}: min {lo:.3f} K, max {hi:.3f} K" + (f", {bad} non-finite values" if bad else ""))
            stop = stop or bad or not (0 < lo <= 1e5) or not (0 < hi <= 1e5)
        if stop:
            print("Terminating program: temperature out of range")
            raise SystemExit(1)

    def load_ckpt(Nonmesh, device):
        d = os.path.join(Nonmesh["save_path"] + "checkpoint", f"Checkpoint{str(Nonmesh['layer_num']).zfill(4)}")
        return load_checkpoint(d, device=device)

    def on_layer_accum(accum_time, Nonmesh):
        np.savez(Nonmesh["save_path"] + "accum_time" + str(Nonmesh["layer_num"]).zfill(4), accum_time=_to_host(accum_time))

    def on_final(Levels, Nonmesh):
        saveState(Levels[0], "Level0_", Nonmesh["layer_num"], Nonmesh["save_path"], 0)
        saveResultsFinal(Levels, Nonmesh)

    return {"on_record": saveResults, "on_checkpoint": on_checkpoint, "on_info": on_info, "on_final": on_final,
            "load_checkpoint": load_ckpt, "on_layer_accum": on_layer_accum,
            "on_layer_state": lambda Levels, Nonmesh: saveState(Levels[0], "Level0_", Nonmesh["layer_num"],
                                                                Nonmesh["save_path"], 0)}


# ----------------------------------------------------------------------------------------------
# checkpoints (raw dumps + JSON header)
# ----------------------------------------------------------------------------------------------
def _to_host(a):
    if _is_tensor(a):
        return a.detach().cpu().numpy()
    return a


def _dump_tree(obj, root, prefix, index):
    """Arrays -> files, everything else -> JSON-able structure with {"__array__": file} placeholders."""
    obj = _to_host(obj)
    if hasattr(obj, "vectors") and hasattr(obj, "__array__"):  # computeFunctions.BoxIndex -> the flat ids it stands for
        obj = np.asarray(obj)
    if isinstance(obj, np.ndarray):
        if obj.ndim == 0:
            return {"__scalar__": obj.item(), "dtype": obj.dtype.str}
        name = f"{prefix}.bin"
        a = np.ascontiguousarray(obj)
        a.astype(a.dtype.newbyteorder("<"), copy=False).tofile(os.path.join(root, name))
        index[name] = {"dtype": a.dtype.str, "shape": list(a.shape)}
        return {"__array__": name}
    if isinstance(obj, dict):
        return {"__dict__": {str(k): _dump_tree(v, root, f"{prefix}.{k}", index) for k, v in obj.items()
                             if not str(k).startswith("_gomelt")}}
    if isinstance(obj, (list, tuple)):
        return {"__list__": [_dump_tree(v, root, f"{prefix}.{i}", index) for i, v in enumerate(obj)],
                "tuple": isinstance(obj, tuple)}
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return {"__scalar__": float(obj), "dtype": np.dtype(type(obj)).str}
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    return obj


def _load_tree(node, root, index, to_array):
    if isinstance(node, dict):
        if "__array__" in node:
            meta = index[node["__array__"]]
            return to_array(np.fromfile(os.path.join(root, node["__array__"]), dtype=np.dtype(meta["dtype"])).reshape(meta["shape"]))
        if "__scalar__" in node:
            return np.dtype(node["dtype"]).type(node["__scalar__"])
        if "__dict__" in node:
            return {k: _load_tree(v, root, index, to_array) for k, v in node["__dict__"].items()}
        if "__list__" in node:
            seq = [_load_tree(v, root, index, to_array) for v in node["__list__"]]
            return tuple(seq) if node.get("tuple") else seq
    return node


def save_checkpoint(path, Levels, accum_time, max_accum_time, time_inc, record_inc):
    """gm:390-402 without pickle: ``path/`` holds one raw dump per array and ``header.json``."""
    os.makedirs(path, exist_ok=True)
    index = {}
    tree = _dump_tree({"Levels": Levels, "accum_time": accum_time, "max_accum_time": max_accum_time}, path, "f", index)
    with open(os.path.join(path, "header.json"), "w") as fh:
        json.dump({"format": "gomelt-b200 checkpoint v1", "time_inc": int(time_inc), "record_inc": float(record_inc),
                   "arrays": index, "tree": tree}, fh)
    return path


def load_checkpoint(path, device=None):
    """-> (Levels, accum_time, max_accum_time, time_inc, record_inc) (the tuple the reference unpickles, gm:113-118).
    ``device``: None keeps NumPy arrays; "cuda" uploads the float32 / bool / uint8 node fields."""
    head = json.load(open(os.path.join(path, "header.json")))

    def to_array(a):
        if device is not None and a.ndim == 1 and a.size > 4096 and a.dtype in (np.float32, np.uint8, np.bool_):
            import torch

            return torch.as_tensor(a).to(device)
        return a

    t = _load_tree(head["tree"], path, head["arrays"], to_array)
    return t["Levels"], t["accum_time"], t["max_accum_time"], head["time_inc"], head["record_inc"]


def toolpath_seek(fh, rows_done):
    """gm:123-126: position an open toolpath file after ``rows_done`` rows (all rows have the same width)."""
    fh.seek(0)
    line_len = len(fh.readline())
    fh.seek(rows_done * line_len)
    return line_len


# ----------------------------------------------------------------------------------------------
# min / max monitor
# ----------------------------------------------------------------------------------------------
def level_minmax(Levels):
    """printLevelMaxMin cF:3635-3665 as one fused reduction per level: [(min, max, n_nonfinite)] for levels 1.."""
    from . import ops

    outs = [ops.minmax(L["T0"]) for L in Levels[1:]]
    return [(float(o[0]), float(o[1]), int(o[2])) for o in (t.cpu().numpy() for t in outs)]
