"""Host-side level setup of the drop-in (JAX-free): ``example.json`` blocks -> the four ``Levels``
dicts with the reference's field names (SetupLevels cF:115-264 and the static-size helpers
cF:447-579).  Geometry only: 1-D node-coordinate arrays, overlap index sets, bounds.  The mutable
fields (T0, Tprime0, S1, S2, S1_storage) are created by ``computeFunctions.SetupLevels`` on the device.

Every level is a uniform tensor-product grid, so the reference's per-axis connectivity tables and
face index lists (createMesh3D cF:32-64, getBCindices cF:520-559) are not built: kernels derive
them from (nx, ny, nz).  Coordinates are float32 and follow jnp.linspace's evaluation
(start*(1-s) + stop*s with s = i/(n-1), end point exact), because cell searches floor((x-x0)/h)
are taken on these values.
"""
import copy

import numpy as np

F32 = np.float32


def axis_coords(lo, hi, n):
    """jnp.linspace(lo, hi, n) in float32 (cF:47)."""
    n = int(n)
    lo, hi = F32(lo), F32(hi)
    if n == 1:
        return np.array([lo], F32)
    s = np.arange(n - 1, dtype=F32) / F32(n - 1)
    body = lo * (F32(1) - s) + hi * s
    return np.concatenate([body, np.array([hi], F32)]).astype(F32)


def _span(bounds):
    return [bounds[a][1] - bounds[a][0] for a in ("x", "y", "z")]


def _parent_nodes_under(child_axis, parent_axis):
    """Indices of the parent's nodes that lie inside the child's extent (cF:858-887)."""
    pmin = parent_axis.min()
    hp = (parent_axis.max() - pmin) / F32(parent_axis.size - 1)
    first = np.round((child_axis.min() - pmin) / hp)
    last = np.round((child_axis.max() - pmin) / hp) + 1
    return np.arange(first, last).astype(int)


def _fine_nodes_on_parent(parent_axis, fine_axis):
    """Indices of the (larger, finer) grid's nodes coincident with ``parent_axis`` nodes (cF:890-925)."""
    fmin = fine_axis.min()
    hf = (fine_axis.max() - fmin) / F32(fine_axis.size - 1)
    hp = (parent_axis.max() - parent_axis.min()) / F32(parent_axis.size - 1)
    first = np.round((parent_axis.min() - fmin) / hf)
    last = np.round((parent_axis.max() - fmin) / hf) + 1
    return np.arange(first, last, int(np.round(hp / hf))).astype(int)


def overlap_ids(index_vectors, nx, ny):
    """Flat ids of a tensor-product index set, x fastest (getOverlapRegion cF:1642-1669)."""
    a, b, c = (np.asarray(v).astype(np.int64) for v in index_vectors)
    return (a[None, None, :] + b[None, :, None] * int(nx) + c[:, None, None] * int(nx) * int(ny)).reshape(-1)


def build_levels(solver_input, properties):
    """[Level0, Level1, Level2, Level3] geometry dicts."""
    lv = [None] + [copy.deepcopy(solver_input.get(f"Level{i}", {})) for i in (1, 2, 3)]
    for L in lv[1:]:
        el = [int(e) for e in L["elements"]]
        span = _span(L["bounds"])
        L["elements"] = el
        L["nodes"] = [e + 1 for e in el]
        L["length"] = [F32(s) for s in span]
        L["h"] = [F32(s / e) for s, e in zip(span, el)]
        L["ne"] = el[0] * el[1] * el[2]
        L["nn"] = L["nodes"][0] * L["nodes"][1] * L["nodes"][2]
        L["node_coords"] = [axis_coords(L["bounds"][a][0], L["bounds"][a][1], n)
                            for a, n in zip(("x", "y", "z"), L["nodes"])]
    L1, L2, L3 = lv[1], lv[2], lv[3]
    layer = properties["layer_height"]
    L1["n_S1_storage"] = int(round((_span(L1["bounds"])[2] / L1["elements"][2]) / layer))
    for L in (L2, L3):  # travel limits of the window relative to Level 1 (find_max_const cF:1886-1913)
        for a in ("x", "y", "z"):
            L["bounds"]["i" + a] = [L1["bounds"][a][0] - L["bounds"][a][0], L1["bounds"][a][1] - L["bounds"][a][1]]
        L["init_node_coors"] = copy.deepcopy(L["node_coords"])
    # Level-1 planes are raised by whole layers until one meets Level-2's top plane (cF:175-182)
    L1["orig_node_coords"] = copy.deepcopy(L1["node_coords"])
    z = L1["orig_node_coords"][2].copy()
    top2 = L2["node_coords"][2][-1]
    for _ in range(1000000):
        if np.isclose(z - top2, 0, atol=1e-4).any():
            break
        z = (z + F32(layer)).astype(F32)
    else:
        raise ValueError("Level-1 and Level-2 z planes never align")
    L1["node_coords"] = [L1["orig_node_coords"][0].copy(), L1["orig_node_coords"][1].copy(), z]

    def overlap(child, parent):
        nodes = [_parent_nodes_under(child["node_coords"][d], parent["node_coords"][d]) for d in range(3)]
        return nodes, [parent["node_coords"][d][nodes[d]].astype(F32) for d in range(3)]

    for child, parent in ((L2, L1), (L3, L2)):
        child["orig_overlap_nodes"], child["orig_overlap_coors"] = overlap(child, parent)
        child["overlapNodes"] = copy.deepcopy(child["orig_overlap_nodes"])
        child["overlapCoords"] = copy.deepcopy(child["orig_overlap_coors"])

    # Level 0: state-only grid at Level-3 resolution over Level-1's footprint x Level-2's depth (cF:206-246)
    s1, s2, h3 = _span(L1["bounds"]), _span(L2["bounds"]), [s / e for s, e in zip(_span(L3["bounds"]), L3["elements"])]
    L0 = {"elements": [round(s1[0] / h3[0]), round(s1[1] / h3[1]), round(s2[2] / h3[2])]}
    L0["nodes"] = [e + 1 for e in L0["elements"]]
    L0["ne"] = L0["elements"][0] * L0["elements"][1] * L0["elements"][2]
    L0["nn"] = L0["nodes"][0] * L0["nodes"][1] * L0["nodes"][2]
    L0["node_coords"] = [axis_coords(L1["bounds"]["x"][0], L1["bounds"]["x"][1], L0["nodes"][0]),
                         axis_coords(L1["bounds"]["y"][0], L1["bounds"]["y"][1], L0["nodes"][1]),
                         axis_coords(L2["bounds"]["z"][0], L2["bounds"]["z"][1], L0["nodes"][2])]
    L0["orig_node_coords"] = copy.deepcopy(L0["node_coords"])
    L0["orig_overlap_nodes"], L0["orig_overlap_coors"] = overlap(L3, L0)
    L0["overlapNodes"] = copy.deepcopy(L0["orig_overlap_nodes"])
    L0["overlapCoords"] = copy.deepcopy(L0["orig_overlap_coors"])
    L0["orig_overlap_nodes_L2"] = [_fine_nodes_on_parent(L2["node_coords"][d], L0["node_coords"][d]) for d in range(3)]
    L0["orig_overlap_coors_L2"] = [L0["node_coords"][d][L0["orig_overlap_nodes_L2"][d]].astype(F32) for d in range(3)]
    L0["overlapNodes_L2"] = copy.deepcopy(L0["orig_overlap_nodes_L2"])
    L0["overlapCoords_L2"] = copy.deepcopy(L0["orig_overlap_coors_L2"])
    L0["idx"] = overlap_ids(L0["overlapNodes"], L0["nodes"][0], L0["nodes"][1])
    L0["idx_L2"] = overlap_ids(L0["overlapNodes_L2"], L0["nodes"][0], L0["nodes"][1])
    L0["layer_idx_delta"] = int(round(layer / h3[2]))
    lv[0] = L0
    return lv


def static_sizes(Levels):
    """(ne2, ne3, nn1, nn2, nn3) - getStaticNodesAndElements cF:447-470."""
    return (int(Levels[2]["ne"]), int(Levels[3]["ne"]), int(Levels[1]["nn"]), int(Levels[2]["nn"]),
            int(Levels[3]["nn"]))


def active_sizes(Levels, laser_pos):
    """(tmp_ne, tmp_nn): Level-1 elements / nodes at or below the laser plane (cF:495-517)."""
    planes = int((Levels[1]["node_coords"][2] <= F32(laser_pos[2]) + F32(1e-5)).sum())
    ex, ey = Levels[1]["elements"][0], Levels[1]["elements"][1]
    return (ex * ey * (planes - 1), (ex + 1) * (ey + 1) * planes)


def substrate_counts(Levels):
    """Per level 0..3: number of nodes in planes with z < 1e-5 (cF:562-579)."""
    return tuple(int((np.asarray(L["node_coords"][2]) < 1e-5).sum()) * L["nodes"][0] * L["nodes"][1]
                 for L in Levels[:4])
