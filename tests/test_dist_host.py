"""Host logic of the distributed drop-in (gomelt_b200/dist.py) that needs no GPU: the Level-1 boxes under a window and
their cuts by the slabs' plane ranges (every box transfer is built from these)."""
import numpy as np

from gomelt_b200 import dist
from gomelt_b200.slab import local_extent, partition_active_planes, partition_planes


def _axis(lo, hi, n):
    return np.linspace(lo, hi, n, dtype=np.float32)


def test_box_cut_by_plane_ranges_partitions_the_box():
    box = dist.Box([3, 1, 5], [7, 4, 9])          # planes 5 .. 13
    parts = partition_planes(31, 4)               # (0,8) (8,16) (16,24) (24,31)
    cuts = [box.zcut(a, b) for a, b in parts]
    assert cuts[2] is None and cuts[3] is None
    assert (cuts[0].lo, cuts[0].n) == ([3, 1, 5], [7, 4, 3]) and (cuts[1].lo, cuts[1].n) == ([3, 1, 8], [7, 4, 6])
    assert sum(c.size for c in cuts if c is not None) == box.size      # owned planes: a partition of the box
    # stored planes (owned + ghosts) overlap by the ghost planes: every rank that needs a plane gets it
    stored = [local_extent(q, 4, *parts[q]) for q in range(4)]
    got = [box.zcut(g0, g0 + nzl) for g0, nzl, _, _ in stored]
    assert got[0].n[2] == 4 and got[1].lo[2] == 7 and got[1].n[2] == 7 and got[2] is None
    assert dist.Box([0, 0, 0], [0, 3, 3]).zcut(0, 10) is None and box.zcut(14, 20) is None


def test_footprint_box_covers_every_parent_cell_under_the_window():
    L1 = {"node_coords": [_axis(0, 10, 51), _axis(0, 4, 21), _axis(-4, 2, 31)]}     # h = 0.2
    win = {"node_coords": [_axis(1.4, 5.4, 101), _axis(0.4, 3.6, 81), _axis(-0.36, 0.04, 11)]}
    b = dist.footprint_box(win, L1, margin=2)
    for d, (lo, hi) in enumerate(((1.4, 5.4), (0.4, 3.6), (-0.36, 0.04))):
        xc = L1["node_coords"][d]
        assert xc[b.lo[d]] <= lo - 0.2 or b.lo[d] == 0
        assert xc[b.lo[d] + b.n[d] - 1] >= hi + 0.2 or b.lo[d] + b.n[d] == xc.size
    # a window pushed against the grid's edge: the box is clipped, never empty
    edge = {"node_coords": [_axis(6.0, 10.0, 101), _axis(0.0, 4.0, 101), _axis(1.6, 2.0, 11)]}
    e = dist.footprint_box(edge, L1)
    assert e.lo[0] + e.n[0] == 51 and e.lo[1] == 0 and e.n[1] == 21 and e.lo[2] + e.n[2] == 31 and e.size > 0


def test_index_box_requires_consecutive_nodes():
    b = dist.index_box([np.arange(7, 28), np.arange(2, 23), np.arange(18, 21)])
    assert (b.lo, b.n) == ([7, 2, 18], [21, 21, 3])
    try:
        dist.index_box([np.array([1, 2, 4]), np.arange(3), np.arange(3)])
    except Exception as exc:
        assert "consecutive" in str(exc)
    else:
        raise AssertionError("a strided index set is not a box")


def test_every_plane_has_one_owner_and_ghosts_mirror_the_neighbours():
    for nz, act, world in ((156, 141, 8), (31, 22, 3), (7, 5, 2)):
        parts = partition_active_planes(nz, act, world)
        owner = np.full(nz, -1)
        for q, (a, b) in enumerate(parts):
            assert (owner[a:b] == -1).all()
            owner[a:b] = q
        assert (owner >= 0).all()
        for q in range(world):
            g0, nzl, zb, ze = local_extent(q, world, *parts[q])
            assert g0 + zb == parts[q][0] and g0 + ze == parts[q][1] and nzl == ze + (1 if q < world - 1 else 0)
