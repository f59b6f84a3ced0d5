"""Oracle helpers for tests (test infrastructure): build a bare level dict without the JSON schema."""
import numpy as np

from . import config
from .fem import createMesh3D
from .setup import getBCindices


def make_level(elements, bounds):
    """Level dict with the fields the per-level oracle functions read (cF:139-149)."""
    FDT = config.FDT
    ex, ey, ez = [int(e) for e in elements]
    nodes = [ex + 1, ey + 1, ez + 1]
    (x0, x1), (y0, y1), (z0, z1) = bounds
    node_coords, connect = createMesh3D((x0, x1, nodes[0]), (y0, y1, nodes[1]), (z0, z1, nodes[2]))
    nn = nodes[0] * nodes[1] * nodes[2]
    return {
        "elements": [ex, ey, ez],
        "nodes": nodes,
        "ne": ex * ey * ez,
        "nn": nn,
        "h": [FDT((x1 - x0) / ex), FDT((y1 - y0) / ey), FDT((z1 - z0) / ez)],
        "node_coords": node_coords,
        "connect": connect,
        "BC": getBCindices(nodes, nn),
        "bounds": {"x": [x0, x1], "y": [y0, y1], "z": [z0, z1]},
    }


def smooth_field(level, rng, lo=300.0, hi=2600.0, noise=50.0):
    """Seeded smooth + noisy temperature-like field on a level (x-fastest flat order)."""
    FDT = config.FDT
    x, y, z = level["node_coords"]
    X = x[None, None, :]
    Y = y[None, :, None]
    Z = z[:, None, None]
    cx, cy = 0.5 * (x[0] + x[-1]), 0.5 * (y[0] + y[-1])
    r2 = ((X - cx) / (x[-1] - x[0])) ** 2 + ((Y - cy) / (y[-1] - y[0])) ** 2
    depth = (z[-1] - Z) / (z[-1] - z[0] + 1e-30)
    T = lo + (hi - lo) * np.exp(-12.0 * r2 - 3.0 * depth)
    T = T + noise * rng.standard_normal(T.shape)
    return np.maximum(T, 250.0).astype(FDT).reshape(-1)
