"""examples/example.json through the driver under different states of the host-side caches (coordinate uploads, per-position
pair descriptors): warm (a previous identical run filled them), cold (cleared before the run), nearly full (the state
other workloads leave behind).  GOMELT_DWELL_TRACE=1 prints the graph-capture decisions."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GOMELT_DWELL_TRACE", "1")
import numpy as np  # noqa: E402

import gomelt_b200 as gm  # noqa: E402
from bench_tools.run_example import load_input  # noqa: E402

gm.load()
cf = gm.computeFunctions


def run(tag, graphs=True):
    g0 = gm.ops.GRAPH_LAUNCHES
    res = gm.driver.go_melt(load_input(tempfile.mkdtemp()), write_final=False, graphs=graphs)
    print("%-28s wall %.4f s, graph kernels %d, pair cache %d, coord cache %d" % (
        tag, res["wall_seconds"], gm.ops.GRAPH_LAUNCHES - g0, len(cf._PAIR_CACHE), len(cf._CACHE.store)), file=sys.stderr)


run("first (imports, allocator)", graphs=False)
run("warm eager", graphs=False)
run("warm graphs")
run("warm graphs")
for _ in range(2):
    cf._PAIR_CACHE.clear()
    cf._CACHE.store.clear()
    run("cold caches, graphs")
for _ in range(2):
    cf._PAIR_CACHE.clear()
    cf._CACHE.store.clear()
    run("cold caches, eager", graphs=False)
# nearly full caches
for i in range(505):
    cf._PAIR_CACHE[("dummy", i)] = None
run("pair cache nearly full")
run("again")
for i in range(4000):
    cf._CACHE.get(np.arange(3, dtype=np.float32) + i, np.float32)
run("coord cache nearly full")
run("again")
run("again")
