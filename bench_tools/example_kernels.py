"""Kernel-time table of one examples/example.json run through the drop-in driver (CUPTI trace via torch.profiler)."""
import collections
import importlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import torch  # noqa: E402

import gomelt_b200 as gm  # noqa: E402
from run_example import load_input  # noqa: E402

drv = importlib.import_module("gomelt_b200.driver")
drv.go_melt(load_input(tempfile.mkdtemp()), write_final=False)  # warm
torch.cuda.synchronize()
t0 = time.time()
res = drv.go_melt(load_input(tempfile.mkdtemp()), write_final=False)
torch.cuda.synchronize()
wall = time.time() - t0
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    drv.go_melt(load_input(tempfile.mkdtemp()), write_final=False)
    torch.cuda.synchronize()
rows = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if "cuda" not in str(getattr(ev, "device_type", "")).lower():
        continue
    us = float(getattr(ev, "device_time", 0.0) or 0.0)
    name = ev.name.split("<")[0].split("(")[0].replace("gomelt::", "").replace("void ", "")
    rows[name][0] += 1
    rows[name][1] += us
tot = sum(v[1] for v in rows.values())
print(json.dumps({"wall_s": wall, "sum_kernel_ms": tot / 1e3,
                  "kernels": [{"kernel": k, "launches": v[0], "ms": round(v[1] / 1e3, 2), "avg_us": round(v[1] / v[0], 1)}
                              for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:18]]}, indent=1))
