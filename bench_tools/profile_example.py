"""cProfile of the example.json run through the drop-in driver (host-side overhead per kernel launch)."""
import cProfile, importlib, io, os, pstats, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "bench_tools"))
import torch
import gomelt_b200 as gm
from run_example import load_input
drv = importlib.import_module("gomelt_b200.driver")
drv.go_melt(load_input(tempfile.mkdtemp()), write_final=False)   # warm
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
res = drv.go_melt(load_input(tempfile.mkdtemp()), write_final=False)
torch.cuda.synchronize()
pr.disable()
print("wall", res["wall_seconds"], "launches", gm.ops.LAUNCHES)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
