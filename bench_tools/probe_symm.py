"""Probe: torch symmetric memory (peer-mapped buffers + device-side barrier) between the ranks of one box."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", hdl.rank, hdl.world_size, [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
t.fill_(float(rank))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
peer[rank + 1] = 100.0 + rank       # store into the neighbour's buffer
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "got from neighbour", float(t[(rank - 1) % world + 1]), flush=True)
for fn, name in ((lambda: hdl.barrier(channel=0), "barrier"),):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(rank, name, "us per call", 10 * e0.elapsed_time(e1), flush=True)
# neighbour signals
nb = [r for r in (rank - 1, rank + 1) if 0 <= r < world]
def sig():
    for r in nb:
        hdl.put_signal(r, channel=1)
    for r in nb:
        hdl.wait_signal(r, channel=1)
for _ in range(5):
    sig()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(100):
    sig()
e1.record()
host = (time.perf_counter() - t0) * 1e4
torch.cuda.synchronize()
print(rank, "put/wait signal us per sweep", 10 * e0.elapsed_time(e1), "host us", host, flush=True)
dist.destroy_process_group()
