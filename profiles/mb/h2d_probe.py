"""Concurrent host -> device bandwidth of N ranks (torchrun) under three placements of each rank's thread before it allocates
and touches its pinned buffer: no affinity change, vCPU group = rank, vCPU group = reversed rank.  The guest sees one NUMA node
(profiles/r02_topology.txt); if the hypervisor places guest memory by the touching vCPU's physical socket, a placement that
matches the GPU's socket shows up here as bandwidth.  Prints one line per mode: per-rank GB/s with all ranks copying at once."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
ncpu = os.cpu_count()
per = max(1, ncpu // world)
all_cpus = set(range(ncpu))
MB = 64
dev = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")


def run(mode):
    if mode == "none":
        os.sched_setaffinity(0, all_cpus)
    else:
        g = rank if mode == "rank" else world - 1 - rank
        os.sched_setaffinity(0, set(range(g * per, (g + 1) * per)))
    host = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)  # first touch from the pinned thread
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([40 * MB / 1024 / dt], device="cuda")
    out = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(out, gbs)
    if rank == 0:
        v = [round(float(o), 1) for o in out]
        print(f"mode={mode:<8} per-rank GB/s {v}  sum {sum(v):.0f}", flush=True)
    del host


for m in ("none", "rank", "reversed", "none"):
    run(m)
dist.destroy_process_group()
