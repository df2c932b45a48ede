"""
Per-rank wall-clock phases of the end-to-end build (new simulation -> sim.G), to see where the
multi-rank e2e number of bench.py goes.  Launch with torchrun (any N) or plain python (N=1).
"""
import gc
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench  # noqa: E402
from simpeg_b200.sharding import Collectives  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
coll = Collectives()
wl = bench.make_workload("cfg4-shard", world, None)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rows = []
for it in range(4):
    barrier()
    t = [time.perf_counter()]
    survey = bench.make_survey(wl, wl["locations"])
    sim = bench.make_simulation(wl, survey, comm=coll)
    t.append(time.perf_counter())            # 0: constructors
    plan = sim._tile_plan()
    t.append(time.perf_counter())            # 1: tile plan (host)
    sim._tile_plan = lambda: plan
    ops = sim._ops
    torch.cuda.synchronize()
    t.append(time.perf_counter())            # 2: engine (uploads)
    sim.G
    t.append(time.perf_counter())            # 3: launch (alloc + enqueue)
    torch.cuda.synchronize()
    t.append(time.perf_counter())            # 4: kernel
    barrier()
    t.append(time.perf_counter())            # 5: barrier
    rows.append(np.diff(t) * 1e3)
    del sim, ops, survey, plan
    gc.collect()

names = ["ctor", "plan", "engine", "launch", "kernel", "barrier"]
msg = f"rank {rank} cpus={os.cpu_count()} affinity={len(os.sched_getaffinity(0))} " + " | ".join(
    f"it{i}: " + " ".join(f"{n}={v:.1f}" for n, v in zip(names, r)) for i, r in enumerate(rows))
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, msg)
    if rank == 0:
        print("\n".join(out))
    dist.destroy_process_group()
else:
    print(msg)
