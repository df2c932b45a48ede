#!/usr/bin/env python
"""Where the host-vector Jvec / Jtvec of a sharded run spend their time (development tool; torchrun).
Stages are separated with device syncs, so their sum is larger than the pipelined call."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist

    import bench
    from simpeg_b200.sharding import Collectives

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    coll = Collectives()
    wl = bench.make_workload("cfg4-shard", world, 2500)
    wl["g_dtype"] = "f64"
    sim = bench.make_simulation(wl, bench.make_survey(wl, wl["locations"]), comm=coll)
    sim.G
    ops, eng = sim._ops, sim._dev
    rng = np.random.default_rng(0)
    nP, nD = eng.nP, sim.survey.nD
    m, v, w = rng.uniform(-0.05, 0.05, nP), rng.uniform(0, 1, nP), rng.uniform(0, 1, nD)

    def sync(barrier=False):
        if barrier and world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stages = {}

    def timed(name, fn):
        t = time.perf_counter()
        r = fn()
        sync()
        stages.setdefault(name, []).append(1e3 * (time.perf_counter() - t))
        return r

    for it in range(6):
        sync(True)
        t0 = time.perf_counter()
        sim.Jvec(m, v)
        sync(True)
        t1 = time.perf_counter()
        sim.Jtvec(m, w)
        sync(True)
        t2 = time.perf_counter()
        stages.setdefault("Jvec_total", []).append(1e3 * (t1 - t0))
        stages.setdefault("Jtvec_total", []).append(1e3 * (t2 - t1))
        # the same calls, stage by stage
        sync(True)
        x = timed("jvec_h2d_cols", lambda: ops.cols_to_device(v))
        y = timed("jvec_kernel", lambda: eng.gemv_dev(x))
        full = timed("jvec_allgather", lambda: coll.all_gather_rows(y, sim._sharding) if world > 1 else y)
        timed("jvec_d2h_rows", lambda: eng.to_host(full))
        sync(True)
        yl = timed("jtvec_h2d_rows", lambda: ops.local_rows_of(w))
        z = timed("jtvec_kernel_allreduce", lambda: ops.gemv_t_allreduce_dev(yl))
        timed("jtvec_d2h_cols", lambda: eng.to_host(z))
    out = {k: round(float(np.median(vals[1:])), 3) for k, vals in stages.items()}
    res = torch.tensor([out[k] for k in sorted(out)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "max_over_ranks_ms": dict(zip(sorted(out), [round(float(x), 3) for x in res.cpu()]))}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
