#!/usr/bin/env python
"""Multi-rank A/B of the sharded Jtvec (development tool; launch with torchrun on N GPUs):
device time of gemv_t + all-reduce for several slab counts, and the wall clock of a device CG
iteration, cfg4 shard (2500 receivers per GPU)."""
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
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG
    from simpeg_b200.potential_fields import base
    from simpeg_b200.sharding import Collectives

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    coll = Collectives()
    wl = bench.make_workload("cfg4-shard", world, int(os.environ.get("PF_RX", "2500")))
    wl["g_dtype"] = "f64"
    sim = bench.make_simulation(wl, bench.make_survey(wl, wl["locations"]), comm=coll)
    sim.G
    ops, eng = sim._ops, sim._dev
    rng = np.random.default_rng(0)
    nP, nD = eng.nP, sim.survey.nD
    w_local = sim._local_rows_of(rng.uniform(0, 1, nD))
    out = {}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for slabs in (1, 2, 4, 6, 8, 12):
        base.JTVEC_SLABS = slabs
        for _ in range(3):
            ops.gemv_t_allreduce_dev(w_local)
        ts = []
        for _ in range(10):
            sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.gemv_t_allreduce_dev(w_local)
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t))
        out[f"jtvec_ms_slabs{slabs}"] = (round(min(ts), 4), round(float(np.median(ts)), 4))
    for slabs in (1, 6):
        base.JTVEC_SLABS = slabs
        solver = DeviceGaussNewtonCG(sim, data_weights=np.ones(nD), reg_hessian=np.ones(nP), beta=1e-3,
                                     cg_maxiter=5, cg_rtol=0.0)
        m = rng.uniform(-0.05, 0.05, nP)
        solver.set_model(m)
        g = rng.uniform(-1, 1, nP)
        solver.find_search_direction(g)
        ts = []
        for _ in range(6):
            sync()
            t0 = time.perf_counter()
            solver.find_search_direction(g)
            sync()
            ts.append((time.perf_counter() - t0) / 5 * 1e3)
        out[f"cg_ms_per_iter_slabs{slabs}"] = (round(min(ts), 3), round(float(np.median(ts)), 3))
        del solver
    if rank == 0:
        print(json.dumps({"world": world, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
