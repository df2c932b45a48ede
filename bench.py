#!/usr/bin/env python
"""
bench.py -- headline benchmark of the potential-field hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our CUDA arm)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU arm)

Metric (BASELINE.json): G-build (receiver x cell) pairs/s and Jvec+Jtvec GB/s.
Workload (default "cfg4-shard"): BASELINE config 4 -- magnetic vector inversion,
TMI, TensorMesh 100x100x100 (h = 10 m, 1 M active cells, 3 M columns), fp64 G --
with 2500 receivers PER GPU (weak scaling; N = 8 is the full 20 000-receiver,
480 GB configuration).  One step = build the local block of G + Jvec + Jtvec.

One JSON line is printed by rank 0.  `value` is the G-build throughput of the
whole job (pairs of all ranks / max-over-ranks device time of the build part of
the steps); `products` carries the Jvec+Jtvec GB/s; `e2e` is the same measured
through the public Simulation3DIntegral API with host (numpy) buffers.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "G-build (receiver x cell) pairs/s"
UNIT = "pairs/s"
# second-order kernel set = 339 flop-eq per node + 64 per cell
FLOP_EQ_PER_NODE, FLOP_EQ_PER_CELL = 339.0, 64.0


# --------------------------------------------------------------------------- workload
def make_workload(name, n_gpus, rx_per_gpu):
    from simpeg_b200.meshes import TensorMesh

    if name == "cfg4-shard":
        shape, h, z, physics, comps, model_type = (100, 100, 100), 10.0, 10.0, "magnetics", ["tmi"], "vector"
        default_rx = 2500
    elif name == "cfg2-shard":      # gravity gradient tensor + gz, 100x100x50
        shape, h, z, physics = (100, 100, 50), 10.0, 5.0, "gravity"
        comps, model_type = ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar"
        default_rx = 1250
    elif name == "cfg1":            # tutorial size, gz
        shape, h, z, physics, comps, model_type = (40, 40, 20), 5.0, 1.0, "gravity", ["gz"], "scalar"
        default_rx = 441
    else:
        raise SystemExit(f"unknown workload {name!r}")
    rx_per_gpu = default_rx if rx_per_gpu is None else rx_per_gpu
    n_rx = rx_per_gpu * n_gpus
    mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
    side = int(np.ceil(np.sqrt(n_rx)))
    xs = np.linspace(-shape[0] * h / 2, shape[0] * h / 2, side)
    ys = np.linspace(-shape[1] * h / 2, shape[1] * h / 2, side)
    X, Y = np.meshgrid(xs, ys)
    rng = np.random.default_rng(1)
    loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, z)][:n_rx]
    loc[:, 2] += rng.uniform(0.0, 1.0, n_rx)          # seeded jitter
    return dict(name=name, mesh=mesh, shape=shape, h=h, physics=physics, components=comps,
                model_type=model_type, locations=loc, n_rx=n_rx, rx_per_gpu=rx_per_gpu,
                b0=(50000.0, 60.0, 250.0))


def make_survey(wl, locations):
    from simpeg_b200.potential_fields import gravity, magnetics

    if wl["physics"] == "gravity":
        rx = gravity.receivers.Point(locations, components=wl["components"])
        return gravity.survey.Survey(gravity.sources.SourceField([rx]))
    rx = magnetics.receivers.Point(locations, components=wl["components"])
    amp, inc, dec = wl["b0"]
    return magnetics.survey.Survey(magnetics.sources.UniformBackgroundField([rx], amp, inc, dec))


def make_simulation(wl, survey, comm=None):
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity, magnetics

    nC = wl["mesh"].n_cells
    dtype = np.float32 if wl.get("g_dtype") == "f32" else np.float64
    if wl["physics"] == "gravity":
        return gravity.Simulation3DIntegral(wl["mesh"], survey=survey, rhoMap=maps.IdentityMap(nP=nC),
                                            sensitivity_dtype=dtype, comm=comm)
    nP = nC * (3 if wl["model_type"] == "vector" else 1)
    return magnetics.Simulation3DIntegral(wl["mesh"], survey=survey, chiMap=maps.IdentityMap(nP=nP),
                                          model_type=wl["model_type"], sensitivity_dtype=dtype,
                                          comm=comm)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples drawing more than half of the peak power seen
        pmax = max(power)
        loaded = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": pmax}


# --------------------------------------------------------------------------- CPU baseline
def cpu_baseline(wl, sample_rx=None, repeats=1):
    """
    The oracle port (oracle/pf_oracle.c: C + OpenMP restatement of the reference's
    numba-parallel loops) timed on the host cores on a bounded receiver
    subsample x the full mesh; products timed with numpy on the sub-matrix
    (the reference's own expressions, gravity/simulation.py:387, 418).
    """
    from oracle import oracle

    mesh = wl["mesh"]
    cores = oracle.num_threads()
    nC = mesh.n_cells
    if sample_rx is None:
        # ~10-20 s of CPU work: the survey probe measured ~5e6 pairs/s/core for TMI
        nblk = 3 if (wl["physics"] == "magnetics" and wl["model_type"] == "vector") else 1
        by_time = (4e6 * cores * 12.0) // nC
        by_mem = 4e9 // (nC * nblk * 8 * len(wl["components"]))      # <= 4 GB sub-matrix on the host
        sample_rx = int(max(8, min(wl["n_rx"], by_time, by_mem)))
    amp, inc, dec = wl["b0"]
    from simpeg_b200.utils import dip_azimuth2cartesian

    b0 = amp * dip_azimuth2cartesian(inc, dec).squeeze()
    # bounded sample: chunks of `sample_rx` receivers (<= 4 GB of G on the host each)
    # until `budget_s` seconds of CPU work or the whole survey is done
    budget_s = 12.0 if repeats else 0.0
    done_rx, best, G, sim = 0, 0.0, None, None
    start = 0
    while True:
        loc = wl["locations"][start:start + sample_rx]
        if loc.shape[0] == 0:
            break
        sim = oracle.OracleSimulation(wl["physics"], mesh.nodes, mesh.cell_nodes,
                                      [(loc, wl["components"])], b0=b0,
                                      model_type=wl["model_type"], backend="c", parallel=True)
        t = time.perf_counter()
        G = sim.sensitivity_matrix()
        best += time.perf_counter() - t
        done_rx += loc.shape[0]
        start += sample_rx
        if best >= budget_s:
            break
    loc = wl["locations"][:done_rx]
    pairs = loc.shape[0] * nC
    rng = np.random.default_rng(2)
    v, w = rng.uniform(size=sim.nP), rng.uniform(size=sim.nD)
    G @ v
    t = time.perf_counter()
    G @ v
    G.T @ w
    tprod = time.perf_counter() - t
    return {"value": pairs / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done_rx} receivers (chunks of {sample_rx}) x full mesh ({nC} cells, "
                      f"{sim.nP} columns), oracle/pf_oracle.c OpenMP, {best:.2f} s",
            "jvec_jtvec_gbs": 2 * G.nbytes / tprod / 1e9, "seconds": best}


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, args.gpus, args.receivers_per_gpu)
    for _ in range(max(args.warmup, 1)):
        cpu_baseline(wl, sample_rx=8, repeats=0)
    times, last = [], None
    for _ in range(args.steps):
        t = time.perf_counter()
        last = cpu_baseline(wl)
        times.append(time.perf_counter() - t)
    out = {
        "impl": "reference", "metric": METRIC, "value": last["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, args.gpus),
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "products": {"jvec_jtvec_gbs": last["jvec_jtvec_gbs"]},
        "e2e": {"value": last["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def workload_config(wl, n_gpus):
    nC = wl["mesh"].n_cells
    nblk = 3 if (wl["physics"] == "magnetics" and wl["model_type"] == "vector") else 1
    return {
        "workload": f"{wl['name']}: {wl['physics']} {'+'.join(wl['components'])} "
                    f"model_type={wl['model_type']} TensorMesh {'x'.join(map(str, wl['shape']))} "
                    f"h={wl['h']:g} m, {wl['rx_per_gpu']} receivers per GPU "
                    f"({wl['n_rx']} total), fp64 G",
        "cells": nC, "columns": nC * nblk, "receivers": wl["n_rx"],
        "rows": wl["n_rx"] * len(wl["components"]),
        "G_bytes_per_gpu": wl["rx_per_gpu"] * len(wl["components"]) * nC * nblk * 8,
        "step": "G build + Jvec + Jtvec", "parallelism": f"receiver-row shards x{n_gpus}",
        "l2": "inputs larger than L2 (G block per GPU >> 126 MB)",
    }


# --------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist

    from simpeg_b200.engine import launch_count, probe_fp64_peak
    from simpeg_b200.potential_fields.base import SLICED_UPLOAD_MIN
    from simpeg_b200.sharding import Collectives

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    coll = Collectives()
    wl = make_workload(args.workload, world, args.receivers_per_gpu)
    wl["g_dtype"] = args.g_dtype
    g_item = 4 if args.g_dtype == "f32" else 8
    nC = wl["mesh"].n_cells
    rng = np.random.default_rng(5)
    survey = make_survey(wl, wl["locations"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- simulation through the public API (also builds the engine) ------------
    sim = make_simulation(wl, survey, comm=coll)
    nP = sim._n_columns()
    nD = survey.nD
    m_host = rng.uniform(-0.05, 0.05, nP)
    v_host = rng.uniform(0.0, 1.0, nP)
    w_host = rng.uniform(0.0, 1.0, nD)
    eng = sim._dev
    pairs_local = eng.n_rows // len(wl["components"]) * nC
    pairs_total = wl["n_rx"] * nC
    g_bytes_local = eng.n_rows * nP * g_item
    g_bytes_total = nD * nP * g_item

    x_dev = eng.to_device(v_host)
    w_local = sim._local_rows_of(w_host)

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        eng.build_G()
        ev[1].record()
        y = eng.gemv_dev(x_dev)                          # Jvec: local rows ...
        full = coll.all_gather_rows(y, sim._sharding)    # ... gathered on every rank
        ev[2].record()
        z = eng.gemv_t_dev(w_local)                      # Jtvec: partial columns ...
        coll.all_reduce_sum(z)                           # ... summed over ranks
        ev[3].record()
        if timers is not None:
            timers.append(ev)
        return full, z

    for _ in range(args.warmup):
        step()
    fp64_peak = probe_fp64_peak()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = launch_count()
    timers = []
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(args.steps):
        step(timers)
    e_stop.record()
    barrier()
    launches = launch_count() - launches0
    clocks = sampler.stop()
    t_total = max_over_ranks(e_start.elapsed_time(e_stop)) * 1e-3
    t_build = max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in timers)) * 1e-3
    t_jvec = max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in timers)) * 1e-3
    t_jtvec = max_over_ranks(sum(e[2].elapsed_time(e[3]) for e in timers)) * 1e-3
    K = args.steps
    value = pairs_total * K / t_build
    prod_gbs = 2 * g_bytes_total * K / (t_jvec + t_jtvec) / 1e9

    # ---- the other sinks of the path (matrix-free forward / G^T v / diag, weighted column
    #      norms of the stored G), timed on the device on a receiver subsample ------------
    extras = {}
    if rank == 0:
        from simpeg_b200.engine import CudaEngine

        n_sub = min(256, eng.n_rows // len(wl["components"]))
        sub = CudaEngine(wl["physics"], None, None, [(wl["locations"][:n_sub], wl["components"])],
                         b0=np.asarray(survey.source_field.b0) if wl["physics"] == "magnetics" else None,
                         model_type=wl["model_type"], plan=eng.host_plan)
        sub_pairs = n_sub * nC
        ws = sub.to_device(w_host[: sub.n_rows])

        def dev_ms(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            best = np.inf
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            return best

        extras = {
            "sample": f"{n_sub} receivers x full mesh, device-resident vectors, best of 3",
            "forward_only_dpred_pairs_s": sub_pairs / dev_ms(lambda: sub.forward_dev(x_dev)) * 1e3,
            "forward_only_gt_dot_v_pairs_s": sub_pairs / dev_ms(lambda: sub.gt_dot_v_dev(ws)) * 1e3,
            "forward_only_gtg_diag_pairs_s": sub_pairs / dev_ms(lambda: sub.gtg_diag_dev(ws)) * 1e3,
            "stored_G_colsumsq_gbs": g_bytes_local / dev_ms(lambda: eng.colsumsq_dev(w_local)) / 1e6,
        }
        del sub, ws

    # ---- device-resident Gauss-Newton CG: 5 iterations of
    #      q = 2 G^T W^2 G p + beta p with all vectors in HBM, one all-reduce per iteration;
    #      host vectors only for the gradient in and the step out (inside the timing) ---------
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

    cg_iters = 5
    jtj = sim.getJtJdiag(m_host)                          # stored-G column norms (not timed)
    solver = DeviceGaussNewtonCG(sim, data_weights=np.ones(nD), reg_hessian=np.ones(nP),
                                 beta=1e-3 * float(jtj.mean()), cg_maxiter=cg_iters, cg_rtol=0.0,
                                 cg_atol=0.0)
    solver.set_model(m_host)
    solver.jacobi_preconditioner()                        # UpdatePreconditioner (cached diagonal)
    g_host = rng.uniform(-1.0, 1.0, nP)
    solver.find_search_direction(g_host)                  # warm-up
    cg_t = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        cg_step = solver.find_search_direction(g_host)
        barrier()
        cg_t.append(max_over_ranks(time.perf_counter() - t0))
    cg_s = float(np.mean(cg_t))
    cg = {"iterations": int(solver.cg_count), "ms_per_iteration": 1e3 * cg_s / max(solver.cg_count, 1),
          "jvec_jtvec_gbs": 2 * g_bytes_total * solver.cg_count / cg_s / 1e9,
          "rel_residual": float(solver.cg_rel_resid), "checksum": float(np.sum(cg_step)),
          "what": "DeviceGaussNewtonCG.find_search_direction: gradient H2D, 5 x (gemv, row weights, "
                  "gemv_t, all-reduce, fused vector kernels), step D2H; wall clock, max over ranks"}
    del solver

    # ---- end to end through the public API, host buffers in and out -------------
    del sim, eng, x_dev, w_local, step, timers
    gc.collect()
    torch.cuda.empty_cache()
    e2e_build, e2e_prod = [], []
    h2d = d2h = 0
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for it in range(e2e_steps + 1):                       # first pass is a warm-up
        barrier()
        t0 = time.perf_counter()
        sim2 = make_simulation(wl, make_survey(wl, wl["locations"]), comm=coll)
        sim2.G                                             # plan + upload + build kernel
        barrier()
        t1 = time.perf_counter()
        jv = sim2.Jvec(m_host, v_host)                     # H2D nP doubles, D2H nD doubles
        jtv = sim2.Jtvec(m_host, w_host)                   # H2D nD doubles, D2H nP doubles
        barrier()
        t2 = time.perf_counter()
        if it > 0:
            e2e_build.append(max_over_ranks(t1 - t0))
            e2e_prod.append(max_over_ranks(t2 - t1))
        plan = sim2._dev.host_plan
        plan_arrays = [plan.tile_xyz, plan.tile_cell_nodes, plan.tile_cell_col, plan.tile_n_nodes,
                       plan.grid_colmap, plan.footprints4, plan.footprints2] + list(plan.grid_axes or [])
        h2d = (sum(a.nbytes for a in plan_arrays if a is not None)
               + sim2._sharding.local_rows.size // len(wl["components"]) * 24
               # Jvec argument: every rank uploads its 1/world slice (all-gathered over NVLink)
               + 8 * (-(-nP // world) if (world > 1 and nP >= SLICED_UPLOAD_MIN) else nP) + 8 * nD)
        d2h = 8 * (nD + nP)
        checksum = float(np.sum(jv)) + float(np.sum(jtv))
        del sim2, plan
        gc.collect()          # G goes back to torch's caching allocator (no cudaFree / cudaMalloc per step)
    e2e_value = pairs_total / float(np.mean(e2e_build))
    e2e_prod_gbs = 2 * g_bytes_total / float(np.mean(e2e_prod)) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    n_nodes = wl["mesh"].n_nodes
    flop_eq_per_pair = FLOP_EQ_PER_NODE * n_nodes / nC + FLOP_EQ_PER_CELL
    build_tflops = (pairs_local * K / t_build) * flop_eq_per_pair / 1e12
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "g_storage": args.g_dtype,
        "config": workload_config(wl, world),
        "ms_build": 1e3 * t_build / K, "ms_jvec": 1e3 * t_jvec / K, "ms_jtvec": 1e3 * t_jtvec / K,
        "products": {"metric": "Jvec+Jtvec GB/s", "jvec_jtvec_gbs": prod_gbs,
                     "frac_of_hbm": prod_gbs / (hbm_peak * world), "G_bytes": g_bytes_total},
        "roofline": {
            "kernel": "pf_march_kernel<12, PROG_TMI_VECTOR, SINK_STORE, double, 4> (G build)", "bound": "fp64",
            "achieved": build_tflops, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": build_tflops / (fp64_peak / 1e12),
            # ncu (profiles/r1g_kernels_summary.md): 7.135 GB written + 0.099 GB read for a
            # 7.104 GB block of G -> DRAM bytes per launch = 1.0183 x the block it writes
            "traffic": 1.0183 * g_bytes_local,
            "note": f"algorithmic flop-eq/pair = 339*nN/nC + 64 = {flop_eq_per_pair:.1f} (SURVEY 8(d)); "
                    "peak = DFMA probe measured in this run (pf_probe_fp64_peak); per GPU; "
                    "ncu: FP64 pipe 58-64 % busy (the kernel issues ~260 FP64 instructions per pair)",
        },
        "roofline_products": {
            "kernel": "gemv_kernel<double> + column_kernel<double,0> (Jvec + Jtvec)", "bound": "hbm",
            "achieved": g_bytes_local * 2 * K / (t_jvec + t_jtvec) / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": g_bytes_local * 2 * K / (t_jvec + t_jtvec) / 1e9 / hbm_peak,
            # ncu: 7.128 GB (gemv) / 7.118 GB (gemv_t) DRAM traffic for a 7.104 GB block
            "traffic": 1.0034 * g_bytes_local,
            "note": f"algorithmic bytes = rows*cols*8 per product, per GPU; traffic = DRAM bytes per "
                    f"launch scaled from the ncu capture; peak = {hbm_src}",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "jvec_jtvec_gbs": e2e_prod_gbs,
                "ms_build": 1e3 * float(np.mean(e2e_build)), "ms_products": 1e3 * float(np.mean(e2e_prod)),
                "steps": e2e_steps, "checksum": checksum,
                "what": "new Simulation3DIntegral(engine='cuda') -> sim.G (tile plan on host, upload, "
                        "build) ; sim.Jvec / sim.Jtvec with numpy vectors in and out"},
        "gpu_launches": int(launches), "clocks": clocks, "other_sinks": extras,
        "gauss_newton_cg": cg,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = {k: v for k, v in cpu_baseline(wl).items() if k != "seconds"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg4-shard")
    ap.add_argument("--receivers-per-gpu", type=int, default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--g-dtype", default="f64", choices=["f64", "f32"],
                    help="storage type of G (sensitivity_dtype); BASELINE.json asks for fp64")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
