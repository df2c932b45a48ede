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
    """`nvidia-smi -lms 100` on this rank's GPU.  Started BEFORE the warm-up steps (on an 8-GPU box
    nvidia-smi needs longer to come up than the timed region lasts); `stop(window)` keeps the
    samples whose timestamps fall inside the timed region and, when the region was too short to
    catch any, the samples of warm-up + timed region (the same steps) -- `window` says which."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
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

    def stop(self, window=None):
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                stamp = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((stamp, float(parts[1]), float(parts[2]), float(parts[3]),
                             [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                 "sw_power_cap"), parts[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        used, which = rows, "whole run"
        if window is not None:
            inside = [r for r in rows if window[0] <= r[0] <= window[1]]
            used, which = (inside, "timed region") if inside else (rows, "warm-up + timed region")
        sm, mx, power = [r[1] for r in used], [r[2] for r in used], [r[3] for r in used]
        reasons = sorted({n for r in used for n in r[4]})
        # "under load" = samples drawing more than half of the peak power seen
        pmax = max(power)
        loaded = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(mx)), "reasons": reasons,
                "samples": len(sm), "power_w_max": pmax, "window": which}


# --------------------------------------------------------------------------- CPU baseline
def cpu_baseline(wl, sample_rx=None, repeats=1, threads=None):
    """
    The oracle port (oracle/pf_oracle.c: C + OpenMP restatement of the reference's
    numba-parallel loops) timed on the host cores on a bounded receiver
    subsample x the full mesh; products timed with numpy on the sub-matrix
    (the reference's own expressions, gravity/simulation.py:387, 418).
    """
    from oracle import oracle

    mesh = wl["mesh"]
    if threads:
        oracle.set_num_threads(threads)       # torchrun exports OMP_NUM_THREADS=1: ask explicitly
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
    try:                                       # numpy's BLAS also obeys OMP_NUM_THREADS=1 under torchrun
        from threadpoolctl import threadpool_limits
        blas = threadpool_limits(limits=cores)
    except Exception:  # pragma: no cover
        blas = None
    G @ v
    t = time.perf_counter()
    G @ v
    G.T @ w
    tprod = time.perf_counter() - t
    if blas is not None:
        blas.restore_original_limits()
    return {"value": pairs / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done_rx} receivers (chunks of {sample_rx}) x full mesh ({nC} cells, "
                      f"{sim.nP} columns), oracle/pf_oracle.c OpenMP, {best:.2f} s",
            "jvec_jtvec_gbs": 2 * G.nbytes / tprod / 1e9, "seconds": best}


def host_cores():
    """Host threads this process may use (affinity mask, not OMP_NUM_THREADS)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, args.gpus, args.receivers_per_gpu)
    threads = host_cores()                    # all the host threads this process may use
    for _ in range(max(args.warmup, 1)):
        cpu_baseline(wl, sample_rx=8, repeats=0, threads=threads)
    times, last = [], None
    for _ in range(args.steps):
        t = time.perf_counter()
        last = cpu_baseline(wl, threads=threads)
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


# --------------------------------------------------------------------------- parity (checker)
PARITY_TOL = 1e-9          # north_star: relative 1e-9 in fp64, norm-wise (max|d| / max|ref|)


def _rel(a, b):
    den = float(np.max(np.abs(b))) if b.size else 0.0
    return float(np.max(np.abs(np.asarray(a) - b)) / den) if den > 0 else float(np.max(np.abs(a), initial=0.0))


def port_rows(physics, nodes, cell_nodes, locations, components, b0, model_type, threads):
    """Rows of G for a few receivers from the CPU port (oracle/pf_oracle.c) -- the checker."""
    from oracle import oracle

    oracle.set_num_threads(max(1, int(threads)))
    sim = oracle.OracleSimulation(physics, nodes, cell_nodes, [(locations, components)], b0=b0,
                                  model_type=model_type, backend="c", parallel=True)
    return sim.sensitivity_matrix()


def sample_positions(n_local_rx, n_sample):
    n = min(n_sample, n_local_rx)
    return np.unique(np.linspace(0, n_local_rx - 1, n).round().astype(np.int64)) if n else np.zeros(0, np.int64)


def stored_g_parity(wl, sim, coll, b0, n_sample, torch, dist):
    """
    Compare, OUTSIDE the timed region, the GPU results with the CPU port on a receiver sample of
    EVERY rank: the rank-local rows of G, Jvec on those rows, and the all-reduced Jtvec /
    column norms for a vector that is non-zero on the sample rows of all ranks only (so that the
    exact answer is the sum over ranks of the port's G_s^T w_s).  Returns the parity dict; the
    maxima are taken over ranks.
    """
    eng, sh = sim._dev, sim._sharding
    C = len(wl["components"])
    nP, nD = eng.nP, sh.nD
    world, rank = coll.world_size, coll.rank
    loc_local = sh.local_receiver_list[0][0]
    pos = sample_positions(loc_local.shape[0], n_sample)
    rows_local = (pos[:, None] * C + np.arange(C)[None, :]).reshape(-1)
    rows_global = sh.local_rows[rows_local]
    nodes, cell_nodes = sim._get_active_nodes()
    Gs = port_rows(wl["physics"], nodes, cell_nodes, loc_local[pos], wl["components"], b0,
                   wl["model_type"], max(1, host_cores() // world))
    # rows of G
    Gd = eng.G[torch.from_numpy(rows_local).to(eng.device), :nP].double().cpu().numpy()
    g_row = float(np.max(np.max(np.abs(Gd - Gs), axis=1) / np.max(np.abs(Gs), axis=1)))
    # Jvec on the sample rows (every rank holds the gathered vector)
    rng = np.random.default_rng(11)
    x = rng.uniform(-1.0, 1.0, nP)
    full = coll.all_gather_rows(eng.gemv_dev(eng.to_device(x)), sh)
    jvec = _rel(full[torch.from_numpy(rows_global).to(eng.device)].cpu().numpy(), Gs @ x)
    # Jtvec / diag with weights living on the sample rows of all ranks
    w = np.zeros(nD)
    for r in range(world):
        rr = sh.rows_of_rank(r)
        n_loc_r = rr.size // C
        pr = sample_positions(n_loc_r, n_sample)
        w[rr[(pr[:, None] * C + np.arange(C)[None, :]).reshape(-1)]] = 1.0
    w *= np.random.default_rng(12).uniform(0.5, 1.5, nD)
    w_local = sim._local_rows_of(w)
    z = coll.all_reduce_sum(eng.gemv_t_dev(w_local)).cpu().numpy()
    d = coll.all_reduce_sum(eng.colsumsq_dev(w_local)).cpu().numpy()
    ws = w[rows_global]
    ref = torch.from_numpy(np.stack([Gs.T @ ws, np.einsum("i,ij,ij->j", ws, Gs, Gs)])).to(eng.device)
    if world > 1:
        dist.all_reduce(ref)          # transport for the checker's partial sums only
    ref = ref.cpu().numpy()
    vals = torch.tensor([g_row, jvec, _rel(z, ref[0]), _rel(d, ref[1])], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    g_row, jvec, jtvec, diag = (float(v) for v in vals.cpu())
    return {"G_rowwise_max": g_row, "jvec": jvec, "jtvec": jtvec, "diag": diag, "tol": PARITY_TOL,
            "n_rx": int(pos.size) * world, "ranks": world, "against": "oracle/pf_oracle.c (port)",
            "ok": bool(max(g_row, jvec, jtvec, diag) <= PARITY_TOL)}


# --------------------------------------------------------------------------- measured DRAM traffic
TRAFFIC_TABLES = ("r2z_traffic.json", "r2_traffic.json")   # newest capture of the round first


def measured_traffic(kernel_name, algorithmic_bytes):
    """DRAM bytes per launch of the dominant kernel from THIS round's committed ncu capture
    (profiles/r2z_traffic.json -- the final kernels of the round -- else r2_traffic.json, written by
    scripts/ncu_traffic.py from `ncu --set full`), scaled by the algorithmic bytes of this run; None
    when the capture does not cover the kernel."""
    table = None
    for name in TRAFFIC_TABLES:
        try:
            table = json.load(open(ROOT / "profiles" / name))
            break
        except (OSError, ValueError):
            continue
    if table is None:
        return None, "no ncu capture committed for this round"
    for entry in table.get("kernels", []):
        if entry["match"] in kernel_name:
            ratio = entry["dram_bytes"] / entry["algorithmic_bytes"]
            return ratio * algorithmic_bytes, (f"dram__bytes_read+write = {ratio:.4f} x algorithmic bytes in "
                                               f"{entry['source']}")
    return None, f"kernel not in profiles/{name}"


def dev_ms(torch, fn, reps=3):
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


# --------------------------------------------------------------------------- other BASELINE configs
# SURVEY.md 8(d) bookkeeping per kernel set: (flop-eq per node, flop-eq per cell)
FLOP_EQ = {"second_order": (339.0, 64.0), "gz_only": (177.0, 8.0), "gravity7": (339.0, 69.0)}


def flop_eq_per_pair(n_nodes, n_cells, kind="second_order"):
    per_node, per_cell = FLOP_EQ[kind]
    return per_node * n_nodes / n_cells + per_cell


def baseline_configs(torch, fp64_peak, hbm_peak):
    """
    One compact entry per BASELINE.json config, each at its per-GPU size, through the public API,
    each with its own throughput, roofline fraction and a parity number against the CPU port on a
    receiver sample (<= ~2 s of GPU time per entry).  cfg4 is the headline workload itself.
    """
    import scipy.sparse as sp

    from simpeg_b200 import maps
    from simpeg_b200.meshes import TensorMesh, synthetic_octree
    from simpeg_b200.potential_fields import gravity, magnetics
    from simpeg_b200.utils import dip_azimuth2cartesian

    threads = host_cores()
    out = {}

    def release():
        gc.collect()
        torch.cuda.empty_cache()

    def sync_time(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t, r

    # ---- cfg1: gravity gz, 40x40x20, 21x21 receivers, stored G (the reference's CPU-runnable case)
    wl = make_workload("cfg1", 1, None)
    survey = make_survey(wl, wl["locations"])
    sim = make_simulation(wl, survey)
    nC = wl["mesh"].n_cells
    rho = np.random.default_rng(3).uniform(-0.5, 0.5, nC)
    sim.G
    t_build = dev_ms(torch, sim._dev.build_G)
    pos = sample_positions(wl["n_rx"], 16)
    Gs = port_rows("gravity", wl["mesh"].nodes, wl["mesh"].cell_nodes, wl["locations"][pos], ["gz"], None,
                   "scalar", threads)
    J = sim.getJ(rho)
    out["cfg1"] = {
        "what": "gravity gz forward + getJ, TensorMesh 40x40x20, 441 receivers, store='ram' (full config)",
        "build_pairs_s": wl["n_rx"] * nC / t_build * 1e3, "ms_build": t_build,
        "kernel": sim._dev.describe("store"),
        "frac_fp64": wl["n_rx"] * nC / t_build * 1e3 * flop_eq_per_pair(wl["mesh"].n_nodes, nC, "gz_only") / fp64_peak,
        "flop_eq_per_pair": flop_eq_per_pair(wl["mesh"].n_nodes, nC, "gz_only"),
        "parity": max(_rel(np.asarray(J)[pos], Gs), _rel(sim.dpred(rho)[pos], Gs @ rho)),
        "getJ_is_ndarray": isinstance(J, np.ndarray),
    }
    del sim, J
    release()

    # ---- cfg2-shard: gravity 6 tensor components + gz, 100x100x50, 1250 receivers (35 GB G)
    wl = make_workload("cfg2-shard", 1, None)
    survey = make_survey(wl, wl["locations"])
    sim = make_simulation(wl, survey)
    eng = sim._dev
    nC, C = wl["mesh"].n_cells, len(wl["components"])
    sim.G
    t_build = dev_ms(torch, eng.build_G)
    g_bytes = eng.n_rows * nC * 8
    xd = eng.to_device(np.random.default_rng(4).uniform(-1, 1, nC))
    wd = eng.to_device(np.random.default_rng(5).uniform(0, 1, eng.n_rows))
    t_prod = dev_ms(torch, lambda: (eng.gemv_dev(xd), eng.gemv_t_dev(wd)))
    pos = sample_positions(wl["n_rx"], 4)
    rows = (pos[:, None] * C + np.arange(C)[None, :]).reshape(-1)
    Gs = port_rows("gravity", wl["mesh"].nodes, wl["mesh"].cell_nodes, wl["locations"][pos],
                   wl["components"], None, "scalar", threads)
    Gd = eng.G[torch.from_numpy(rows).to(eng.device), :nC].cpu().numpy()
    out["cfg2-shard"] = {
        "what": "gravity gxx,gxy,gxz,gyy,gyz,gzz,gz, TensorMesh 100x100x50, 1250 receivers per GPU "
                "(8750 rows, 35 GB fp64 G): one-pass build + Jvec/Jtvec",
        "build_pairs_s": wl["n_rx"] * nC / t_build * 1e3, "ms_build": t_build,
        "kernel": eng.describe("store"),
        "frac_fp64": wl["n_rx"] * nC / t_build * 1e3 * flop_eq_per_pair(wl["mesh"].n_nodes, nC, "gravity7") / fp64_peak,
        "flop_eq_per_pair": flop_eq_per_pair(wl["mesh"].n_nodes, nC, "gravity7"),
        "jvec_jtvec_gbs": 2 * g_bytes / t_prod / 1e6, "frac_hbm": 2 * g_bytes / t_prod / 1e6 / hbm_peak,
        "parity": float(np.max(np.max(np.abs(Gd - Gs), axis=1) / np.max(np.abs(Gs), axis=1))),
    }
    del xd, wd, sim, eng, Gd
    release()

    # ---- cfg3-shard: TMI scalar forward_only on 200x200x100 (4 M cells), 5000 receivers
    n_rx = 5000
    mesh = TensorMesh([np.full(200, 25.0), np.full(200, 25.0), np.full(100, 25.0)], "CCN")
    side = int(np.ceil(np.sqrt(n_rx)))
    xs = np.linspace(-2500.0, 2500.0, side)
    X, Y = np.meshgrid(xs, xs)
    loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, 30.0)][:n_rx]
    rx = magnetics.receivers.Point(loc, components="tmi")
    src = magnetics.sources.UniformBackgroundField([rx], 50000.0, 60.0, 250.0)
    survey = magnetics.survey.Survey(src)
    nC = mesh.n_cells
    rng = np.random.default_rng(4)
    chi = rng.uniform(0.0, 0.05, nC)
    sim = magnetics.Simulation3DIntegral(mesh, survey=survey, chiMap=maps.IdentityMap(nP=nC),
                                         store_sensitivities="forward_only")
    pos = sample_positions(n_rx, 4)
    b0 = 50000.0 * dip_azimuth2cartesian(60.0, 250.0).squeeze()
    Gs = port_rows("magnetics", mesh.nodes, mesh.cell_nodes, loc[pos], ["tmi"], b0, "scalar", threads)
    v = np.zeros(n_rx)
    v[pos] = rng.uniform(0.5, 1.5, pos.size)
    # warm-up: tile plan, node table and tile arrays are built / uploaded on first use (once per mesh)
    sim.dpred(chi)
    sim.Jtvec(chi, v)
    sim.getJtJdiag(chi, W=sp.diags(np.sqrt(0.5 * v)))
    def best_of_two(fn_a, fn_b):
        (ta, ra), (tb, _) = sync_time(fn_a), sync_time(fn_b)
        return min(ta, tb), ra

    t_fwd, d = best_of_two(lambda: sim.dpred(chi), lambda: sim.dpred(chi))
    t_jt, jt = best_of_two(lambda: sim.Jtvec(chi, v), lambda: sim.Jtvec(chi, v))
    # different weights per call: the sha256 cache of getJtJdiag must not answer
    t_dg, dg = best_of_two(lambda: sim.getJtJdiag(chi, W=sp.diags(np.sqrt(v))),
                           lambda: sim.getJtJdiag(chi, W=sp.diags(np.sqrt(0.25 * v))))
    pairs = n_rx * nC
    fe = flop_eq_per_pair(mesh.n_nodes, nC)
    out["cfg3-shard"] = {
        "what": "magnetic TMI scalar, store_sensitivities='forward_only', TensorMesh 200x200x100 (4 M cells), "
                "5000 receivers per GPU; public API with numpy vectors (wall clock)",
        "dpred_pairs_s": pairs / t_fwd, "jtvec_pairs_s": pairs / t_jt, "getJtJdiag_pairs_s": pairs / t_dg,
        "kernels": {"dpred": sim._dev.describe("forward"), "jtvec": sim._dev.describe("gtv"),
                    "diag": sim._dev.describe("diag")},
        "frac_fp64": {"dpred": pairs / t_fwd * fe / fp64_peak, "jtvec": pairs / t_jt * fe / fp64_peak,
                      "diag": pairs / t_dg * fe / fp64_peak},
        "parity": max(_rel(d[pos], Gs @ chi), _rel(jt, Gs.T @ v[pos]),
                      _rel(dg, np.einsum("i,ij,ij->j", v[pos], Gs, Gs))),
    }
    del Gs, sim
    release()

    # ---- cfg5-shard: gravity gz on an octree below topography, 1.96 M cells, 6250 receivers (98 GB G)
    def topo(x, y):
        return 50.0 * np.exp(-0.5 * ((x / 100.0) ** 2 + (y / 100.0) ** 2)) - 50.0

    t0 = time.perf_counter()
    mesh = synthetic_octree((140, 140, 12), 10.0, levels=3, topo=topo)
    t_mesh = time.perf_counter() - t0
    n_rx = 6250
    rng = np.random.default_rng(6)
    half = 140 * 40.0 / 2
    xy = rng.uniform(-half, half, (n_rx, 2))
    loc = np.c_[xy, topo(xy[:, 0], xy[:, 1]) + 5.0]
    rx = gravity.receivers.Point(loc, components=["gz"])
    survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
    nC = mesh.n_cells
    sim = gravity.Simulation3DIntegral(mesh, survey=survey, rhoMap=maps.IdentityMap(nP=nC),
                                       sensitivity_dtype=np.float64)
    t_plan, eng = sync_time(lambda: sim._dev)          # generic tile plan (host, native) + uploads
    t_build, _ = sync_time(lambda: sim.G)
    t_build2 = dev_ms(torch, eng.build_G, reps=2) * 1e-3
    wts = rng.uniform(0.5, 1.5, n_rx)
    m = rng.uniform(-0.5, 0.5, nC)
    sim.getJtJdiag(m, W=sp.diags(wts))
    t_diag, diag = sync_time(lambda: sim.getJtJdiag(m, W=sp.diags(wts * 1.0001)))
    g_bytes = n_rx * nC * 8
    # well-posed Gauss-Newton system: gradient of the misfit of a synthetic anomaly
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

    m_true = np.zeros(nC)
    cc = mesh.cell_centers
    m_true[np.linalg.norm(cc - np.array([0.0, 0.0, -150.0]), axis=1) < 120.0] = 0.5
    grad = 2.0 * sim.Jtvec(m, wts**2 * (sim.dpred(m) - sim.dpred(m_true)))   # misfit gradient at m
    solver = DeviceGaussNewtonCG(sim, data_weights=wts, reg_hessian=np.ones(nC), beta=1e-3 * float(diag.mean()),
                                 cg_maxiter=10, cg_rtol=0.0)
    solver.set_model(m)
    solver.jacobi_preconditioner()
    solver.find_search_direction(grad)
    t_cg, _ = sync_time(lambda: solver.find_search_direction(grad))
    pos = sample_positions(n_rx, 4)
    nodes, cell_nodes = sim._get_active_nodes()
    Gs = port_rows("gravity", nodes, cell_nodes, loc[pos], ["gz"], None, "scalar", threads)
    Gd = eng.G[torch.from_numpy(pos).to(eng.device), :nC].cpu().numpy()
    out["cfg5-shard"] = {
        "what": "gravity gz, synthetic octree below topography (1.96 M active cells, hanging nodes), 6250 "
                "receivers per GPU (98 GB fp64 G): build, getJtJdiag, device Gauss-Newton CG (public API)",
        "cells": nC, "mesh_s": t_mesh, "plan_upload_s": t_plan, "tiles": int(eng.host_plan.n_tiles),
        "build_pairs_s": n_rx * nC / t_build2, "ms_build": 1e3 * t_build2, "first_build_s": t_build,
        "kernel": eng.describe("store"),
        "frac_fp64": n_rx * nC / t_build2 * flop_eq_per_pair(nodes.shape[0], nC, "gz_only") / fp64_peak,
        "flop_eq_per_pair": flop_eq_per_pair(nodes.shape[0], nC, "gz_only"),   # unique nodes, not tile slots
        "node_evaluations_per_pair": float(eng.host_plan.nodes_per_cell),
        "getJtJdiag_gbs": g_bytes / t_diag / 1e9, "frac_hbm_getJtJdiag": g_bytes / t_diag / 1e9 / hbm_peak,
        "cg_ms_per_iteration": 1e3 * t_cg / max(solver.cg_count, 1),
        "cg_jvec_jtvec_gbs": 2 * g_bytes * solver.cg_count / t_cg / 1e9,
        "cg_rel_residual": float(solver.cg_rel_resid),
        "parity": float(np.max(np.max(np.abs(Gd - Gs), axis=1) / np.max(np.abs(Gs), axis=1))),
    }
    del solver, sim, eng, Gd
    release()
    return out


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
    b0 = np.asarray(survey.source_field.b0) if wl["physics"] == "magnetics" else None

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
    ops = sim._ops
    token = torch.zeros(1, dtype=torch.float64, device=eng.device)

    def step(timers=None):
        # events: build | rank alignment | Jvec kernel | all-gather | Jtvec kernel + all-reduce.
        # The alignment is a one-element all-reduce ON THE DEVICE (no host sync): it makes every
        # rank wait for the slowest build, so that build skew is reported as `ms_skew` instead
        # of leaking into the Jvec timer.
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        eng.build_G()
        ev[1].record()
        coll.all_reduce_sum(token)
        ev[2].record()
        y = eng.gemv_dev(x_dev)                          # Jvec: local rows ...
        ev[3].record()
        full = coll.all_gather_rows(y, sim._sharding)    # ... gathered on every rank
        ev[4].record()
        z = ops.gemv_t_allreduce_dev(w_local)            # Jtvec: column partials summed over ranks
        ev[5].record()
        if timers is not None:
            timers.append(ev)
        return full, z

    import datetime

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    fp64_peak = probe_fp64_peak()
    barrier()
    t_region0 = datetime.datetime.now()
    launches0 = launch_count()
    timers = []
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(args.steps):
        step(timers)
    e_stop.record()
    barrier()
    launches = launch_count() - launches0
    clocks = sampler.stop((t_region0, datetime.datetime.now()))
    K = args.steps

    def part(i, j):
        return max_over_ranks(sum(e[i].elapsed_time(e[j]) for e in timers)) * 1e-3

    t_total = max_over_ranks(e_start.elapsed_time(e_stop)) * 1e-3
    t_build, t_skew = part(0, 1), part(1, 2)
    t_gemv, t_gather, t_jtvec = part(2, 3), part(3, 4), part(4, 5)
    t_jvec = t_gemv + t_gather
    value = pairs_total * K / t_build
    prod_gbs = 2 * g_bytes_total * K / (t_jvec + t_jtvec) / 1e9
    # the Jtvec kernel alone (no collective), for the cost of the all-reduce inside ms_jtvec
    t_gemv_t_alone = max_over_ranks(dev_ms(torch, lambda: eng.gemv_t_dev(w_local))) * 1e-3

    # ---- parity against the CPU port on a receiver sample of every rank (not timed) --------
    parity = stored_g_parity(wl, sim, coll, b0, args.parity_rx, torch, dist)

    # ---- the other sinks of the path (matrix-free forward / G^T v / diag, weighted column
    #      norms of the stored G), timed on the device on a receiver subsample ------------
    extras = {}
    if rank == 0:
        from simpeg_b200.engine import CudaEngine

        n_sub = min(256, eng.n_rows // len(wl["components"]))
        sub = CudaEngine(wl["physics"], None, None, [(wl["locations"][:n_sub], wl["components"])],
                         b0=b0, model_type=wl["model_type"], plan=eng.host_plan)
        sub_pairs = n_sub * nC
        ws = sub.to_device(w_host[: sub.n_rows])
        extras = {
            "sample": f"{n_sub} receivers x full mesh, device-resident vectors, best of 3",
            "forward_only_dpred_pairs_s": sub_pairs / dev_ms(torch, lambda: sub.forward_dev(x_dev)) * 1e3,
            "forward_only_gt_dot_v_pairs_s": sub_pairs / dev_ms(torch, lambda: sub.gt_dot_v_dev(ws)) * 1e3,
            "forward_only_gtg_diag_pairs_s": sub_pairs / dev_ms(torch, lambda: sub.gtg_diag_dev(ws)) * 1e3,
            "stored_G_colsumsq_gbs": g_bytes_local / dev_ms(torch, lambda: eng.colsumsq_dev(w_local)) / 1e6,
            "kernels": {"forward": sub.describe("forward"), "gt_dot_v": sub.describe("gtv"),
                        "gtg_diag": sub.describe("diag")},
        }
        del sub, ws

    # ---- device-resident Gauss-Newton CG: 5 iterations of
    #      q = 2 G^T W^2 G p + beta p with all vectors in HBM, one all-reduce per iteration;
    #      host vectors only for the gradient in and the step out (inside the timing).
    #      The system is a real Gauss-Newton step: the gradient is that of the misfit of a
    #      synthetic anomaly, g = -2 J^T W^2 (d_obs - d(m)), so it lies in range(J^T) and the
    #      preconditioned CG residual must go DOWN (checked below) -------------------------
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

    cg_iters = 5
    jtj = sim.getJtJdiag(m_host)                          # stored-G column norms (not timed)
    cc = wl["mesh"].cell_centers
    m_true = np.zeros(nP)
    blob = np.linalg.norm(cc - np.array([0.0, 0.0, -0.4 * wl["shape"][2] * wl["h"]]), axis=1) \
        < 0.15 * wl["shape"][0] * wl["h"]
    m_true[:nC][blob] = 0.05                              # first column block (x component / scalar)
    d_obs = sim.dpred(m_true)
    wd = np.full(nD, 1.0 / max(float(np.abs(d_obs).max()) * 0.02, 1e-30))   # 2 % of the peak anomaly
    g_host = 2.0 * sim.Jtvec(m_host, wd**2 * (sim.dpred(m_host) - d_obs))   # gradient of the misfit at m_host
    solver = DeviceGaussNewtonCG(sim, data_weights=wd, reg_hessian=np.ones(nP),
                                 beta=1e-3 * float((wd[0] ** 2) * jtj.mean()), cg_maxiter=cg_iters, cg_rtol=0.0,
                                 cg_atol=0.0)
    solver.set_model(m_host)
    solver.jacobi_preconditioner()                        # UpdatePreconditioner (cached diagonal)
    solver.find_search_direction(g_host)                  # warm-up
    cg_t = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        cg_step = solver.find_search_direction(g_host)
        barrier()
        cg_t.append(max_over_ranks(time.perf_counter() - t0))
    cg_s = float(np.median(cg_t))         # wall clock with one host sync per iteration: median of 3
    cg = {"iterations": int(solver.cg_count), "ms_per_iteration": 1e3 * cg_s / max(solver.cg_count, 1),
          "jvec_jtvec_gbs": 2 * g_bytes_total * solver.cg_count / cg_s / 1e9,
          "ms_per_iteration_runs": [1e3 * t / max(solver.cg_count, 1) for t in cg_t],
          "rel_residual": float(solver.cg_rel_resid), "residual_decreased": bool(solver.cg_rel_resid < 1.0),
          "checksum": float(np.sum(cg_step)),
          "what": "DeviceGaussNewtonCG.find_search_direction on g = -2 J^T W^2 r of a synthetic anomaly: "
                  "gradient H2D, 5 x (gemv, row weights, gemv_t + all-reduce, fused vector kernels), step "
                  "D2H; wall clock, max over ranks"}
    del solver
    kernel_build = eng.describe("store")

    # ---- end to end through the public API, host buffers in and out -------------
    del sim, eng, x_dev, w_local, step, timers, ops
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

    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"

    configs = None
    if world == 1 and args.configs:
        configs = baseline_configs(torch, fp64_peak, hbm_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if not parity["ok"]:
            sys.exit(3)
        return

    n_nodes = wl["mesh"].n_nodes
    fe_pair = flop_eq_per_pair(n_nodes, nC)
    build_tflops = (pairs_local * K / t_build) * fe_pair / 1e12
    traffic_b, traffic_b_src = measured_traffic(kernel_build, g_bytes_local)
    traffic_p, traffic_p_src = measured_traffic("gemv_kernel", g_bytes_local)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "g_storage": args.g_dtype,
        "config": workload_config(wl, world),
        "ms_build": 1e3 * t_build / K, "ms_skew": 1e3 * t_skew / K,
        "ms_jvec": 1e3 * t_jvec / K, "ms_jvec_kernel": 1e3 * t_gemv / K, "ms_allgather": 1e3 * t_gather / K,
        "ms_jtvec": 1e3 * t_jtvec / K, "ms_jtvec_kernel_alone": 1e3 * t_gemv_t_alone,
        "ms_allreduce_exposed": 1e3 * (t_jtvec / K - t_gemv_t_alone),
        "products": {"metric": "Jvec+Jtvec GB/s", "jvec_jtvec_gbs": prod_gbs,
                     "frac_of_hbm": prod_gbs / (hbm_peak * world), "G_bytes": g_bytes_total,
                     "per_gpu_gbs": prod_gbs / world},
        "parity": parity,
        "roofline": {
            "kernel": kernel_build + " (G build)", "bound": "fp64",
            "achieved": build_tflops, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": build_tflops / (fp64_peak / 1e12),
            "traffic": traffic_b, "traffic_source": traffic_b_src,
            "note": f"algorithmic flop-eq/pair = 339*nN/nC + 64 = {fe_pair:.1f} (SURVEY 8(d)); "
                    "peak = DFMA probe measured in this run (pf_probe_fp64_peak); per GPU",
        },
        "roofline_products": {
            "kernel": "gemv_kernel + column_kernel (Jvec + Jtvec)", "bound": "hbm",
            "achieved": g_bytes_local * 2 * K / (t_jvec + t_jtvec) / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": g_bytes_local * 2 * K / (t_jvec + t_jtvec) / 1e9 / hbm_peak,
            "traffic": traffic_p, "traffic_source": traffic_p_src,
            "note": f"algorithmic bytes = rows*cols*{g_item} per product, per GPU; peak = {hbm_src}",
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
    if configs is not None:
        out["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = {k: v for k, v in cpu_baseline(wl, threads=host_cores()).items() if k != "seconds"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    bad = [k for k, v in (configs or {}).items() if not v["parity"] <= PARITY_TOL]
    if not parity["ok"] or bad:
        sys.stderr.write(f"PARITY FAILURE: {parity} {bad}\n")
        sys.exit(3)


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
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="skip the per-BASELINE-config block (cfg1, cfg2-shard, cfg3-shard, cfg5-shard; N=1 only)")
    ap.add_argument("--parity-rx", type=int, default=8,
                    help="receivers per rank checked against the CPU port after the timed loop")
    ap.add_argument("--g-dtype", default="f64", choices=["f64", "f32"],
                    help="storage type of G (sensitivity_dtype); BASELINE.json asks for fp64")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
