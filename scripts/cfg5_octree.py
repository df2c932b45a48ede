"""
BASELINE config 5 at its per-GPU size: gravity gz on a synthetic octree below Gaussian-hill
topography (1.96 M active cells incl. hanging nodes), 6250 receivers (= 50 000 / 8) draped 5 m
above the topography, fp64 G (98 GB): G build (tile kernel), getJtJdiag from the stored G, and
10 iterations of the device-resident Gauss-Newton CG (Jtvec(W^2 Jvec(p)) + beta p).
Through the public API; prints one JSON line.
"""
import json
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, ".")
from simpeg_b200 import maps  # noqa: E402
from simpeg_b200.gauss_newton import DeviceGaussNewtonCG  # noqa: E402
from simpeg_b200.meshes import synthetic_octree  # noqa: E402
from simpeg_b200.potential_fields import gravity  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 140
n_rx = int(sys.argv[2]) if len(sys.argv) > 2 else 6250


def topo(x, y):
    return 50.0 * np.exp(-0.5 * ((x / 100.0) ** 2 + (y / 100.0) ** 2)) - 50.0


t = time.perf_counter()
mesh = synthetic_octree((n, n, 12), 10.0, levels=3, topo=topo)
t_mesh = time.perf_counter() - t
rng = np.random.default_rng(6)
half = n * 40.0 / 2
xy = rng.uniform(-half, half, (n_rx, 2))
loc = np.c_[xy, topo(xy[:, 0], xy[:, 1]) + 5.0]
rx = gravity.receivers.Point(loc, components=["gz"])
survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
nC = mesh.n_cells
sim = gravity.Simulation3DIntegral(mesh, survey=survey, rhoMap=maps.IdentityMap(nP=nC),
                                   sensitivity_dtype=np.float64)
t = time.perf_counter()
eng = sim._dev                        # tile plan (host) + upload
t_plan = time.perf_counter() - t
torch.cuda.synchronize()
t = time.perf_counter()
sim.G
torch.cuda.synchronize()
t_build = time.perf_counter() - t
w = rng.uniform(0.5, 1.5, n_rx)
m = rng.uniform(-0.5, 0.5, nC)
sim.getJtJdiag(m, W=sp.diags(w))      # warm-up
t = time.perf_counter()
diag = sim.getJtJdiag(m, W=sp.diags(w * 1.0001))
t_diag = time.perf_counter() - t
iters = 10
solver = DeviceGaussNewtonCG(sim, data_weights=w, reg_hessian=np.ones(nC), beta=1e-3 * float(diag.mean()),
                             cg_maxiter=iters, cg_rtol=0.0)
solver.set_model(m)
solver.jacobi_preconditioner()
g = rng.uniform(-1.0, 1.0, nC)
solver.find_search_direction(g)       # warm-up
t = time.perf_counter()
step = solver.find_search_direction(g)
t_cg = time.perf_counter() - t
g_bytes = n_rx * nC * 8
print(json.dumps({
    "workload": "cfg5 per-GPU shard: gravity gz, synthetic octree below topography",
    "cells": nC, "nodes": int(mesh.total_nodes.shape[0]), "receivers": n_rx, "G_bytes": g_bytes,
    "tiles": int(eng.host_plan.n_tiles), "nodes_per_cell": float(eng.host_plan.nodes_per_cell),
    "mesh_s": t_mesh, "plan_upload_s": t_plan,
    "build_s": t_build, "build_pairs_s": n_rx * nC / t_build,
    "getJtJdiag_s": t_diag, "getJtJdiag_gbs": g_bytes / t_diag / 1e9,
    "cg_iterations": int(solver.cg_count), "cg_s_per_iteration": t_cg / solver.cg_count,
    "cg_jvec_jtvec_gbs": 2 * g_bytes * solver.cg_count / t_cg / 1e9,
    "cg_rel_residual": float(solver.cg_rel_resid), "checksum": float(step.sum()),
}), flush=True)
