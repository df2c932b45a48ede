"""
BASELINE config 3 at its per-GPU size: magnetic TMI, induced scalar susceptibility,
store_sensitivities='forward_only' dpred on a 200 x 200 x 100 tensor mesh (4 M cells), 5000 receivers
(= 40 000 / 8) at z = +30 m; plus the matrix-free Jtvec and getJtJdiag of the same simulation.
Through the public API with numpy vectors; prints one JSON line.
"""
import json
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, ".")
from simpeg_b200 import maps  # noqa: E402
from simpeg_b200.meshes import TensorMesh  # noqa: E402
from simpeg_b200.potential_fields import magnetics  # noqa: E402

n_rx = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
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


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t, out


t = time.perf_counter()
sim = magnetics.Simulation3DIntegral(mesh, survey=survey, chiMap=maps.IdentityMap(nP=nC),
                                     store_sensitivities="forward_only")
sim._dev
t_setup = time.perf_counter() - t
t_fwd, d = timed(lambda: sim.dpred(chi))
v = rng.uniform(0.0, 1.0, survey.nD)
t_jt, jt = timed(lambda: sim.Jtvec(chi, v))
w = rng.uniform(0.5, 1.5, survey.nD)
k = [0]


def diag():
    k[0] += 1
    return sim.getJtJdiag(chi, W=sp.diags(w * (1.0 + 1e-6 * k[0])))      # defeat the sha256 cache


t_dg, dg = timed(diag)
pairs = n_rx * nC
print(json.dumps({
    "workload": "cfg3 per-GPU shard: magnetics tmi scalar, forward_only, TensorMesh 200x200x100",
    "cells": nC, "receivers": n_rx, "setup_s": t_setup,
    "dpred_s": t_fwd, "dpred_pairs_s": pairs / t_fwd,
    "jtvec_s": t_jt, "jtvec_pairs_s": pairs / t_jt,
    "getJtJdiag_s": t_dg, "getJtJdiag_pairs_s": pairs / t_dg,
    "checksum": float(d.sum() + jt.sum() + dg.sum()),
}), flush=True)
