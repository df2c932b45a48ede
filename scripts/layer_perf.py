"""Device throughput of the equivalent-source layer simulations (pairs/s): flat layer on a tensor
mesh (z-marching kernel, ~2 nodes per pair) and a draped layer (tile kernel, 8 nodes per pair),
gravity gz and magnetic TMI, stored G build and forward_only dpred."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from simpeg_b200 import maps  # noqa: E402
from simpeg_b200.meshes import TensorMesh  # noqa: E402
from simpeg_b200.potential_fields import gravity, magnetics  # noqa: E402


def run(physics, flat, n_side=400, n_rx=2048):
    mesh = TensorMesh([np.full(n_side, 25.0), np.full(n_side, 25.0)], "CC")
    nC = mesh.n_cells
    rng = np.random.default_rng(0)
    if flat:
        top, bottom = -10.0, -60.0
    else:
        cc = 0.5 * (mesh.cell_bounds[:, [0, 2]] + mesh.cell_bounds[:, [1, 3]])
        top = 50.0 * np.exp(-0.5 * ((cc[:, 0] / 1000.0) ** 2 + (cc[:, 1] / 1000.0) ** 2)) - 30.0
        bottom = top - 50.0
    half = n_side * 12.5
    loc = np.c_[rng.uniform(-half, half, n_rx), rng.uniform(-half, half, n_rx), np.full(n_rx, 60.0)]
    if physics == "gravity":
        rx = gravity.receivers.Point(loc, components=["gz"])
        survey = gravity.survey.Survey(gravity.sources.SourceField(receiver_list=[rx]))
        mk = lambda store: gravity.SimulationEquivalentSourceLayer(  # noqa: E731
            mesh, top, bottom, survey=survey, rhoMap=maps.IdentityMap(nP=nC),
            store_sensitivities=store, sensitivity_dtype=np.float64)
    else:
        rx = magnetics.receivers.Point(loc, components=["tmi"])
        src = magnetics.sources.UniformBackgroundField([rx], 50000.0, 60.0, 250.0)
        survey = magnetics.survey.Survey(src)
        mk = lambda store: magnetics.SimulationEquivalentSourceLayer(  # noqa: E731
            mesh, top, bottom, survey=survey, chiMap=maps.IdentityMap(nP=nC),
            store_sensitivities=store, sensitivity_dtype=np.float64)
    out = {"physics": physics, "layer": "flat" if flat else "draped", "cells": nC, "receivers": n_rx}
    sim = mk("ram")
    t0 = time.perf_counter()
    eng = sim._dev
    out["plan_s"] = time.perf_counter() - t0
    out["nodes_per_cell"] = None if eng.host_plan.grid_shape is not None else eng.host_plan.nodes_per_cell

    def dev_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    pairs = n_rx * nC
    out["build_G_pairs_s"] = pairs / dev_ms(eng.build_G) * 1e3
    m = eng.to_device(rng.uniform(size=eng.nP))
    out["forward_only_pairs_s"] = pairs / dev_ms(lambda: eng.forward_dev(m)) * 1e3
    v = eng.to_device(rng.uniform(size=eng.n_rows))
    out["gt_dot_v_pairs_s"] = pairs / dev_ms(lambda: eng.gt_dot_v_dev(v)) * 1e3
    return out


if __name__ == "__main__":
    for physics in ("gravity", "magnetics"):
        for flat in (True, False):
            print(json.dumps(run(physics, flat)), flush=True)
