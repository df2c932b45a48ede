#!/usr/bin/env python
"""
Calibration of the CPU baseline: the reference's OWN numba loops against the C/OpenMP port
(``oracle/pf_oracle.c``) that ``bench.py`` times on the GPU box (``cpu_baseline.kind = "port"``).

The reference cannot be imported on the GPU box (``/root/reference`` does not travel, choclo /
discretize are absent), so ``bench.py --impl reference`` times the port.  This script measures,
in the build container where the reference tree exists, how far the port is from the code it
stands in for, on the same host, the same threads, the same inputs:

    reference: simpeg/potential_fields/gravity/_numba/_3d_mesh.py   _sensitivity_gravity (97-163)
               simpeg/potential_fields/magnetics/_numba/_3d_mesh.py _sensitivity_tmi     (208-372)
               (loaded unmodified by path on the choclo shim, oracle/ref_loader.py)
    port:      pfo_* loops of oracle/pf_oracle.c through oracle.OracleSimulation(backend="c")

Shape: BASELINE config 1 mesh (40 x 40 x 20 cells, h = 5 m) with 21 x 21 receivers; cases: gravity gz
(scalar) and magnetic TMI with a vector model (the bench workload's kernel family).

    python scripts/calibrate_port.py [--out profiles/r2_port_calibration.json]

The record holds pairs/s of both, their ratio, and the bit-level agreement of the two matrices.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def best_of(fn, reps):
    best = np.inf
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_port_calibration.json"))
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()

    from oracle import oracle, ref_loader
    from simpeg_b200.meshes import TensorMesh
    from simpeg_b200.utils import dip_azimuth2cartesian

    if not ref_loader.reference_available():
        raise SystemExit("the reference tree is not available here (build container only)")
    import numba

    threads = min(oracle.num_threads(), numba.get_num_threads())
    oracle.set_num_threads(threads)
    numba.set_num_threads(threads)

    mesh = TensorMesh([np.full(40, 5.0), np.full(40, 5.0), np.full(20, 5.0)], "CCN")
    x = np.linspace(-80.0, 80.0, 21)
    X, Y = np.meshgrid(x, x)
    rng = np.random.default_rng(0)
    loc = np.c_[X.ravel(), Y.ravel(), 1.0 + rng.uniform(0, 1, X.size)]
    b0 = 50000.0 * dip_azimuth2cartesian(60.0, 250.0).squeeze()
    pairs = loc.shape[0] * mesh.n_cells

    cases = [("gravity gz (scalar)", "gravity", ["gz"], "scalar"),
             ("magnetics tmi (vector model)", "magnetics", ["tmi"], "vector")]
    record = {"host_threads": threads, "cpu_count": os.cpu_count(), "mesh": "40x40x20 h=5 m",
              "receivers": int(loc.shape[0]), "pairs": int(pairs), "reps": args.reps, "cases": []}
    for label, physics, comps, model_type in cases:
        sims = {}
        for backend in ("ref", "c"):
            sims[backend] = oracle.OracleSimulation(physics, mesh.nodes, mesh.cell_nodes,
                                                    [(loc, comps)], b0=b0, model_type=model_type,
                                                    backend=backend, parallel=True)
        G_ref = sims["ref"].sensitivity_matrix()       # includes the numba JIT: not timed
        G_c = sims["c"].sensitivity_matrix()
        t_ref = best_of(sims["ref"].sensitivity_matrix, args.reps)
        t_c = best_of(sims["c"].sensitivity_matrix, args.reps)
        record["cases"].append({
            "case": label,
            "reference_numba_pairs_s": pairs / t_ref,
            "port_pairs_s": pairs / t_c,
            "port_over_reference": t_ref / t_c,
            "bit_identical": bool(np.array_equal(G_ref, G_c)),
            "max_abs_diff": float(np.max(np.abs(G_ref - G_c))),
        })
        print(json.dumps(record["cases"][-1]))
    record["reading"] = ("port_over_reference ~ 1: the port is a fair stand-in for the reference's "
                         "numba-parallel path on the same cores (cpu_baseline.kind = 'port')")
    Path(args.out).write_text(json.dumps(record, indent=1) + "\n")


if __name__ == "__main__":
    main()
