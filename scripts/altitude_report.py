#!/usr/bin/env python
"""Deviation of the CUDA engine, and of the reference (through the bit-identical C port), from the
extended-precision evaluation at airborne heights (GPU box; prints one JSON line per case)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    from helpers import rowwise_rel
    from oracle import longdouble_oracle as ld
    from oracle.make_golden_altitude import B0, CASES, geometry, truth
    from oracle.oracle import OracleSimulation
    from simpeg_b200.engine import CudaEngine
    from simpeg_b200.tiling import build_tile_plan_tensor

    nodes, cell_nodes, rx = geometry()
    axes = (-80.0 + 10.0 * np.arange(17), -80.0 + 10.0 * np.arange(17), -80.0 + 10.0 * np.arange(9))
    for name, (physics, comps, model_type) in CASES.items():
        exact = truth(name, nodes, cell_nodes, rx)
        G_port = OracleSimulation(physics, nodes, cell_nodes, [(rx, comps)], b0=np.array(B0),
                                  model_type=model_type, backend="c").sensitivity_matrix()
        ref = ld.rowwise_deviation(G_port, exact)
        out = {"case": name, "heights": [200, 500, 1000],
               "reference_vs_exact": [float(ref[3 * i:3 * i + 3].max()) for i in range(3)]}
        for use_marching in (True, False):
            eng = CudaEngine(physics, None, None, [(rx, comps)], b0=np.array(B0), model_type=model_type,
                             sensitivity_dtype=np.float64, plan=build_tile_plan_tensor(axes, None),
                             use_marching=use_marching)
            eng.build_G()
            G = eng.G_to_numpy()
            de, dr = ld.rowwise_deviation(G, exact), rowwise_rel(G, G_port)
            key = "march" if use_marching else "tile"
            out[key + "_vs_exact"] = [float(de[3 * i:3 * i + 3].max()) for i in range(3)]
            out[key + "_vs_reference"] = [float(dr[3 * i:3 * i + 3].max()) for i in range(3)]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
