#!/usr/bin/env python
"""One short launch of a build / sink kernel for `ncu` (development tool, GPU box).

    python scripts/prof_one.py <case> <n_rx> [sink]

case: tmi_vector | tmi_scalar | grav_gz | grav7 | octree_gz ; sink: store (default) | forward | gtv | diag | products
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from simpeg_b200.engine import CudaEngine
    from simpeg_b200.meshes import TensorMesh, synthetic_octree
    from simpeg_b200.tiling import build_tile_plan, build_tile_plan_tensor, tensor_axes_of
    from simpeg_b200.utils import dip_azimuth2cartesian

    case = sys.argv[1]
    n_rx = int(sys.argv[2])
    sink = sys.argv[3] if len(sys.argv) > 3 else "store"
    b0 = 50000.0 * dip_azimuth2cartesian(60.0, 250.0).squeeze()
    spec = {
        "tmi_vector": ((100, 100, 100), 10.0, "magnetics", ["tmi"], "vector"),
        "tmi_scalar": ((200, 200, 100), 25.0, "magnetics", ["tmi"], "scalar"),
        "grav_gz": ((100, 100, 100), 10.0, "gravity", ["gz"], "scalar"),
        "grav7": ((100, 100, 50), 10.0, "gravity", ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar"),
    }
    rng = np.random.default_rng(0)
    if case == "octree_gz":
        mesh = synthetic_octree((60, 60, 12), 10.0, levels=3)
        physics, comps, model_type = "gravity", ["gz"], "scalar"
        half = 60 * 40.0 / 2
        loc = np.c_[rng.uniform(-half, half, (n_rx, 2)), np.full(n_rx, 5.0)]
        plan = build_tile_plan(mesh.total_nodes, mesh.cell_nodes, "rcb")
    else:
        shape, h, physics, comps, model_type = spec[case]
        mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
        side = int(np.ceil(np.sqrt(n_rx)))
        xs = np.linspace(-shape[0] * h / 2, shape[0] * h / 2, side)
        X, Y = np.meshgrid(xs, xs)
        loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, h)][:n_rx]
        loc[:, 2] += rng.uniform(0, 1, n_rx)
        plan = build_tile_plan_tensor(tensor_axes_of(mesh), None, with_tiles=False)
    eng = CudaEngine(physics, None, None, [(loc, comps)], b0=b0, model_type=model_type,
                     sensitivity_dtype=np.float64, plan=plan)
    x = eng.to_device(rng.uniform(-1, 1, eng.nP))
    w = eng.to_device(rng.uniform(0, 1, eng.n_rows))
    fn = {"store": eng.build_G, "forward": lambda: eng.forward_dev(x), "gtv": lambda: eng.gt_dot_v_dev(w),
          "diag": lambda: eng.gtg_diag_dev(w),
          "products": lambda: (eng.gemv_dev(x), eng.gemv_t_dev(w), eng.colsumsq_dev(w))}[sink]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(case, sink, eng.describe("store" if sink == "products" else sink))


if __name__ == "__main__":
    main()
