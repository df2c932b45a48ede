#!/usr/bin/env python
"""
A/B harness for builds of libpf_cuda.so (development tool, GPU box).

    python scripts/exp_build_variants.py lib1.so lib2.so ...      (driver: one subprocess per library)
    PF_CUDA_LIB=lib.so python scripts/exp_build_variants.py --one  (one library, prints one JSON line)

Per library: device-timed G build of four programs (TMI vector 100^3 = the bench workload, TMI scalar
200x200x100, gravity gz 100^3, gravity 7 components 100x100x50), the matrix-free sinks on a receiver
subsample, and the deviation of a few rows from the CPU port (so that a fast but wrong variant
is caught on the spot).
"""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def one():
    import torch

    from oracle import oracle
    from simpeg_b200.engine import CudaEngine
    from simpeg_b200.meshes import TensorMesh
    from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
    from simpeg_b200.utils import dip_azimuth2cartesian

    def ms(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    b0 = 50000.0 * dip_azimuth2cartesian(60.0, 250.0).squeeze()
    cases = [
        ("tmi_vector_100^3", (100, 100, 100), 10.0, "magnetics", ["tmi"], "vector", 2500),
        ("tmi_scalar_200x200x100", (200, 200, 100), 25.0, "magnetics", ["tmi"], "scalar", 1250),
        ("grav_gz_100^3", (100, 100, 100), 10.0, "gravity", ["gz"], "scalar", 2500),
        ("grav_7comp_100x100x50", (100, 100, 50), 10.0, "gravity",
         ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar", 1250),
        ("tmideriv3_scalar_100^3", (100, 100, 100), 10.0, "magnetics", ["tmi_x", "tmi_y", "tmi_z"], "scalar", 1000),
        ("maggrad6_scalar_100x100x50", (100, 100, 50), 10.0, "magnetics",
         ["bxx", "bxy", "bxz", "byy", "byz", "bzz"], "scalar", 1000),
        ("bx_scalar_100^3", (100, 100, 100), 10.0, "magnetics", ["bx"], "scalar", 2000),
        ("gzz_100^3", (100, 100, 100), 10.0, "gravity", ["gzz"], "scalar", 2500),
    ]
    only = os.environ.get("PF_EXP_CASES")
    out = {"lib": os.environ.get("PF_CUDA_LIB", "default")}
    rng = np.random.default_rng(0)
    for name, shape, h, physics, comps, model_type, n_rx in cases:
        if only and name not in only.split(","):
            continue
        mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
        side = int(np.ceil(np.sqrt(n_rx)))
        xs = np.linspace(-shape[0] * h / 2, shape[0] * h / 2, side)
        X, Y = np.meshgrid(xs, xs)
        loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, h)][:n_rx]
        loc[:, 2] += rng.uniform(0, 1, n_rx)
        plan = build_tile_plan_tensor(tensor_axes_of(mesh), None, with_tiles=False)
        eng = CudaEngine(physics, None, None, [(loc, comps)], b0=b0, model_type=model_type,
                         sensitivity_dtype=np.float64, plan=plan)
        nC = mesh.n_cells
        t = ms(eng.build_G)
        res = {"build_ms": t, "build_gpairs": n_rx * nC / t / 1e6, "kernel": eng.describe("store")}
        # parity on 3 receivers
        pos = np.array([0, n_rx // 2, n_rx - 1])
        C = len(comps)
        rows = (pos[:, None] * C + np.arange(C)[None, :]).reshape(-1)
        Gs = oracle.OracleSimulation(physics, mesh.nodes, mesh.cell_nodes, [(loc[pos], comps)], b0=b0,
                                     model_type=model_type, backend="c").sensitivity_matrix()
        Gd = eng.G[torch.from_numpy(rows).to(eng.device), : eng.nP].cpu().numpy()
        res["parity"] = float(np.max(np.max(np.abs(Gd - Gs), axis=1) / np.max(np.abs(Gs), axis=1)))
        eng.G = None
        torch.cuda.empty_cache()
        # sinks on 256 receivers
        sub = CudaEngine(physics, None, None, [(loc[:256], comps)], b0=b0, model_type=model_type, plan=plan)
        xd = sub.to_device(rng.uniform(-1, 1, sub.nP))
        wd = sub.to_device(rng.uniform(0, 1, sub.n_rows))
        for key, fn in (("forward", lambda: sub.forward_dev(xd)), ("gtv", lambda: sub.gt_dot_v_dev(wd)),
                        ("diag", lambda: sub.gtg_diag_dev(wd))):
            res[key + "_gpairs"] = 256 * nC / ms(fn, 3) / 1e6
        res["diag_kernel"] = sub.describe("diag")
        # sink parity against the port on the same 3 receivers
        sub3 = CudaEngine(physics, None, None, [(loc[pos], comps)], b0=b0, model_type=model_type, plan=plan)
        x = rng.uniform(-1, 1, sub3.nP)
        w = rng.uniform(0.5, 1.5, sub3.n_rows)

        def rel(a, b):
            return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

        res["parity_sinks"] = max(
            rel(sub3.forward_dev(sub3.to_device(x)).cpu().numpy(), Gs @ x),
            rel(sub3.gt_dot_v_dev(sub3.to_device(w)).cpu().numpy(), Gs.T @ w),
            rel(sub3.gtg_diag_dev(sub3.to_device(w)).cpu().numpy(), np.einsum("i,ij,ij->j", w, Gs, Gs)))
        out[name] = res
        del eng, sub, sub3, xd, wd
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


def main():
    if "--one" in sys.argv:
        one()
        return
    libs = [a for a in sys.argv[1:] if not a.startswith("--")]
    for lib in libs:
        env = dict(os.environ, PF_CUDA_LIB=str(Path(lib).resolve()))
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            print(json.dumps({"lib": lib, "error": r.stderr[-1500:]}), flush=True)
            continue
        d = json.loads(line[-1])
        print(json.dumps(d), flush=True)
        short = {k: (round(v["build_gpairs"], 2), round(v["forward_gpairs"], 1), round(v["gtv_gpairs"], 1),
                     round(v["diag_gpairs"], 1), "%.1e" % v["parity"], "%.1e" % v["parity_sinks"])
                 for k, v in d.items() if isinstance(v, dict)}
        print("SUMMARY", Path(lib).name, short, flush=True)


if __name__ == "__main__":
    main()
