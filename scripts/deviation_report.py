"""Observed norm-wise deviation from the oracle / golden vectors (development aid)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import GOLDEN, load_golden, rowwise_rel, vec_rel
from oracle.oracle import OracleSimulation
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of

worst = {"G": 0.0, "fwd": 0.0, "gtv": 0.0}
for name in GOLDEN:
    c = load_golden(name)
    e = CudaEngine(c["physics"], c["nodes"], c["cell_nodes"], c["rlist"], b0=c["b0"], model_type=c["model_type"])
    e.build_G()
    worst["G"] = max(worst["G"], rowwise_rel(e.G_to_numpy(), c["G"]).max())
    worst["fwd"] = max(worst["fwd"], vec_rel(e.to_host(e.forward_dev(e.to_device(c["model"]))), c["fwd"]))
    worst["gtv"] = max(worst["gtv"], vec_rel(e.to_host(e.gt_dot_v_dev(e.to_device(c["v"]))), c["gtv"]))
print("golden (18 cases): ", {k: f"{v:.2e}" for k, v in worst.items()})
B0 = np.array([-8.55e3, -2.35e4, -4.33e4])
for label, shape, h, z, phys, comps, mt in [
        ("cfg1", (40, 40, 20), 5.0, 1.0, "gravity", ["gz"], "scalar"),
        ("cfg2", (100, 100, 50), 10.0, 5.0, "gravity", ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar"),
        ("cfg3", (200, 200, 100), 25.0, 30.0, "magnetics", ["tmi"], "scalar"),
        ("cfg4", (100, 100, 100), 10.0, 10.0, "magnetics", ["tmi"], "vector"),
        ("cfg4 high receivers (z=+500 m)", (100, 100, 100), 10.0, 500.0, "magnetics", ["tmi"], "vector")]:
    mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
    rng = np.random.default_rng(3)
    half = shape[0] * h / 2
    loc = np.c_[rng.uniform(-half, half, (6, 2)), np.full(6, z) + rng.uniform(0, 1, 6)]
    ora = OracleSimulation(phys, mesh.nodes, mesh.cell_nodes, [(loc, comps)], b0=B0, model_type=mt)
    eng = CudaEngine(phys, None, None, [(loc, comps)], b0=B0, model_type=mt, plan=build_tile_plan_tensor(tensor_axes_of(mesh)))
    G = ora.sensitivity_matrix(); eng.build_G()
    lo, hi = (-0.5, 0.5) if phys == "gravity" else (0.0, 0.05)
    m = rng.uniform(lo, hi, eng.nP)
    print(f"{label:32s} G {rowwise_rel(eng.G_to_numpy(), G).max():.2e}  fwd {vec_rel(eng.to_host(eng.forward_dev(eng.to_device(m))), ora.forward(m)):.2e}")
