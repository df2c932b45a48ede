"""ncu driver for one named case: python scripts/prof_case.py grav7|tmideriv|gz|tmi [n_rx]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
case = sys.argv[1]; n_rx = int(sys.argv[2]) if len(sys.argv) > 2 else 296
cfg = {"grav7": ((100, 100, 50), 10.0, "gravity", ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar"),
       "gz": ((100, 100, 50), 10.0, "gravity", ["gz"], "scalar"),
       "tmideriv": ((100, 100, 100), 10.0, "magnetics", ["tmi_x", "tmi_y", "tmi_z"], "scalar"),
       "tmi": ((100, 100, 100), 10.0, "magnetics", ["tmi"], "vector")}[case]
shape, h, phys, comps, mt = cfg
mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
side = int(np.ceil(np.sqrt(n_rx))); xs = np.linspace(-450, 450, side)
X, Y = np.meshgrid(xs, xs); rx = np.c_[X.ravel() + 0.3, Y.ravel() - 0.2, np.full(X.size, 7.0)][:n_rx]
eng = CudaEngine(phys, None, None, [(rx, comps)], b0=np.array([1e4, -2e4, -4e4]), model_type=mt,
                 plan=build_tile_plan_tensor(tensor_axes_of(mesh)))
eng.build_G(); torch.cuda.synchronize()
x = torch.rand(eng.nP, dtype=torch.float64, device="cuda"); y = torch.rand(eng.n_rows, dtype=torch.float64, device="cuda")
eng.forward_dev(x); eng.gt_dot_v_dev(y); torch.cuda.synchronize()
