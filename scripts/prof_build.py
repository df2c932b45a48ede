"""Small driver for ncu captures of the build / product kernels."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
n_rx = int(sys.argv[1]) if len(sys.argv) > 1 else 296
mesh = TensorMesh([np.full(100, 10.0)] * 3, "CCN")
side = int(np.ceil(np.sqrt(n_rx)))
xs = np.linspace(-500, 500, side)
X, Y = np.meshgrid(xs, xs); rx = np.c_[X.ravel(), Y.ravel(), np.full(X.size, 10.0)][:n_rx]
eng = CudaEngine("magnetics", None, None, [(rx, ["tmi"])], b0=np.array([1e4, -2e4, -4e4]), model_type="vector",
                 plan=build_tile_plan_tensor(tensor_axes_of(mesh)), use_marching="--no-march" not in sys.argv)
eng.build_G(); torch.cuda.synchronize()
x = torch.rand(eng.nP, dtype=torch.float64, device="cuda"); y = torch.rand(eng.n_rows, dtype=torch.float64, device="cuda")
eng.gemv_dev(x); eng.gemv_t_dev(y); eng.colsumsq_dev(y); eng.forward_dev(x); eng.gt_dot_v_dev(y)
torch.cuda.synchronize()
