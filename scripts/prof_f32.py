import sys, numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
mesh = TensorMesh([np.full(100, 10.0)] * 3, "CCN")
rx = np.c_[np.random.default_rng(0).uniform(-400, 400, (592, 2)), np.full(592, 10.0)]
eng = CudaEngine("magnetics", None, None, [(rx, ["tmi"])], b0=np.array([1e4, -2e4, -4e4]), model_type="vector",
                 sensitivity_dtype=np.float32, plan=build_tile_plan_tensor(tensor_axes_of(mesh)))
eng.build_G(); torch.cuda.synchronize()
x = torch.rand(eng.nP, dtype=torch.float64, device="cuda"); y = torch.rand(eng.n_rows, dtype=torch.float64, device="cuda")
eng.gemv_dev(x); eng.gemv_t_dev(y); torch.cuda.synchronize()
