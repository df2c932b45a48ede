import sys, numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
mesh = TensorMesh([np.full(100, 10.0), np.full(100, 10.0), np.full(50, 10.0)], "CCN")
plan = build_tile_plan_tensor(tensor_axes_of(mesh))
n_rx = 256; side = 16; xs = np.linspace(-450, 450, side); X, Y = np.meshgrid(xs, xs)
rx = np.c_[X.ravel() + 0.3, Y.ravel() - 0.2, np.full(X.size, 7.0)]
def timed(fn, n=3):
    torch.cuda.synchronize(); best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
allc = ["gz", "gzz", "gxx", "gyy", "gxy", "gxz", "gyz", "gx", "gy", "guv"]
for march in (True, False):
    for nc in (1, 2, 3, 4, 5, 7, 10):
        eng = CudaEngine("gravity", None, None, [(rx, allc[:nc])], plan=plan, use_marching=march)
        t = timed(eng.build_G)
        print(f"march={march} ncomp={nc:2d}: {t:7.2f} ms  {n_rx*eng.nC/t/1e6:6.1f} Gpairs/s  G write {eng.n_rows*eng.nP*8/t/1e6:7.0f} GB/s", flush=True)
        del eng; torch.cuda.empty_cache()
