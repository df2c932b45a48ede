"""Engine-level timing probe (development aid; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine, probe_fp64_peak
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
MARCH = '--no-march' not in sys.argv

def timed(fn, n=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def run(name, shape, h, n_rx, physics, comps, model_type, dtype=np.float64, z=10.0):
    mesh = TensorMesh([np.full(n, h) for n in shape], "CCN")
    side = int(np.ceil(np.sqrt(n_rx)))
    xs = np.linspace(-shape[0]*h/2, shape[0]*h/2, side); ys = np.linspace(-shape[1]*h/2, shape[1]*h/2, side)
    X, Y = np.meshgrid(xs, ys); rx = np.c_[X.ravel(), Y.ravel(), np.full(X.size, z)][:n_rx]
    t = time.time()
    plan = build_tile_plan_tensor(tensor_axes_of(mesh))
    eng = CudaEngine(physics, None, None, [(rx, comps)], b0=np.array([1e4,-2e4,-4e4]), model_type=model_type, sensitivity_dtype=dtype, plan=plan, use_marching=MARCH)
    tplan = time.time() - t
    pairs = n_rx * eng.nC
    tb = timed(eng.build_G)
    gbytes = eng.n_rows * eng.nP * (8 if dtype is np.float64 else 4)
    x = torch.rand(eng.nP, dtype=torch.float64, device="cuda"); y = torch.rand(eng.n_rows, dtype=torch.float64, device="cuda")
    tg = timed(lambda: eng.gemv_dev(x), 5); tt = timed(lambda: eng.gemv_t_dev(y), 5); tc = timed(lambda: eng.colsumsq_dev(y), 5)
    tf = timed(lambda: eng.forward_dev(x), 2); tv = timed(lambda: eng.gt_dot_v_dev(y), 2); td = timed(lambda: eng.gtg_diag_dev(y), 2)
    print(f"{name}: plan {tplan:.2f}s tiles {eng.host_plan.n_tiles} nu {eng.host_plan.nodes_per_cell:.3f} | build {tb:.2f} ms {pairs/tb/1e6:.1f} Gpairs/s | "
          f"G {gbytes/1e9:.2f} GB gemv {tg:.3f} ms {gbytes/tg/1e6:.0f} GB/s gemv_t {tt:.3f} ms {gbytes/tt/1e6:.0f} GB/s colsumsq {tc:.3f} ms {gbytes/tc/1e6:.0f} GB/s | "
          f"fwd {tf:.2f} ms {pairs/tf/1e6:.1f} Gp/s gtv {tv:.2f} ms {pairs/tv/1e6:.1f} Gp/s diag {td:.2f} ms {pairs/td/1e6:.1f} Gp/s", flush=True)
    del eng; torch.cuda.empty_cache()

print("fp64 peak TFLOP/s", probe_fp64_peak()/1e12)
run("cfg1 gz 40x40x20 441rx", (40,40,20), 5.0, 441, "gravity", ["gz"], "scalar", z=1.0)
run("cfg2-like grav7 100x100x50 512rx", (100,100,50), 10.0, 512, "gravity", ["gxx","gxy","gxz","gyy","gyz","gzz","gz"], "scalar", z=5.0)
run("cfg4-like tmi vector 100^3 1024rx", (100,100,100), 10.0, 1024, "magnetics", ["tmi"], "vector")
run("cfg3-like tmi scalar 200x200x100 256rx", (200,200,100), 25.0, 256, "magnetics", ["tmi"], "scalar", z=30.0)
run("tmi_deriv scalar 100^3 512rx", (100,100,100), 10.0, 512, "magnetics", ["tmi_x","tmi_y","tmi_z"], "scalar")
run("cfg4-like f32 1024rx", (100,100,100), 10.0, 1024, "magnetics", ["tmi"], "vector", np.float32)
