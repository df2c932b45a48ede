"""cfg5-style probe: gravity gz on a synthetic octree below Gaussian-hill topography."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import synthetic_octree
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
topo = lambda x, y: 50.0 * np.exp(-0.5 * ((x / 100.0) ** 2 + (y / 100.0) ** 2)) - 50.0
t = time.time(); mesh = synthetic_octree((n, n, 12), 10.0, levels=3, topo=topo); t_mesh = time.time() - t
rng = np.random.default_rng(0)
half = n * 40.0 / 2
xy = rng.uniform(-half, half, (256, 2)); loc = np.c_[xy, topo(xy[:, 0], xy[:, 1]) + 5.0]
t = time.time(); eng = CudaEngine("gravity", mesh.total_nodes, mesh.cell_nodes, [(loc, ["gz"])]); t_plan = time.time() - t
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
pairs = 256 * eng.nC
x = torch.rand(eng.nP, dtype=torch.float64, device="cuda"); y = torch.rand(eng.n_rows, dtype=torch.float64, device="cuda")
tb = timed(eng.build_G); tf = timed(lambda: eng.forward_dev(x)); tv = timed(lambda: eng.gt_dot_v_dev(y)); td = timed(lambda: eng.gtg_diag_dev(y))
gb = eng.n_rows * eng.nP * 8
tg = timed(lambda: eng.gemv_dev(x)); tt = timed(lambda: eng.gemv_t_dev(y))
hp = eng.host_plan
print(f"octree cells {eng.nC} nodes {mesh.total_nodes.shape[0]} (mesh {t_mesh:.1f}s, plan+upload {t_plan:.1f}s) tiles {hp.n_tiles} "
      f"cells/tile {eng.nC/hp.n_tiles:.0f} nu {hp.nodes_per_cell:.3f} | build {pairs/tb/1e6:.1f} fwd {pairs/tf/1e6:.1f} "
      f"gtv {pairs/tv/1e6:.1f} diag {pairs/td/1e6:.1f} Gpairs/s | gemv {gb/tg/1e6:.0f} gemv_t {gb/tt/1e6:.0f} GB/s")
