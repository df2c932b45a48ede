"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): marching, tile, node-space
kernels, stored-G products (fp64 + fp32), layer simulation, device CG."""
import sys

import numpy as np

sys.path.insert(0, ".")
from simpeg_b200 import maps  # noqa: E402
from simpeg_b200.gauss_newton import DeviceGaussNewtonCG  # noqa: E402
from simpeg_b200.meshes import TensorMesh, synthetic_octree  # noqa: E402
from simpeg_b200.potential_fields import gravity, magnetics  # noqa: E402

rng = np.random.default_rng(0)
mesh = TensorMesh([[(5.0, 9)], [(5.0, 7)], [(5.0, 5)]], "CCN")
loc = np.c_[rng.uniform(-20, 20, 11), rng.uniform(-15, 15, 11), rng.uniform(1, 4, 11)]
active = rng.uniform(size=mesh.n_cells) < 0.8
nA = int(active.sum())
for store in ("ram", "forward_only"):
    for dtype in (np.float64, np.float32):
        rx = magnetics.receivers.Point(loc, components=["tmi", "bx", "tmi_z"])
        src = magnetics.sources.UniformBackgroundField([rx], 50000.0, 60.0, 250.0)
        sim = magnetics.Simulation3DIntegral(mesh, survey=magnetics.survey.Survey(src), active_cells=active,
                                             chiMap=maps.IdentityMap(nP=3 * nA), model_type="vector",
                                             store_sensitivities=store, sensitivity_dtype=dtype)
        m = rng.uniform(0, 0.05, 3 * nA)
        d = sim.dpred(m)
        sim.Jtvec(m, d)
        sim.getJtJdiag(m)
rxg = gravity.receivers.Point(loc, components=["gz", "gxx", "gzz"])
sg = gravity.survey.Survey(gravity.sources.SourceField([rxg]))
simg = gravity.Simulation3DIntegral(mesh, survey=sg, rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                    sensitivity_dtype=np.float64)
rho = rng.uniform(-0.5, 0.5, mesh.n_cells)
simg.dpred(rho)
solver = DeviceGaussNewtonCG(simg, reg_hessian=np.ones(mesh.n_cells), beta=1e-3, lower=-0.4, upper=0.4)
solver.set_model(np.clip(rho, -0.4, 0.4))
solver.jacobi_preconditioner()
solver.find_search_direction(rng.uniform(-1, 1, mesh.n_cells))
tree = synthetic_octree((6, 6, 4), 10.0, levels=2, topo=lambda x, y: -5.0 + 0 * x)
rxt = gravity.receivers.Point(loc * 2, components=["gz"])
simt = gravity.Simulation3DIntegral(tree, survey=gravity.survey.Survey(gravity.sources.SourceField([rxt])),
                                    rhoMap=maps.IdentityMap(nP=tree.n_cells), sensitivity_dtype=np.float64)
simt.dpred(rng.uniform(size=tree.n_cells))
mesh2 = TensorMesh([[(5.0, 9)], [(5.0, 7)]], "CC")
top = -1.0 - rng.uniform(size=mesh2.n_cells)
lay = gravity.SimulationEquivalentSourceLayer(mesh2, top, top - 3.0, survey=sg,
                                              rhoMap=maps.IdentityMap(nP=mesh2.n_cells),
                                              sensitivity_dtype=np.float64)
lay.dpred(rng.uniform(size=mesh2.n_cells))
# round 2: derived gzz (receivers on the mesh surface take the exact S-sum path), blocked diag,
# device CG with a sparse model-map derivative, tile containers
import scipy.sparse as sp  # noqa: E402

loc0 = loc.copy()
loc0[:4, 2] = 0.0                                   # on the top face of the mesh ("CCN": top at z = 0)
loc0[0, :2] = mesh.nodes_x[3], mesh.nodes_y[2]      # exactly on a node
rxd = gravity.receivers.Point(loc0, components=["gxx", "gyy", "gzz", "gz"])
sd = gravity.survey.Survey(gravity.sources.SourceField([rxd]))
for store in ("ram", "forward_only"):
    simd = gravity.Simulation3DIntegral(mesh, survey=sd, rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                        store_sensitivities=store, sensitivity_dtype=np.float64)
    simd.dpred(rho)
    simd.getJtJdiag(rho)
P = sp.random(mesh.n_cells, 40, density=0.1, random_state=1, format="csr")
simp = gravity.Simulation3DIntegral(mesh, survey=sg, rhoMap=maps.LinearMap(P), sensitivity_dtype=np.float64)
cgp = DeviceGaussNewtonCG(simp, reg_hessian=np.ones(40), beta=1e-3)
cgp.set_model(rng.uniform(size=40))
cgp.find_search_direction(rng.uniform(-1, 1, 40))
print("sanitize pass done")
