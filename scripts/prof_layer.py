"""ncu driver: draped equivalent-source layer (tile kernel, 8 nodes per cell), gravity gz."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from simpeg_b200 import maps
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.potential_fields import gravity
mesh = TensorMesh([np.full(300, 25.0), np.full(300, 25.0)], "CC")
cc = 0.5 * (mesh.cell_bounds[:, [0, 2]] + mesh.cell_bounds[:, [1, 3]])
top = 50.0 * np.exp(-0.5 * ((cc[:, 0] / 1000.0) ** 2 + (cc[:, 1] / 1000.0) ** 2)) - 30.0
rng = np.random.default_rng(0)
loc = np.c_[rng.uniform(-3000, 3000, 1024), rng.uniform(-3000, 3000, 1024), np.full(1024, 60.0)]
rx = gravity.receivers.Point(loc, components=["gz"])
survey = gravity.survey.Survey(gravity.sources.SourceField(receiver_list=[rx]))
sim = gravity.SimulationEquivalentSourceLayer(mesh, top, top - 50.0, survey=survey,
                                              rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                              sensitivity_dtype=np.float64)
sim.G
torch.cuda.synchronize()
