"""
GPU parity of ``SimulationEquivalentSourceLayer(engine="cuda")`` against
golden vectors generated from the reference's own equivalent-source loops
(gravity/_numba/_2d_mesh.py, magnetics/_numba/_2d_mesh.py; oracle/make_golden_layer.py), through
the public API: stored G, forward_only dpred, on-the-fly G.T v and diag(G.T W G).
Bar: 1e-9 norm-wise per row of G / per vector (fp64), 1e-6 for float32 storage.
"""
import numpy as np
import pytest

from helpers import RTOL_NORMWISE, rowwise_rel, single_component, vec_rel
from layer_helpers import LAYER_GOLDEN, layer_simulation, load_layer_golden
from simpeg_b200.meshes import TensorMesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", LAYER_GOLDEN)
def test_layer_stored_G_and_products(name):
    case = load_layer_golden(name)
    sim = layer_simulation(case, store="ram")
    G = sim.G.to_numpy()
    assert G.shape == case["G"].shape
    assert rowwise_rel(G, case["G"]).max() <= RTOL_NORMWISE
    assert vec_rel(sim.dpred(case["model"]), case["fwd"]) <= RTOL_NORMWISE
    assert vec_rel(sim.Jvec(case["model"], case["model"]), case["fwd"]) <= RTOL_NORMWISE
    assert vec_rel(sim.Jtvec(case["model"], case["v"]), case["gtv"]) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", LAYER_GOLDEN)
def test_layer_forward_only(name):
    case = load_layer_golden(name)
    sim = layer_simulation(case, store="forward_only")
    assert vec_rel(sim.dpred(case["model"]), case["fwd"]) <= RTOL_NORMWISE
    assert vec_rel(sim.Jtvec(case["model"], case["v"]), case["gtv"]) <= RTOL_NORMWISE
    diag = sim._gtg_diagonal_without_building_g(case["w"])
    expected = np.einsum("i,ij,ij->j", case["w"], case["G"], case["G"])
    assert vec_rel(diag, expected) <= RTOL_NORMWISE
    if single_component(case):      # the reference's weights[i] quirk is invisible here
        assert vec_rel(diag, case["diag_ref"]) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", ["layer_grav_flat_all10", "layer_mag_tmi_flat_vector"])
def test_flat_layer_on_a_tensor_mesh_takes_the_marching_kernel(name):
    """Same numbers when the 2D mesh is a tensor mesh and the layer is flat (one-cell-thick 3D
    tensor mesh -> z-marching kernel) as through the generic tile kernel."""
    case = load_layer_golden(name)
    xs = np.unique(case["bounds"][:, :2])
    ys = np.unique(case["bounds"][:, 2:])
    mesh = TensorMesh([np.diff(xs), np.diff(ys)], origin=[xs[0], ys[0]])
    assert np.array_equal(mesh.cell_bounds, case["bounds"])
    sim = layer_simulation(case, mesh=mesh)
    assert sim._tile_plan().grid_shape == (len(xs) - 1, len(ys) - 1, 1)
    assert rowwise_rel(sim.G.to_numpy(), case["G"]).max() <= RTOL_NORMWISE
    fo = layer_simulation(case, store="forward_only", mesh=mesh)
    assert vec_rel(fo.dpred(case["model"]), case["fwd"]) <= RTOL_NORMWISE


def test_layer_float32_storage():
    case = load_layer_golden("layer_grav_topo_gz")
    sim = layer_simulation(case, dtype=np.float32)
    G = sim.G.to_numpy()
    assert G.dtype == np.float32
    assert rowwise_rel(G, case["G"]).max() <= 1e-6
    assert vec_rel(sim.dpred(case["model"]), case["fwd"]) <= 1e-6


def test_layer_getJtJdiag_and_inactive_cells():
    import scipy.sparse as sp

    case = load_layer_golden("layer_grav_inactive_gz")
    sim = layer_simulation(case)
    W = sp.diags(np.sqrt(case["w"]))
    expected = np.einsum("i,ij,ij->j", case["w"], case["G"], case["G"])
    assert vec_rel(sim.getJtJdiag(case["model"], W=W), expected) <= RTOL_NORMWISE


def test_large_draped_layer_against_oracle_subsample():
    """A layer at a realistic size (120 x 100 cells draped under topography, 400 receivers):
    random rows against the C oracle of the reference loops."""
    from oracle.oracle import OracleLayerSimulation
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity

    mesh = TensorMesh([np.full(120, 25.0), np.full(100, 25.0)], "CC")
    cc = 0.5 * (mesh.cell_bounds[:, [0, 2]] + mesh.cell_bounds[:, [1, 3]])
    topo = 50.0 * np.exp(-0.5 * ((cc[:, 0] / 400.0) ** 2 + (cc[:, 1] / 400.0) ** 2))
    top, bottom = topo - 30.0, topo - 80.0
    rng = np.random.default_rng(3)
    loc = np.c_[rng.uniform(-1400, 1400, 400), rng.uniform(-1200, 1200, 400)]
    loc = np.c_[loc, 50.0 * np.exp(-0.5 * ((loc[:, 0] / 400.0) ** 2 + (loc[:, 1] / 400.0) ** 2)) + 5.0]
    rx = gravity.receivers.Point(loc, components=["gz"])
    survey = gravity.survey.Survey(gravity.sources.SourceField(receiver_list=[rx]))
    sim = gravity.SimulationEquivalentSourceLayer(mesh, top, bottom, survey=survey,
                                                  rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                                  sensitivity_dtype=np.float64)
    G = sim.G.to_numpy()
    rows = rng.choice(400, 24, replace=False)
    ref = OracleLayerSimulation("gravity", mesh.cell_bounds, top, bottom, [(loc[rows], ["gz"])])
    assert rowwise_rel(G[rows], ref.sensitivity_matrix()).max() <= RTOL_NORMWISE
    model = rng.uniform(-0.5, 0.5, mesh.n_cells)
    full = OracleLayerSimulation("gravity", mesh.cell_bounds, top, bottom, [(loc, ["gz"])])
    assert vec_rel(sim.dpred(model), full.forward(model)) <= RTOL_NORMWISE
