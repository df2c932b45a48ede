"""
CPU: equivalent-source layer.

* the C restatement of the reference's 2D-mesh loops (oracle/pf_oracle.c ``pfo_layer_*``) against
  golden vectors produced by those loops themselves (gravity/_numba/_2d_mesh.py,
  magnetics/_numba/_2d_mesh.py, run unmodified by oracle/make_golden_layer.py): G and forward are
  bit-identical, the threaded column sums agree to rounding;
* the mapping the CUDA engine relies on: the layer's prisms, handed to the 3D node-sharing loops
  as (nodes, cell_nodes), give the same matrix (vertex-sum order aside);
* host-side behaviour of ``SimulationEquivalentSourceLayer`` (validation and error texts of
  simpeg/potential_fields/base.py:411-474).
"""
import numpy as np
import pytest

from helpers import vec_rel, rowwise_rel
from layer_helpers import LAYER_GOLDEN, layer_simulation, load_layer_golden
from oracle.oracle import OracleLayerSimulation, OracleSimulation
from simpeg_b200.meshes import CellBoundsMesh2D, TensorMesh
from simpeg_b200.potential_fields import gravity, magnetics
from simpeg_b200.potential_fields.base import layer_nodes_and_cells


def _oracle(case, **kw):
    return OracleLayerSimulation(case["physics"], case["bounds"], case["top"], case["bottom"],
                                 case["rlist"], b0=case["b0"], model_type=case["model_type"],
                                 backend="c", **kw)


def test_fixtures_exist():
    assert len(LAYER_GOLDEN) >= 10


@pytest.mark.parametrize("name", LAYER_GOLDEN)
def test_c_layer_oracle_matches_reference_golden(name):
    case = load_layer_golden(name)
    sim = _oracle(case)
    assert np.array_equal(sim.sensitivity_matrix(), case["G"])
    assert np.array_equal(sim.forward(case["model"]), case["fwd"])
    assert vec_rel(sim.gt_dot_v(case["v"]), case["gtv"]) < 1e-14
    assert vec_rel(sim.diag_gtg(case["w"]), case["diag_ref"]) < 1e-14


@pytest.mark.parametrize("name", LAYER_GOLDEN)
def test_golden_layer_consistency(name):
    case = load_layer_golden(name)
    assert vec_rel(case["G"] @ case["model"], case["fwd"]) < 1e-12
    assert vec_rel(case["G"].T @ case["v"], case["gtv"]) < 1e-12


@pytest.mark.parametrize("name", LAYER_GOLDEN)
def test_layer_prisms_as_shared_node_mesh(name):
    """What engine='cuda' does: (nodes, cell_nodes) of the prisms -> the 3D loops.  1e-9 norm-wise
    per row (observed ~1e-15: only the order of the 8-term vertex sum differs)."""
    case = load_layer_golden(name)
    nodes, cell_nodes = layer_nodes_and_cells(case["bounds"], case["top"], case["bottom"])
    sim3d = OracleSimulation(case["physics"], nodes, cell_nodes, case["rlist"], b0=case["b0"],
                             model_type=case["model_type"], backend="c")
    assert rowwise_rel(sim3d.sensitivity_matrix(), case["G"]).max() <= 1e-9
    assert vec_rel(sim3d.forward(case["model"]), case["fwd"]) <= 1e-9
    flat = np.all(case["top"] == case["top"][0]) and np.all(case["bottom"] == case["bottom"][0])
    assert nodes.shape[0] < 8 * case["bounds"].shape[0] if flat else True


def test_layer_nodes_vertex_order_and_sharing():
    mesh = TensorMesh([np.full(4, 2.0), np.full(3, 5.0)], "CC")
    bounds = mesh.cell_bounds
    nodes, cn = layer_nodes_and_cells(bounds, np.full(12, -1.0), np.full(12, -3.0))
    assert nodes.shape == (5 * 4 * 2, 3)           # a flat layer shares every coincident vertex
    c = 6
    x0, x1, y0, y1 = bounds[c]
    expect = [(x0, y0, -3.0), (x1, y0, -3.0), (x0, y1, -3.0), (x1, y1, -3.0),
              (x0, y0, -1.0), (x1, y0, -1.0), (x0, y1, -1.0), (x1, y1, -1.0)]
    assert np.array_equal(nodes[cn[c]], np.array(expect))
    rng = np.random.default_rng(0)
    nodes, cn = layer_nodes_and_cells(bounds, -1.0 - rng.uniform(size=12), -4.0 - rng.uniform(size=12))
    assert nodes.shape[0] == 8 * 12                # nothing to share when every face is its own


def test_layer_simulation_validation():
    mesh2 = TensorMesh([np.full(4, 2.0), np.full(3, 5.0)], "CC")
    mesh3 = TensorMesh([np.full(4, 2.0), np.full(3, 5.0), np.full(2, 1.0)], "CCN")
    for cls in (gravity.SimulationEquivalentSourceLayer, magnetics.SimulationEquivalentSourceLayer):
        # base.py:462-469
        with pytest.raises(ValueError, match="mesh must be 2D, received a 3D mesh"):
            cls(mesh3, 0.0, -1.0)
        # base.py:436-440
        with pytest.raises(AttributeError):
            cls(mesh2, np.zeros(5), -1.0)
        sim = cls(mesh2, 0, -2.0)
        assert sim.cell_z_top.shape == (12,) and np.all(sim.cell_z_bottom == -2.0)
        active = np.zeros(12, dtype=bool)
        active[::2] = True
        sim = cls(mesh2, np.zeros(6), -np.ones(6), active_cells=active)
        assert sim.nC == 6
        nodes, cn = sim._get_active_nodes()
        assert cn.shape == (6, 8)
    with pytest.raises(ValueError, match="mesh must be 3D, received a 2D mesh"):
        gravity.Simulation3DIntegral(mesh2)


def test_flat_tensor_layer_uses_the_marching_plan():
    mesh2 = TensorMesh([np.full(40, 2.0), np.full(30, 5.0)], "CC")
    sim = gravity.SimulationEquivalentSourceLayer(mesh2, -1.0, -4.0)
    plan = sim._tile_plan()
    assert plan.grid_shape == (40, 30, 1)
    draped = gravity.SimulationEquivalentSourceLayer(mesh2, -1.0 - np.linspace(0, 1, 1200), -4.0)
    plan = draped._tile_plan()
    assert plan.grid_shape is None and plan.n_cells == 1200
    tree = gravity.SimulationEquivalentSourceLayer(CellBoundsMesh2D(mesh2.cell_bounds), -1.0, -4.0)
    assert tree._tile_plan().n_cells == 1200


def test_layer_helper_builds_simulations():
    case = load_layer_golden(LAYER_GOLDEN[0])
    sim = layer_simulation(case)
    assert sim.survey.nD == case["G"].shape[0]


def test_generic_tile_plan_is_cached_on_the_mesh():
    mesh2 = CellBoundsMesh2D(TensorMesh([np.full(12, 2.0), np.full(9, 5.0)], "CC").cell_bounds)
    top = -1.0 - np.linspace(0, 1, mesh2.n_cells)
    a = gravity.SimulationEquivalentSourceLayer(mesh2, top, -4.0)
    b = magnetics.SimulationEquivalentSourceLayer(mesh2, top, -4.0)
    assert a._tile_plan() is b._tile_plan()            # same mesh, cells and faces: planned once
    c = gravity.SimulationEquivalentSourceLayer(mesh2, top - 0.5, -4.0)
    assert c._tile_plan() is not a._tile_plan()        # different prisms: new plan
    active = np.ones(mesh2.n_cells, dtype=bool)
    active[::3] = False
    d = gravity.SimulationEquivalentSourceLayer(mesh2, top[active], -4.0, active_cells=active)
    assert d._tile_plan().n_cells == int(active.sum())
