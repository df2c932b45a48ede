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
    assert vec_rel(fo.Jtvec(case["model"], case["v"]), case["gtv"]) <= RTOL_NORMWISE
    expected = np.einsum("i,ij,ij->j", case["w"], case["G"], case["G"])
    assert vec_rel(fo._gtg_diagonal_without_building_g(case["w"]), expected) <= RTOL_NORMWISE


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


# ---------------------------------------------------------------------------
# mirrors of tests/pf/test_equivalent_sources.py (reference), engine="cuda"
# ---------------------------------------------------------------------------
def _coordinates():
    x, y = np.meshgrid(np.linspace(-50, 50, 11), np.linspace(-50, 50, 11))
    return np.c_[x.ravel(order="F"), y.ravel(order="F"), np.full(x.size, 5.0)]


def _tensor_mesh_2d():
    # mesh_builder_xyz(coordinates, h=[5, 5], padding 50 m): 40 x 40 cells of 5 m around the data
    return TensorMesh([np.full(40, 5.0), np.full(40, 5.0)], origin=[-100.0, -100.0])


def _block_model(mesh, value):
    cc = 0.5 * (mesh.cell_bounds[:, [0, 2]] + mesh.cell_bounds[:, [1, 3]])
    inside = (np.abs(cc[:, 0]) <= 20) & (np.abs(cc[:, 1]) <= 20)
    return np.where(inside, value, 0.0)


@pytest.mark.parametrize("store", ["ram", "forward_only"])
@pytest.mark.parametrize("components", ["gz", ["gz", "gzz"], ["gx", "gxy", "guv"]])
def test_gravity_layer_vs_3d_simulation(store, components):
    """test_equivalent_sources.py:255-328: the layer equals the 3D integral simulation on the
    analogous one-cell-thick 3D mesh (atol 1e-7)."""
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity

    mesh = _tensor_mesh_2d()
    top, bottom = -20.0, -50.0
    mesh_3d = TensorMesh([mesh.h[0], mesh.h[1], np.array([top - bottom])], origin=[-100.0, -100.0, bottom])
    rx = gravity.receivers.Point(_coordinates(), components=components)
    survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
    mapping = maps.IdentityMap(nP=mesh.n_cells)
    sim_3d = gravity.Simulation3DIntegral(mesh_3d, survey=survey, rhoMap=mapping, sensitivity_dtype=np.float64)
    eq = gravity.SimulationEquivalentSourceLayer(mesh, top, bottom, survey=survey, rhoMap=mapping,
                                                 store_sensitivities=store, sensitivity_dtype=np.float64)
    model = _block_model(mesh, 2.67)
    np.testing.assert_allclose(sim_3d.dpred(model), eq.dpred(model), atol=1e-7)


@pytest.mark.parametrize("model_type", ["scalar", "vector"])
@pytest.mark.parametrize("components", ["tmi", ["bx", "by", "bz"], ["tmi_x", "bzz"]])
def test_magnetic_layer_vs_3d_simulation_with_active_cells(model_type, components):
    """test_equivalent_sources.py:610-749 (forward vs 3D simulation, with active cells)."""
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import magnetics

    mesh = _tensor_mesh_2d()
    top, bottom = -20.0, -50.0
    mesh_3d = TensorMesh([mesh.h[0], mesh.h[1], np.array([top - bottom])], origin=[-100.0, -100.0, bottom])
    cc = 0.5 * (mesh.cell_bounds[:, [0, 2]] + mesh.cell_bounds[:, [1, 3]])
    active = (np.abs(cc[:, 0]) < 75) & (np.abs(cc[:, 1]) < 75)
    n = int(active.sum()) * (3 if model_type == "vector" else 1)
    rx = magnetics.receivers.Point(_coordinates(), components=components)
    src = magnetics.sources.UniformBackgroundField([rx], amplitude=51_000.0, inclination=71.0,
                                                   declination=12.0)
    survey = magnetics.survey.Survey(src)
    kw = dict(survey=survey, chiMap=maps.IdentityMap(nP=n), model_type=model_type,
              active_cells=active, sensitivity_dtype=np.float64)
    sim_3d = magnetics.Simulation3DIntegral(mesh_3d, **kw)
    eq = magnetics.SimulationEquivalentSourceLayer(mesh, top, bottom, **kw)
    chi = _block_model(mesh, 0.2e-3)[active]
    model = chi if model_type == "scalar" else np.hstack([chi, 0.3 * chi, -0.5 * chi])
    np.testing.assert_allclose(sim_3d.dpred(model), eq.dpred(model), atol=1e-7)


@pytest.mark.parametrize("top_bottom_as_array", [False, True])
def test_fit_gravity_equivalent_sources(top_bottom_as_array):
    """test_equivalent_sources.py:874-927: equivalent sources trained on synthetic data reproduce
    them (atol 0.005 mGal).  The reference runs its inversion stack (L2DataMisfit, ProjectedGNCG
    with cg_maxiter=20, cg_rtol=1e-3, UpdatePreconditioner, beta cooled by 3 per iteration down to
    the target misfit); here the same outer loop drives the device-resident CG."""
    from simpeg_b200 import maps
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG
    from simpeg_b200.potential_fields import gravity

    mesh = _tensor_mesh_2d()
    top, bottom = -20.0, -50.0
    if top_bottom_as_array:
        rng = np.random.default_rng(seed=42)
        top = np.full(mesh.n_cells, top) + rng.normal(scale=0.5, size=mesh.n_cells)
        bottom = np.full(mesh.n_cells, bottom) + rng.normal(scale=0.5, size=mesh.n_cells)
    rx = gravity.receivers.Point(_coordinates(), components="gz")
    survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
    sim = gravity.SimulationEquivalentSourceLayer(mesh, top, bottom, survey=survey,
                                                  rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                                  sensitivity_dtype=np.float64)
    std = 1e-3
    dobs = sim.dpred(_block_model(mesh, 2.67)) + np.random.default_rng(1).normal(scale=std, size=survey.nD)
    nD, nP = survey.nD, mesh.n_cells
    # phi = |W (G m - d)|^2 + beta |m|^2 ; H = 2 G^T W^2 G + beta * 2 I
    solver = DeviceGaussNewtonCG(sim, data_weights=np.full(nD, 1.0 / std), reg_hessian=2.0 * np.ones(nP),
                                 beta=0.0, cg_maxiter=20, cg_rtol=1e-3)
    m = np.zeros(nP)
    solver.set_model(m)
    v = np.random.default_rng(42).normal(size=nP)          # BetaEstimate_ByEig(beta0_ratio=1e-1)
    for _ in range(4):
        hv = solver.hessian_vector(v)
        lam, v = np.linalg.norm(hv) / np.linalg.norm(v), hv / np.linalg.norm(hv)
    solver.beta = 1e-1 * lam / 2.0
    for _ in range(20):
        resid = sim.dpred(m) - dobs
        if np.sum((resid / std) ** 2) <= nD:                # TargetMisfit(chifact=1)
            break
        g = 2.0 * sim.Jtvec(m, resid / std**2) + 2.0 * solver.beta * m
        solver.set_model(m)
        solver.jacobi_preconditioner()                      # UpdatePreconditioner
        m = m + solver.find_search_direction(g)             # linear problem: full Newton step
        solver.beta /= 3.0                                  # BetaSchedule(coolingFactor=3, coolingRate=1)
    np.testing.assert_allclose(sim.dpred(m), dobs, atol=0.005, rtol=1e-5)


def test_layer_edge_cases_single_cell_and_no_receivers():
    """One prism / an empty receiver object / no active cells (shapes only: nothing to evaluate)."""
    from oracle.oracle import OracleLayerSimulation
    from simpeg_b200 import maps
    from simpeg_b200.meshes import CellBoundsMesh2D
    from simpeg_b200.potential_fields import gravity

    mesh = CellBoundsMesh2D(np.array([[-5.0, 5.0, -2.0, 4.0]]))
    loc = np.array([[0.0, 1.0, 2.0], [20.0, -3.0, 1.0], [-5.0, -2.0, 0.5]])   # last one above a corner
    rx = gravity.receivers.Point(loc, components=["gz", "gxy"])
    empty = gravity.receivers.Point(np.zeros((0, 3)), components=["gx"])
    survey = gravity.survey.Survey(gravity.sources.SourceField([empty, rx]))
    sim = gravity.SimulationEquivalentSourceLayer(mesh, -1.0, -3.0, survey=survey,
                                                  rhoMap=maps.IdentityMap(nP=1), sensitivity_dtype=np.float64)
    ref = OracleLayerSimulation("gravity", mesh.cell_bounds, -1.0, -3.0,
                                [(np.zeros((0, 3)), ["gx"]), (loc, ["gz", "gxy"])])
    G = sim.G.to_numpy()
    assert G.shape == (6, 1)
    assert rowwise_rel(G, ref.sensitivity_matrix()).max() <= RTOL_NORMWISE
    assert vec_rel(sim.dpred(np.array([0.3])), ref.forward(np.array([0.3]))) <= RTOL_NORMWISE
    # no active cells at all
    none = gravity.SimulationEquivalentSourceLayer(mesh, np.zeros(0), np.zeros(0), survey=survey,
                                                   active_cells=np.array([False]),
                                                   rhoMap=maps.IdentityMap(nP=0), sensitivity_dtype=np.float64)
    assert none.G.shape == (6, 0)
    assert np.array_equal(none.dpred(np.zeros(0)), np.zeros(6))
