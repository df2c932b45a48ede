"""GPU: edge cases -- empty and tiny inputs, ragged receiver objects, a public-API run on a
tensor mesh large enough for the z-marching kernel with topography (inactive cells)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import RTOL_NORMWISE, rowwise_rel, vec_rel
from oracle.oracle import OracleSimulation
from sim_fixtures import oracle_for
from simpeg_b200 import maps
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.potential_fields import gravity, magnetics

pytestmark = pytest.mark.gpu


def test_single_cell_single_receiver():
    mesh = TensorMesh([[(2.0, 1)], [(3.0, 1)], [(4.0, 1)]], "CCN")
    loc = np.array([[0.3, -0.4, 1.5]])
    eng = CudaEngine("gravity", mesh.nodes, mesh.cell_nodes, [(loc, ["gz", "gxx"])])
    ora = OracleSimulation("gravity", mesh.nodes, mesh.cell_nodes, [(loc, ["gz", "gxx"])])
    eng.build_G()
    assert rowwise_rel(eng.G_to_numpy(), ora.sensitivity_matrix()).max() <= RTOL_NORMWISE
    m = np.array([0.7])
    assert vec_rel(eng.to_host(eng.forward_dev(eng.to_device(m))), ora.forward(m)) <= RTOL_NORMWISE


def test_empty_receiver_object_and_ragged_objects():
    mesh = TensorMesh([[(5.0, 6)], [(5.0, 5)], [(5.0, 4)]], "CCN")
    rng = np.random.default_rng(2)
    a = np.c_[rng.uniform(-10, 10, (3, 2)), np.full(3, 2.0)]
    b = np.c_[rng.uniform(-10, 10, (5, 2)), np.full(5, 3.0)]
    rlist = [(a, ["gz"]), (np.zeros((0, 3)), ["gxx", "gz"]), (b, ["gzz", "gxy", "gx"])]
    eng = CudaEngine("gravity", mesh.nodes, mesh.cell_nodes, rlist)
    ora = OracleSimulation("gravity", mesh.nodes, mesh.cell_nodes, rlist)
    assert eng.n_rows == 3 + 0 + 15
    eng.build_G()
    G = ora.sensitivity_matrix()
    assert rowwise_rel(eng.G_to_numpy(), G).max() <= RTOL_NORMWISE
    v = rng.uniform(size=eng.n_rows)
    assert vec_rel(eng.to_host(eng.gt_dot_v_dev(eng.to_device(v))), G.T @ v) <= RTOL_NORMWISE
    assert vec_rel(eng.to_host(eng.gtg_diag_dev(eng.to_device(v))), np.einsum("i,ij,ij->j", v, G, G)) <= RTOL_NORMWISE


def test_no_receivers_and_no_active_cells():
    mesh = TensorMesh([[(5.0, 4)], [(5.0, 4)], [(5.0, 4)]], "CCN")
    eng = CudaEngine("gravity", mesh.nodes, mesh.cell_nodes, [(np.zeros((0, 3)), ["gz"])])
    assert eng.n_rows == 0
    eng.build_G()
    out = eng.to_host(eng.gt_dot_v_dev(eng.to_device(np.zeros(0))))
    assert out.shape == (64,) and not out.any()
    # all cells inactive
    rx = gravity.receivers.Point(np.array([[0.0, 0.0, 1.0]]), components="gz")
    survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
    sim = gravity.Simulation3DIntegral(mesh, survey=survey, rhoMap=maps.IdentityMap(nP=0),
                                       active_cells=np.zeros(64, dtype=bool), sensitivity_dtype=np.float64)
    assert sim.nC == 0
    d = sim.dpred(np.zeros(0))
    assert d.shape == (1,) and d[0] == 0.0


@pytest.mark.parametrize("store", ["ram", "forward_only"])
def test_public_api_tensor_mesh_with_topography(store):
    """Simulation3DIntegral on a mesh that takes the marching-kernel path, with active_cells."""
    mesh = TensorMesh([[(4.0, 30)], [(4.0, 24)], [(3.0, 10)]], "CCN")
    cc = mesh.cell_centers
    active = cc[:, 2] < -4.0 + 3.0 * np.exp(-((cc[:, 0] / 30.0) ** 2 + (cc[:, 1] / 25.0) ** 2))
    rng = np.random.default_rng(4)
    loc = np.c_[rng.uniform(-50, 50, (10, 2)), np.full(10, 2.0)]
    rx = magnetics.receivers.Point(loc, components=["tmi"])
    src = magnetics.sources.UniformBackgroundField([rx], 52000.0, 65.0, 20.0)
    survey = magnetics.survey.Survey(src)
    nC = int(active.sum())
    sim = magnetics.Simulation3DIntegral(mesh, survey=survey, chiMap=maps.IdentityMap(nP=3 * nC),
                                         model_type="vector", active_cells=active,
                                         store_sensitivities=store, sensitivity_dtype=np.float64)
    assert sim._dev.use_marching and sim._dev.plan.grid_colmap
    ora = oracle_for(sim)
    G = ora.sensitivity_matrix()
    chi = rng.uniform(-0.05, 0.05, 3 * nC)
    v = rng.uniform(size=survey.nD)
    assert vec_rel(sim.dpred(chi), G @ chi) <= RTOL_NORMWISE
    assert vec_rel(sim.Jtvec(chi, v), G.T @ v) <= RTOL_NORMWISE
    w = rng.uniform(0.5, 1.5, size=survey.nD)
    assert vec_rel(sim.getJtJdiag(chi, W=sp.diags(w)), np.einsum("i,ij,ij->j", w**2, G, G)) <= RTOL_NORMWISE
    if store == "ram":
        assert rowwise_rel(sim.G.to_numpy(), G).max() <= RTOL_NORMWISE


@pytest.mark.parametrize("physics,comps,mt", [
    ("gravity", ["gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv"], "scalar"),
    ("magnetics", ["tmi", "bx", "by", "bz"], "vector"),
    ("magnetics", ["tmi_x", "tmi_y", "tmi_z", "bxx", "byz"], "scalar"),
])
def test_receivers_inside_below_and_beside_the_mesh(physics, comps, mt):
    """Borehole-style geometry: nodes on every side of the receivers (all sign combinations of
    e, n, u, including zeros: receivers in node planes, on cell faces and on mesh edges' lines)."""
    mesh = TensorMesh([[(3.0, 14)], [(4.0, 12)], [(2.5, 12)]], "CCN")
    rng = np.random.default_rng(12)
    inside = np.c_[rng.uniform(-15, 15, (6, 2)), rng.uniform(-25, -5, 6)]      # inside cells (generic)
    below = np.c_[rng.uniform(-30, 30, (3, 2)), np.full(3, -45.0)]             # below the mesh
    beside = np.array([[40.0, 3.3, -12.2], [-3.1, -39.0, -7.7]])               # outside, at depth
    on_planes = np.array([[mesh.nodes_x[5], 1.234, 3.0],                        # e == 0 for a node plane
                          [0.77, mesh.nodes_y[4], 2.0],                         # n == 0
                          [mesh.nodes_x[9], mesh.nodes_y[7], 4.0],              # e == n == 0 (above a node)
                          [50.0, 2.2, mesh.nodes_z[6]]])                        # u == 0, outside the mesh
    loc = np.vstack([inside, below, beside, on_planes])
    b0 = np.array([1.2e4, -2.1e4, -4.4e4])
    from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of
    plan = build_tile_plan_tensor(tensor_axes_of(mesh))
    ora = OracleSimulation(physics, mesh.nodes, mesh.cell_nodes, [(loc, comps)], b0=b0, model_type=mt)
    G = ora.sensitivity_matrix()
    assert np.isfinite(G).all()
    for march in (True, False):
        eng = CudaEngine(physics, None, None, [(loc, comps)], b0=b0, model_type=mt, plan=plan,
                         use_marching=march, use_node_space=march)
        eng.build_G()
        assert rowwise_rel(eng.G_to_numpy(), G).max() <= RTOL_NORMWISE
        m = rng.uniform(-0.5, 0.5, eng.nP)
        v = rng.uniform(size=eng.n_rows)
        assert vec_rel(eng.to_host(eng.forward_dev(eng.to_device(m))), G @ m) <= RTOL_NORMWISE
        assert vec_rel(eng.to_host(eng.gt_dot_v_dev(eng.to_device(v))), G.T @ v) <= RTOL_NORMWISE


def test_store_sensitivities_disk_roundtrip(tmp_path):
    """store_sensitivities='disk' (gravity/simulation.py:599-606): G is built on the GPU and also
    written as the reference's raw row-major memmap; load_sensitivities() of a second simulation
    -- or of a file the reference wrote -- uploads it instead of evaluating it."""
    from helpers import RTOL_NORMWISE, load_golden, vec_rel
    from sim_fixtures import gravity_survey, small_mesh
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity

    mesh = small_mesh()
    path = str(tmp_path / "sens.bin")
    kw = dict(survey=gravity_survey(["gz", "gzz"]), rhoMap=maps.IdentityMap(nP=mesh.n_cells),
              sensitivity_dtype=np.float64, sensitivity_path=path)
    sim = gravity.Simulation3DIntegral(mesh, store_sensitivities="disk", **kw)
    rho = np.random.default_rng(0).uniform(-0.5, 0.5, mesh.n_cells)
    d = sim.dpred(rho)
    G = sim.G.to_numpy()
    on_disk = np.memmap(path, dtype=np.float64, mode="r", shape=G.shape, order="C")
    assert np.array_equal(np.asarray(on_disk), G)
    # a second simulation reads the file; make sure it does not evaluate anything
    sim2 = gravity.Simulation3DIntegral(mesh, store_sensitivities="ram", **kw)
    sim2._sensitivity_matrix = None          # calling it would raise
    sim2.load_sensitivities()
    assert np.array_equal(sim2.dpred(rho), d)
    assert np.array_equal(sim2.G.to_numpy(), G)
    # float32 file as the reference writes it by default (base.py:86)
    path32 = str(tmp_path / "sens32.bin")
    ref32 = np.memmap(path32, dtype=np.float32, mode="w+", shape=G.shape, order="C")
    ref32[:] = G.astype(np.float32)
    ref32.flush()
    kw32 = dict(kw, sensitivity_dtype=np.float32, sensitivity_path=path32)
    sim3 = gravity.Simulation3DIntegral(mesh, store_sensitivities="ram", **kw32)
    sim3.load_sensitivities()
    assert vec_rel(sim3.dpred(rho), G @ rho) <= 1e-6


def test_store_sensitivities_disk_zarr_roundtrip(tmp_path):
    """A sensitivity_path that ends in .zarr: the matrix goes to an uncompressed zarr (format 2)
    directory store in row chunks, the container of the reference's dask path
    (simpeg/dask/potential_fields/base.py:46-66); load_sensitivities() takes the store itself, or
    the directory that holds `sensitivity.zarr`, and uploads the rows instead of evaluating them."""
    import json

    from sim_fixtures import gravity_survey, small_mesh
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity
    from simpeg_b200.zarr_store import ZarrMatrixReader

    mesh = small_mesh()
    store = tmp_path / "sensitivity.zarr"
    kw = dict(survey=gravity_survey(["gz", "gxx"]), rhoMap=maps.IdentityMap(nP=mesh.n_cells),
              sensitivity_dtype=np.float32, sensitivity_path=str(store))
    sim = gravity.Simulation3DIntegral(mesh, store_sensitivities="disk", **kw)
    rho = np.random.default_rng(0).uniform(-0.5, 0.5, mesh.n_cells)
    d = sim.dpred(rho)
    G = sim.G.to_numpy()
    meta = json.load(open(store / ".zarray"))
    assert meta["shape"] == list(G.shape) and meta["compressor"] is None and meta["dtype"] == "<f4"
    assert np.array_equal(ZarrMatrixReader(store).read_rows(np.arange(G.shape[0])), G.astype(np.float32))
    for path in (str(store), str(tmp_path)):          # the store, or the dask path's directory around it
        sim2 = gravity.Simulation3DIntegral(mesh, store_sensitivities="ram", **dict(kw, sensitivity_path=path))
        sim2._sensitivity_matrix = None               # calling it would raise
        sim2.load_sensitivities()
        assert np.array_equal(sim2.dpred(rho), d)
    # a store of another shape is refused
    other = gravity.Simulation3DIntegral(mesh, store_sensitivities="ram",
                                         **dict(kw, survey=gravity_survey(["gz"])))
    with pytest.raises(ValueError, match="zarr store holds"):
        other.load_sensitivities()


@pytest.mark.parametrize("components", [["gxx", "gyy", "gzz"], ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"]],
                         ids=["trace", "cfg2-set"])
def test_derived_gzz_with_receivers_touching_cells(components):
    """gzz is derived from gxx and gyy in the z-marching kernel (no plane of its own); the term
    sum_v s_v S_v it drops is non-zero exactly when the receiver touches a prism.  Receivers on
    the top face of the mesh -- inside a face, on an edge, on a node -- and one buried inside a
    cell must still match the oracle (the reference's singular-branch conventions) and the tile
    kernel, which evaluates kernel_uu at every node."""
    from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of

    mesh = TensorMesh([[(4.0, 40)], [(5.0, 30)], [(3.0, 12)]], "CCN")
    xs, ys, zs = mesh.nodes_x, mesh.nodes_y, mesh.nodes_z
    loc = np.array([
        [xs[5] + 1.3, ys[4] + 2.1, 0.0],              # inside a top face
        [xs[16], ys[13] + 1.7, 0.0],                  # on an edge of the top face
        [xs[27], ys[15], 0.0],                        # on a node of the top face (footprint border row)
        [xs[2] + 0.9, ys[22] + 1.1, zs[3] + 1.4],     # inside a cell
        [xs[33], ys[6] + 2.0, zs[7]],                 # on an interior edge
        [1.234, -2.345, 2.5],                         # ordinary receiver above the mesh
    ])
    plan = build_tile_plan_tensor(tensor_axes_of(mesh), None)
    march = CudaEngine("gravity", None, None, [(loc, components)], plan=plan, use_marching=True)
    tile = CudaEngine("gravity", None, None, [(loc, components)], plan=plan, use_marching=False)
    assert "pf_march_kernel" in march.describe("store") and "pf_tile_kernel" in tile.describe("store")
    ora = OracleSimulation("gravity", mesh.nodes, mesh.cell_nodes, [(loc, components)])
    G = ora.sensitivity_matrix()
    march.build_G()
    tile.build_G()
    assert rowwise_rel(march.G_to_numpy(), G).max() <= RTOL_NORMWISE
    assert rowwise_rel(march.G_to_numpy(), tile.G_to_numpy()).max() <= 1e-12
    m = np.random.default_rng(4).uniform(-0.5, 0.5, mesh.n_cells)
    assert vec_rel(march.to_host(march.forward_dev(march.to_device(m))), ora.forward(m)) <= RTOL_NORMWISE
    w = np.random.default_rng(5).uniform(0.5, 1.5, march.n_rows)
    assert vec_rel(march.to_host(march.gtg_diag_dev(march.to_device(w))),
                   np.einsum("i,ij,ij->j", w, G, G)) <= RTOL_NORMWISE
