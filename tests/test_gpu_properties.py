"""
GPU: parity at the shapes BASELINE.json names, through (a) the oracle on a
receiver subsample of the FULL mesh and (b) size-independent properties:
adjointness <G x, y> = <x, G^T y>, linearity, stored-G vs matrix-free
agreement, row-block independence (what the multi-GPU sharding relies on),
tree meshes with hanging nodes, determinism.
"""
import numpy as np
import pytest

from helpers import RTOL_NORMWISE, rowwise_rel, vec_rel
from oracle.oracle import OracleSimulation
from simpeg_b200.engine import CudaEngine
from simpeg_b200.meshes import TensorMesh, synthetic_octree
from simpeg_b200.tiling import build_tile_plan_tensor, march_footprints, tensor_axes_of

pytestmark = pytest.mark.gpu

B0 = np.array([-8.55e3, -2.35e4, -4.33e4])   # 50 000 nT, I=60, D=250 (test_forward_Mag_Linear.py:214)


def grid_receivers(n, half, z, seed):
    side = int(np.ceil(np.sqrt(n)))
    xs = np.linspace(-half, half, side)
    X, Y = np.meshgrid(xs, xs)
    loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, z)][:n]
    loc[:, 2] += np.random.default_rng(seed).uniform(0, 1, n)
    return loc


def tensor_engine(mesh, physics, rlist, **kw):
    plan = build_tile_plan_tensor(tensor_axes_of(mesh))
    return CudaEngine(physics, None, None, rlist, plan=plan, **kw)


CONFIGS = {
    # BASELINE configs 1, 2, 3 (mesh shape / spacing / components); receivers subsampled
    "cfg1_gz": dict(shape=(40, 40, 20), h=5.0, z=1.0, physics="gravity", comps=["gz"], mt="scalar", n_rx=441),
    "cfg2_tensor7": dict(shape=(100, 100, 50), h=10.0, z=5.0, physics="gravity",
                         comps=["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], mt="scalar", n_rx=96),
    "cfg3_tmi_scalar": dict(shape=(200, 200, 100), h=25.0, z=30.0, physics="magnetics", comps=["tmi"],
                            mt="scalar", n_rx=64),
    "cfg4_tmi_vector": dict(shape=(100, 100, 100), h=10.0, z=10.0, physics="magnetics", comps=["tmi"],
                            mt="vector", n_rx=128),
}


def make(cfg):
    mesh = TensorMesh([np.full(n, cfg["h"]) for n in cfg["shape"]], "CCN")
    loc = grid_receivers(cfg["n_rx"], cfg["shape"][0] * cfg["h"] / 2, cfg["z"], 1)
    eng = tensor_engine(mesh, cfg["physics"], [(loc, cfg["comps"])], b0=B0, model_type=cfg["mt"])
    return mesh, loc, eng


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_mesh_subsample_vs_oracle(name):
    """G rows / forward / G^T v of a few receivers over the FULL mesh against the oracle."""
    cfg = CONFIGS[name]
    mesh, loc, eng = make(cfg)
    sub = loc[:: max(1, len(loc) // 4)][:4]
    ora = OracleSimulation(cfg["physics"], mesh.nodes, mesh.cell_nodes, [(sub, cfg["comps"])], b0=B0,
                           model_type=cfg["mt"])
    sub_eng = tensor_engine(mesh, cfg["physics"], [(sub, cfg["comps"])], b0=B0, model_type=cfg["mt"])
    G_ref = ora.sensitivity_matrix()
    sub_eng.build_G()
    assert rowwise_rel(sub_eng.G_to_numpy(), G_ref).max() <= RTOL_NORMWISE
    rng = np.random.default_rng(4)
    lo, hi = (-0.5, 0.5) if cfg["physics"] == "gravity" else (0.0, 0.05)
    m = rng.uniform(lo, hi, sub_eng.nP)
    v = rng.uniform(size=sub_eng.n_rows)
    assert vec_rel(sub_eng.to_host(sub_eng.forward_dev(sub_eng.to_device(m))), ora.forward(m)) <= RTOL_NORMWISE
    assert vec_rel(sub_eng.to_host(sub_eng.gt_dot_v_dev(sub_eng.to_device(v))), ora.gt_dot_v(v)) <= RTOL_NORMWISE
    assert vec_rel(sub_eng.to_host(sub_eng.gemv_dev(sub_eng.to_device(m))), G_ref @ m) <= RTOL_NORMWISE
    assert vec_rel(sub_eng.to_host(sub_eng.gemv_t_dev(sub_eng.to_device(v))), G_ref.T @ v) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_adjoint_linearity_and_store_modes(name):
    cfg = CONFIGS[name]
    mesh, loc, eng = make(cfg)
    rng = np.random.default_rng(9)
    x1, x2 = eng.to_device(rng.normal(size=eng.nP)), eng.to_device(rng.normal(size=eng.nP))
    y = eng.to_device(rng.normal(size=eng.n_rows))
    eng.build_G()
    Gx1, Gx2 = eng.gemv_dev(x1), eng.gemv_dev(x2)
    Gty = eng.gemv_t_dev(y)
    # adjointness of the stored-G products
    lhs, rhs = float(Gx1 @ y), float(x1 @ Gty)
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), float(Gx1.norm() * y.norm()))
    # linearity
    comb = eng.gemv_dev(2.5 * x1 - 0.75 * x2)
    assert vec_rel(eng.to_host(comb), eng.to_host(2.5 * Gx1 - 0.75 * Gx2)) <= 1e-11
    # matrix-free == stored
    assert vec_rel(eng.to_host(eng.forward_dev(x1)), eng.to_host(Gx1)) <= RTOL_NORMWISE
    assert vec_rel(eng.to_host(eng.gt_dot_v_dev(y)), eng.to_host(Gty)) <= RTOL_NORMWISE
    w = eng.to_device(rng.uniform(0.5, 1.5, size=eng.n_rows))
    assert vec_rel(eng.to_host(eng.gtg_diag_dev(w)), eng.to_host(eng.colsumsq_dev(w))) <= RTOL_NORMWISE
    # determinism: same bits on a second evaluation
    assert np.array_equal(eng.to_host(eng.gemv_t_dev(y)), eng.to_host(Gty))
    assert np.array_equal(eng.to_host(eng.forward_dev(x1)), eng.to_host(eng.forward_dev(x1)))


MARCH_CASES = [
    ("gravity", ["gz"], "scalar"),
    ("gravity", ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar"),
    ("gravity", ["gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv"], "scalar"),
    ("magnetics", ["tmi"], "scalar"),
    ("magnetics", ["tmi"], "vector"),
    ("magnetics", ["bx", "by", "bz"], "vector"),
    ("magnetics", ["tmi", "bx"], "scalar"),
    ("magnetics", ["tmi_x", "tmi_y", "tmi_z"], "scalar"),
    ("magnetics", ["bxx", "bxy", "bxz", "byy", "byz", "bzz", "tmi_z"], "vector"),
]


@pytest.mark.parametrize("physics,comps,mt", MARCH_CASES, ids=lambda v: "+".join(v) if isinstance(v, list) else v)
@pytest.mark.parametrize("masked", [False, True], ids=["all_active", "topography"])
def test_marching_kernel_vs_tile_kernel_and_oracle(physics, comps, mt, masked):
    """The z-marching kernel (tensor meshes) against the generic tile kernel and the oracle:
    G rows and forward_only, all component families, with and without inactive cells."""
    mesh = TensorMesh([[(4.0, 3, -1.2), (4.0, 34), (4.0, 3, 1.2)], [(5.0, 30)], [(3.0, 12)]], "CCN")
    active = None
    if masked:
        cc = mesh.cell_centers
        active = np.nonzero(cc[:, 2] < -6.0 + 5.0 * np.exp(-((cc[:, 0] / 40.0) ** 2 + (cc[:, 1] / 50.0) ** 2)))[0]
    loc = grid_receivers(12, 60.0, 2.0, 3)
    # generic positions (a receiver ~1e-14 m off a node column makes the reference's own
    # third-order kernels overflow: -x_i / ((x_k + r) r) with x_k + r == 0) ...
    loc[:, :2] += np.random.default_rng(5).uniform(-1.0, 1.0, (12, 2))
    loc[0, :2] = mesh.nodes_x[20], mesh.nodes_y[11]        # ... and one exactly above a node
    plan = build_tile_plan_tensor(tensor_axes_of(mesh), active)
    assert plan.footprints4.shape[0] > 0 and plan.footprints2.shape[0] > 0
    kw = dict(b0=B0, model_type=mt, plan=plan)
    march = CudaEngine(physics, None, None, [(loc, comps)], use_marching=True, **kw)
    tile = CudaEngine(physics, None, None, [(loc, comps)], use_marching=False, **kw)
    assert march.plan.grid_nx == 40 and tile.plan.grid_nx == 0
    cn = mesh.cell_nodes if active is None else mesh.cell_nodes[active]
    uniq, inv = np.unique(cn, return_inverse=True)
    ora = OracleSimulation(physics, mesh.nodes[uniq], inv.reshape(cn.shape), [(loc, comps)], b0=B0, model_type=mt)
    G_ref = ora.sensitivity_matrix()
    march.build_G()
    tile.build_G()
    Gm, Gt = march.G_to_numpy(), tile.G_to_numpy()
    assert rowwise_rel(Gm, G_ref).max() <= RTOL_NORMWISE
    assert rowwise_rel(Gm, Gt).max() <= 1e-12
    rng = np.random.default_rng(8)
    m = rng.uniform(-0.5, 0.5, march.nP)
    fm = march.to_host(march.forward_dev(march.to_device(m)))
    assert vec_rel(fm, ora.forward(m)) <= RTOL_NORMWISE
    assert vec_rel(fm, tile.to_host(tile.forward_dev(tile.to_device(m)))) <= 1e-11
    # on-the-fly G^T v: node-space kernel (default) vs tile kernel vs oracle
    v = rng.uniform(size=march.n_rows)
    gtv_node = march.to_host(march.gt_dot_v_dev(march.to_device(v)))
    assert march.plan.n_nodes > 0
    tile_only = CudaEngine(physics, None, None, [(loc, comps)], use_marching=False, use_node_space=False, **kw)
    gtv_tile = tile_only.to_host(tile_only.gt_dot_v_dev(tile_only.to_device(v)))
    assert tile_only.plan.n_nodes == 0
    assert vec_rel(gtv_node, ora.gt_dot_v(v)) <= RTOL_NORMWISE
    assert vec_rel(gtv_node, gtv_tile) <= 1e-11
    # float32 storage through the marching kernel
    m32 = CudaEngine(physics, None, None, [(loc, comps)], sensitivity_dtype=np.float32, **kw)
    m32.build_G()
    assert rowwise_rel(m32.G_to_numpy(), G_ref).max() <= 1e-6


@pytest.mark.parametrize("shape", [(151, 7, 3), (3, 60, 4), (255, 2, 2), (300, 5, 2), (64, 64, 3), (1, 1, 5), (37, 23, 6)],
                         ids=lambda s: "x".join(map(str, s)))
def test_marching_kernel_footprint_shapes(shape):
    """Node phase of the z-marching kernel on footprints of every aspect ratio: one node row per
    round (151 or 255 nodes wide), 64 rows per round (4 nodes wide), full-width meshes that need
    several footprints, non-uniform spacing, a single cell column.  G rows, forward and the
    blocked diagonal against the oracle; TMI vector (3 planes, 1024-node footprints) and the
    seven-component gravity set (6 planes, 512-node footprints, derived gzz)."""
    rng = np.random.default_rng(21)
    hx, hy, hz = (rng.uniform(2.0, 5.0, n) for n in shape)
    mesh = TensorMesh([hx, hy, hz], "CCN")
    # receivers over the mesh (far outside it the reference's own rounding exceeds the bar, see
    # test_altitude_parity.py)
    # ... and off its symmetry planes (rows that vanish by symmetry have no row-wise scale)
    loc = grid_receivers(6, 1.0, 1.5, 4)
    loc[:, :2] += rng.uniform(-0.3, 0.3, (6, 2))
    loc[:, 0] *= 0.4 * hx.sum()
    loc[:, 1] *= 0.4 * hy.sum()
    loc[0] = [mesh.nodes_x[shape[0] // 2], mesh.nodes_y[shape[1] // 2], 0.0]    # on a surface node
    plan = build_tile_plan_tensor(tensor_axes_of(mesh))
    # the planner drops footprint sets that are no better than bricks on such meshes: force them
    plan.footprints4 = march_footprints(*shape, 1024)
    plan.footprints2 = march_footprints(*shape, 512)
    plan.footprints4h = march_footprints(*shape, 512, threads=128)     # 128-thread CTAs (TMI vector)
    for physics, comps, mt in (("magnetics", ["tmi"], "vector"),
                               ("gravity", ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"], "scalar")):
        eng = CudaEngine(physics, None, None, [(loc, comps)], b0=B0, model_type=mt, plan=plan, use_marching=True)
        assert eng.plan.grid_nx == shape[0] and "pf_march_kernel" in eng.describe("store")
        assert eng.describe("store").endswith(",4,128>") == (mt == "vector")
        ora = OracleSimulation(physics, mesh.nodes, mesh.cell_nodes, [(loc, comps)], b0=B0, model_type=mt)
        G_ref = ora.sensitivity_matrix()
        eng.build_G()
        assert rowwise_rel(eng.G_to_numpy(), G_ref).max() <= RTOL_NORMWISE
        m = rng.uniform(-0.5, 0.5, eng.nP)
        assert vec_rel(eng.to_host(eng.forward_dev(eng.to_device(m))), G_ref @ m) <= RTOL_NORMWISE
        w = rng.uniform(0.5, 1.5, eng.n_rows)
        assert vec_rel(eng.to_host(eng.gtg_diag_dev(eng.to_device(w))),
                       np.einsum("i,ij,ij->j", w, G_ref, G_ref)) <= RTOL_NORMWISE


def test_row_blocks_are_independent():
    """Rows built from a receiver block equal the same rows of the whole build (sharding invariant)."""
    cfg = CONFIGS["cfg4_tmi_vector"]
    mesh, loc, eng = make(cfg)
    eng.build_G()
    lo, hi = 37, 91
    part = tensor_engine(mesh, cfg["physics"], [(loc[lo:hi], cfg["comps"])], b0=B0, model_type=cfg["mt"])
    part.build_G()
    assert bool((part.G[:, : part.nP] == eng.G[lo:hi, : eng.nP]).all())


def test_octree_below_topography():
    """TreeMesh-style input (hanging nodes, cells of 3 sizes, topography) through the generic planner."""
    topo = lambda x, y: -12.0 + 10.0 * np.exp(-0.5 * ((x / 60.0) ** 2 + (y / 60.0) ** 2))  # noqa: E731
    mesh = synthetic_octree((10, 10, 5), 5.0, levels=3, topo=topo)
    assert mesh.n_cells > 5000 and len(np.unique(np.round(mesh.cell_volumes, 6))) == 3
    assert mesh.cell_centers[:, 2].max() < -2.0 + 5.0    # cells above the topography were dropped
    rng = np.random.default_rng(6)
    xy = rng.uniform(-90, 90, size=(24, 2))
    loc = np.c_[xy, topo(xy[:, 0], xy[:, 1]) + 5.0]
    eng = CudaEngine("gravity", mesh.total_nodes, mesh.cell_nodes, [(loc, ["gz", "gzz"])])
    ora = OracleSimulation("gravity", mesh.total_nodes, mesh.cell_nodes, [(loc, ["gz", "gzz"])])
    eng.build_G()
    assert rowwise_rel(eng.G_to_numpy(), ora.sensitivity_matrix()).max() <= RTOL_NORMWISE
    m = rng.uniform(-0.5, 0.5, eng.nP)
    w = rng.uniform(0.5, 1.5, eng.n_rows)
    assert vec_rel(eng.to_host(eng.forward_dev(eng.to_device(m))), ora.forward(m)) <= RTOL_NORMWISE
    G = ora.sensitivity_matrix()
    assert vec_rel(eng.to_host(eng.gtg_diag_dev(eng.to_device(w))), np.einsum("i,ij,ij->j", w, G, G)) <= RTOL_NORMWISE


def test_physics_sanity_buried_sphere():
    """Independent known answer: gz of a discretised sphere ~ point mass (gravity/analytics.py:6-86 idea)."""
    import scipy.constants as constants

    mesh = TensorMesh([np.full(40, 2.5)] * 3, "CCN")
    centre, radius, drho = np.array([0.0, 0.0, -50.0]), 20.0, 0.5
    inside = np.linalg.norm(mesh.cell_centers - centre, axis=1) < radius
    rho = np.where(inside, drho, 0.0)
    loc = np.array([[0.0, 0.0, 10.0], [30.0, -20.0, 10.0], [80.0, 60.0, 15.0]])
    eng = tensor_engine(mesh, "gravity", [(loc, ["gz"])])
    gz = eng.to_host(eng.forward_dev(eng.to_device(rho)))
    mass = drho * 1e3 * inside.sum() * 2.5**3            # kg (g/cc -> kg/m3)
    d = centre - loc
    expected = constants.G * mass * d[:, 2] / np.linalg.norm(d, axis=1) ** 3 * 1e5   # m/s2 -> mGal
    np.testing.assert_allclose(gz, expected, rtol=2e-3)
