"""CPU: the tile planner covers every active cell exactly once and reproduces
the (nodes, cell_nodes) geometry bit-for-bit inside every tile."""
import numpy as np
import pytest

from simpeg_b200.meshes import TensorMesh, synthetic_octree
from simpeg_b200.tiling import TILE_CELLS, TILE_NODES, build_tile_plan


def check_plan(nodes, cell_nodes, method="auto"):
    plan = build_tile_plan(nodes, cell_nodes, method)
    cols = plan.tile_cell_col
    valid = cols >= 0
    assert valid.sum() == cell_nodes.shape[0]
    assert np.array_equal(np.sort(cols[valid]), np.arange(cell_nodes.shape[0]))
    tt, ss = np.nonzero(valid)
    local = plan.tile_cell_nodes[tt, ss]
    assert (local < plan.tile_n_nodes[tt][:, None]).all()
    xyz = plan.tile_xyz[tt[:, None], :, local]
    assert np.array_equal(xyz, nodes[cell_nodes[cols[tt, ss]]])
    assert plan.tile_n_nodes.max() <= TILE_NODES
    assert valid.sum(axis=1).max() <= TILE_CELLS
    # columns ascend inside a tile (contiguous G writes)
    for t in range(min(plan.n_tiles, 8)):
        c = cols[t][cols[t] >= 0]
        assert np.all(np.diff(c) > 0)
    return plan


def test_tensor_mesh_bricks():
    mesh = TensorMesh([np.full(40, 5.0), np.full(40, 5.0), np.full(20, 5.0)], "CCN")
    plan = check_plan(mesh.nodes, mesh.cell_nodes)
    assert plan.method == "brick"
    assert plan.n_tiles == 64            # 10 x 10 x 5 bricks of 500 cells
    assert plan.nodes_per_cell < 1.5


def test_rcb_on_tensor_mesh():
    mesh = TensorMesh([np.full(20, 5.0), np.full(17, 5.0), np.full(9, 5.0)], "CCN")
    plan = check_plan(mesh.nodes, mesh.cell_nodes, "rcb")
    assert plan.method == "rcb"


def test_nonuniform_and_small():
    mesh = TensorMesh([[(2.0, 3, -1.3), (2.0, 5), (2.0, 3, 1.3)], [(3.0, 4)], [(1.0, 2)]], "CCN")
    plan = check_plan(mesh.nodes, mesh.cell_nodes)
    assert plan.n_tiles == 1


def test_inactive_cells_compact_nodes():
    mesh = TensorMesh([np.full(12, 1.0)] * 3, "CCN")
    cc = mesh.cell_centers
    active = cc[:, 2] < -2.0 + 0.3 * cc[:, 0]
    cn = mesh.cell_nodes[active]
    uniq, inv = np.unique(cn, return_inverse=True)
    check_plan(mesh.nodes[uniq], inv.reshape(cn.shape))


def test_octree_with_hanging_nodes():
    topo = lambda x, y: 5.0 * np.exp(-0.5 * ((x / 30.0) ** 2 + (y / 30.0) ** 2))  # noqa: E731
    mesh = synthetic_octree((6, 6, 4), 2.5, levels=3, topo=topo)
    assert mesh.n_cells > 2000
    widths = np.unique(np.round(mesh.cell_volumes ** (1 / 3), 6))
    assert len(widths) == 3              # three refinement levels present
    check_plan(mesh.total_nodes, mesh.cell_nodes)
    native = check_plan(mesh.total_nodes, mesh.cell_nodes, "rcb")
    # the native planner (csrc/pf_plan.cu) cuts where a whole number of expected tiles fits on either
    # side, the numpy statement of the bisection at the median: both are valid plans (check_plan), and
    # the native one must not cost more 256-thread rounds (4 per node round, 1 per cell round)
    numpy_plan = check_plan(mesh.total_nodes, mesh.cell_nodes, "rcb_numpy")

    def rounds(plan):
        cells = (plan.tile_cell_col >= 0).sum(axis=1)
        return 4 * np.ceil(plan.tile_n_nodes / 256).sum() + np.ceil(cells / 256).sum()

    assert rounds(native) <= 1.03 * rounds(numpy_plan)


def test_native_planner_fills_tiles_on_a_large_octree():
    """The proportional split of the native planner: on an octree with hanging nodes the tiles hold
    clearly more cells than median bisection leaves them (which costs the same rounds per tile)."""
    topo = lambda x, y: 50.0 * np.exp(-0.5 * ((x / 100.0) ** 2 + (y / 100.0) ** 2)) - 50.0  # noqa: E731
    mesh = synthetic_octree((40, 40, 12), 10.0, levels=3, topo=topo)
    plan = check_plan(mesh.total_nodes, mesh.cell_nodes, "rcb")
    cells = (plan.tile_cell_col >= 0).sum(axis=1)
    assert plan.tile_n_nodes.mean() >= 0.74 * TILE_NODES          # median bisection: ~0.69
    assert cells.mean() >= 340


def _check_tensor_plan(mesh, active=None):
    from simpeg_b200.tiling import build_tile_plan_tensor, tensor_axes_of

    plan = build_tile_plan_tensor(tensor_axes_of(mesh), active)
    cn = mesh.cell_nodes if active is None else mesh.cell_nodes[active]
    cols = plan.tile_cell_col
    valid = cols >= 0
    assert np.array_equal(np.sort(cols[valid]), np.arange(cn.shape[0]))
    tt, ss = np.nonzero(valid)
    local = plan.tile_cell_nodes[tt, ss]
    assert (local < plan.tile_n_nodes[tt][:, None]).all()
    assert np.array_equal(plan.tile_xyz[tt[:, None], :, local], mesh.nodes[cn[cols[tt, ss]]])
    assert plan.tile_n_nodes.max() <= TILE_NODES
    return plan


def test_native_tensor_planner_matches_mesh_geometry():
    """pf_plan_tensor_* (csrc/pf_plan.cu): same geometry as nodes[cell_nodes], bit for bit."""
    mesh = TensorMesh([[(2.0, 3, -1.3), (2.0, 17), (2.0, 3, 1.3)], [(3.0, 13)], [(1.0, 9)]], "CCN")
    plan = _check_tensor_plan(mesh)
    assert plan.method == "tensor" and plan.n_cells == mesh.n_cells
    active = np.nonzero(mesh.cell_centers[:, 2] < -1.0 - 0.1 * np.abs(mesh.cell_centers[:, 0]))[0]
    plan = _check_tensor_plan(mesh, active)
    assert plan.n_cells == active.size
    big = TensorMesh([np.full(40, 5.0), np.full(40, 5.0), np.full(20, 5.0)], "CCN")
    assert _check_tensor_plan(big).n_tiles == 64


def test_empty_and_invalid():
    plan = build_tile_plan(np.zeros((0, 3)), np.zeros((0, 8), dtype=np.int64))
    assert plan.n_tiles == 0
    with pytest.raises(ValueError):
        build_tile_plan(np.zeros((4, 2)), np.zeros((1, 8), dtype=np.int64))
    with pytest.raises(ValueError):
        build_tile_plan(np.zeros((4, 3)), np.full((1, 8), 7, dtype=np.int64))


@pytest.mark.parametrize("shape", [(100, 100, 10), (40, 40, 20), (200, 200, 4), (600, 3, 2), (3, 600, 2),
                                   (151, 7, 3), (255, 2, 1), (256, 1, 1), (1, 1, 1), (130, 97, 5)])
def test_march_footprints_cover_the_mesh_within_the_kernel_limits(shape):
    """Footprints of the z-marching kernel: every (x, y) cell column exactly once, node capacity
    per layer respected, and no footprint wider than the CTA (one node column per thread)."""
    from simpeg_b200.tiling import march_footprints, march_node_rounds

    nx, ny, nz = shape
    for ncap, threads in ((1024, 256), (512, 256), (512, 128)):
        fps = march_footprints(nx, ny, nz, ncap, threads=threads)
        cover = np.zeros((nx, ny), dtype=int)
        for ix0, iy0, w, h, z0, z1 in fps:
            assert (w + 1) * (h + 1) <= ncap and w + 1 <= threads and w >= 1 - (nx == 1) and h >= 1 - (ny == 1)
            assert (z0, z1) == (0, nz)
            cover[ix0:ix0 + w, iy0:iy0 + h] += 1
        assert (cover == 1).all()
        # the node-round count the planner minimises is the kernel's: whole rows per round
        w, h = fps[:, 2].astype(np.int64), fps[:, 3].astype(np.int64)
        assert np.array_equal(march_node_rounds(w, h, threads), -(-(h + 1) // (threads // (w + 1))))


def test_half_size_footprints_only_when_they_cost_no_more():
    """The plan carries the set for 128-thread CTAs (footprints4h) only when it evaluates at most 2 %
    more node slots per layer than the 256-thread set."""
    from simpeg_b200.tiling import build_tile_plan_tensor, march_node_rounds

    for shape in [(100, 100, 10), (200, 200, 10), (40, 40, 20), (130, 97, 5)]:
        plan = build_tile_plan_tensor([np.arange(n + 1, dtype=float) for n in shape], None, with_tiles=False)
        if plan.footprints4h.shape[0] == 0:
            continue
        def slots(fps, t):
            return (march_node_rounds(fps[:, 2].astype(np.int64), fps[:, 3].astype(np.int64), t) * t).sum()
        assert slots(plan.footprints4h, 128) <= 1.02 * slots(plan.footprints4, 256)
        assert (plan.footprints4h[:, 2] + 1 <= 128).all()
        assert ((plan.footprints4h[:, 2] + 1) * (plan.footprints4h[:, 3] + 1) <= 512).all()
