"""
Tiled simulations (simpeg_b200/meta.py, mirror of simpeg/meta/simulation.py:24-424).

CPU: the containers' host logic with dense numpy tiles (no GPU).
GPU: real engine="cuda" tiles against the single-mesh oracle -- the construction of
examples/02-gravity/plot_inv_grav_tiled.py:155-168 (one Simulation3DIntegral per tile with a
sparse tile mapping as its rhoMap) -- and the device Gauss-Newton CG with that sparse mapping.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import RTOL_NORMWISE, vec_rel
from simpeg_b200 import maps
from simpeg_b200.meta import MetaSimulation, SumMetaSimulation


class DenseTile:
    """numpy stand-in for a linear tile simulation with an identity model map."""

    def __init__(self, G):
        self.G = np.asarray(G, dtype=np.float64)
        self.model = None
        self.modelMap = maps.IdentityMap(nP=self.G.shape[1])

        class _Survey:
            nD = self.G.shape[0]

        self.survey = _Survey()

    def fields(self, m):
        self.model = m
        return self.G @ m

    def dpred(self, m=None, f=None):
        return f if f is not None else self.fields(m)

    def Jvec(self, m, v, f=None):
        return self.G @ v

    def Jtvec(self, m, v, f=None):
        return self.G.T @ v

    def getJtJdiag(self, m, W=None, f=None):
        w = np.ones(self.G.shape[0]) if W is None else W.diagonal() ** 2
        return np.einsum("i,ij,ij->j", w, self.G, self.G)


def test_sum_meta_over_model_tiles_equals_the_whole():
    rng = np.random.default_rng(0)
    G = rng.normal(size=(9, 20))
    parts = [np.arange(0, 7), np.arange(7, 12), np.arange(12, 20)]
    sims = [DenseTile(G[:, p]) for p in parts]
    mappings = [maps.LinearMap(sp.identity(20, format="csr")[p]) for p in parts]
    meta = SumMetaSimulation(sims, mappings)
    m, v, w = rng.normal(size=20), rng.normal(size=20), rng.normal(size=9)
    wd = rng.uniform(0.5, 1.5, 9)
    assert np.allclose(meta.dpred(m), G @ m)
    assert np.allclose(meta.Jvec(m, v), G @ v)
    assert np.allclose(meta.Jtvec(m, w), G.T @ w)
    assert np.allclose(meta.getJtJdiag(m, W=sp.diags(wd)), np.einsum("i,ij,ij->j", wd**2, G, G))
    with pytest.raises(ValueError, match="same number of data"):
        SumMetaSimulation([DenseTile(G[:5, :7]), DenseTile(G[:, 7:12])], mappings[:2])


def test_meta_over_receiver_tiles_with_averaging_maps():
    rng = np.random.default_rng(1)
    n_model = 16
    tiles, mappings, blocks = [], [], []
    for n_rows, n_local in ((4, 8), (6, 4), (3, 16)):
        G = rng.normal(size=(n_rows, n_local))
        P = sp.random(n_local, n_model, density=0.3, random_state=int(n_rows), format="csr")
        tiles.append(DenseTile(G))
        mappings.append(maps.LinearMap(P))
        blocks.append(G @ P.toarray())
    J = np.vstack(blocks)
    meta = MetaSimulation(tiles, mappings)
    m, v, w = rng.normal(size=n_model), rng.normal(size=n_model), rng.normal(size=J.shape[0])
    assert meta.nD == 13
    assert np.allclose(meta.dpred(m), J @ m)
    assert np.allclose(meta.Jvec(m, v), J @ v)
    assert np.allclose(meta.Jtvec(m, w), J.T @ w)
    # meta/simulation.py:281-292: exact for one-entry-per-column maps, the reference's
    # approximation otherwise -- what is pinned here is the reference's formula itself
    expected = 0.0
    for tile, mapping in zip(tiles, mappings):
        d = sp.diags(np.sqrt(np.einsum("ij,ij->j", tile.G, tile.G)))
        expected = expected + np.asarray((d @ mapping.A).power(2).sum(axis=0)).ravel()
    assert np.allclose(meta.getJtJdiag(m), expected)
    with pytest.raises(ValueError, match="same number of mappings"):
        MetaSimulation(tiles, mappings[:2])
    with pytest.raises(ValueError, match="inconsistent"):
        MetaSimulation(tiles, [mappings[1], mappings[0], mappings[2]])


# ------------------------------------------------------------------------------------- GPU
def _mesh_and_survey():
    from simpeg_b200.meshes import TensorMesh
    from simpeg_b200.potential_fields import gravity

    mesh = TensorMesh([[(5.0, 8)], [(5.0, 6)], [(5.0, 4)]], "CCN")
    x = np.linspace(-18.0, 18.0, 5)
    X, Y = np.meshgrid(x, x)
    loc = np.c_[X.ravel(), Y.ravel(), np.full(X.size, 3.0)] + 0.3
    rx = gravity.receivers.Point(loc, components=["gz", "gzz"])
    return mesh, loc, gravity.survey.Survey(gravity.sources.SourceField([rx]))


@pytest.mark.gpu
def test_sum_meta_of_cuda_tiles_matches_single_mesh_oracle():
    from oracle.oracle import OracleSimulation
    from simpeg_b200.meshes import TensorMesh
    from simpeg_b200.potential_fields import gravity

    mesh, loc, survey = _mesh_and_survey()
    G = OracleSimulation("gravity", mesh.nodes, mesh.cell_nodes, [(loc, ["gz", "gzz"])]).sensitivity_matrix()
    # two tiles over the model: the west and the east half of the mesh, each its own TensorMesh
    nx, ny, nz = mesh.shape_cells
    ix = np.arange(mesh.n_cells) % nx
    sims, mappings = [], []
    for lo, hi in ((0, 3), (3, 8)):
        sub = TensorMesh([np.full(hi - lo, 5.0), np.full(ny, 5.0), np.full(nz, 5.0)],
                         origin=mesh.origin + np.array([5.0 * lo, 0.0, 0.0]))
        cells = np.nonzero((ix >= lo) & (ix < hi))[0]
        sims.append(gravity.Simulation3DIntegral(sub, survey=survey, rhoMap=maps.IdentityMap(nP=sub.n_cells),
                                                 sensitivity_dtype=np.float64))
        mappings.append(maps.LinearMap(sp.identity(mesh.n_cells, format="csr")[cells]))
    meta = SumMetaSimulation(sims, mappings)
    rng = np.random.default_rng(2)
    m, v, w = rng.uniform(-0.5, 0.5, mesh.n_cells), rng.normal(size=mesh.n_cells), rng.normal(size=survey.nD)
    wd = rng.uniform(0.5, 1.5, survey.nD)
    assert vec_rel(meta.dpred(m), G @ m) <= RTOL_NORMWISE
    assert vec_rel(meta.Jvec(m, v), G @ v) <= RTOL_NORMWISE
    assert vec_rel(meta.Jtvec(m, w), G.T @ w) <= RTOL_NORMWISE
    assert vec_rel(meta.getJtJdiag(m, W=sp.diags(wd)), np.einsum("i,ij,ij->j", wd**2, G, G)) <= RTOL_NORMWISE


def _coarse_tile(mesh):
    """Local mesh with cells twice as large and the volume-averaging matrix global -> local
    (what maps.TileMap builds between octrees)."""
    from simpeg_b200.meshes import TensorMesh

    nx, ny, nz = mesh.shape_cells
    local = TensorMesh([np.full(nx // 2, 10.0), np.full(ny // 2, 10.0), np.full(nz // 2, 10.0)], origin=mesh.origin)
    c = np.arange(mesh.n_cells)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    parent = (i // 2) + (nx // 2) * ((j // 2) + (ny // 2) * (k // 2))
    P = sp.csr_matrix((np.full(mesh.n_cells, 0.125), (parent, c)), shape=(local.n_cells, mesh.n_cells))
    return local, P


@pytest.mark.gpu
def test_receiver_tiles_with_tile_maps_and_device_cg():
    from oracle.oracle import OracleSimulation
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG
    from simpeg_b200.potential_fields import gravity
    from test_gpu_gauss_newton import reference_search_direction

    mesh, loc, _ = _mesh_and_survey()
    local, P = _coarse_tile(mesh)
    G_local = OracleSimulation("gravity", local.nodes, local.cell_nodes, [(loc, ["gz"])]).sensitivity_matrix()
    J = G_local @ P.toarray()
    # plot_inv_grav_tiled.py:155-168: one simulation per receiver tile, rhoMap = the tile mapping
    sims = []
    for rows in (slice(0, 10), slice(10, 25)):
        rx = gravity.receivers.Point(loc[rows], components=["gz"])
        survey = gravity.survey.Survey(gravity.sources.SourceField([rx]))
        sims.append(gravity.Simulation3DIntegral(local, survey=survey, rhoMap=maps.LinearMap(P),
                                                 sensitivity_dtype=np.float64))
    meta = MetaSimulation(sims, [maps.IdentityMap(nP=mesh.n_cells)] * 2)
    rng = np.random.default_rng(3)
    m, v, w = rng.uniform(-0.5, 0.5, mesh.n_cells), rng.normal(size=mesh.n_cells), rng.normal(size=25)
    assert vec_rel(meta.dpred(m), J @ m) <= RTOL_NORMWISE
    assert vec_rel(meta.Jvec(m, v), J @ v) <= RTOL_NORMWISE
    assert vec_rel(meta.Jtvec(m, w), J.T @ w) <= RTOL_NORMWISE
    # device CG on one tile: H = 2 P^T G^T W^2 G P + beta I with the sparse P applied on the device
    sim, Jt = sims[1], J[10:25]
    wd = rng.uniform(0.5, 1.5, 15) / np.abs(Jt).max()
    solver = DeviceGaussNewtonCG(sim, data_weights=wd, reg_hessian=np.ones(mesh.n_cells), beta=0.05,
                                 cg_maxiter=8, cg_rtol=1e-10)
    solver.set_model(m)
    H = 2.0 * Jt.T @ ((wd**2)[:, None] * Jt) + 0.05 * np.eye(mesh.n_cells)
    assert vec_rel(solver.hessian_vector(v), H @ v) <= RTOL_NORMWISE
    g = rng.normal(size=mesh.n_cells)
    step, count, _ = reference_search_direction(H, g, m, np.ones(mesh.n_cells), -np.inf, np.inf, 8, 1e-10, 0.0)
    got = solver.find_search_direction(g)
    assert solver.cg_count == count
    assert vec_rel(got, step) <= 1e-7
