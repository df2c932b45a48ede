"""
GPU: gravity.Simulation3DIntegral(engine="cuda") against the oracle, mirroring
the reference's tests/pf/test_forward_Grav_Linear.py:423-863 (same mesh,
receivers, model, assertions and tolerances; the oracle plays the part of
engine="choclo").
"""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator

from helpers import RTOL_NORMWISE, rowwise_rel, vec_rel
from sim_fixtures import densities, gravity_survey, make_mapping, oracle_for, small_mesh
from simpeg_b200.potential_fields import gravity

pytestmark = pytest.mark.gpu

ALL = ["gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv"]


def make_sim(components="gz", store="ram", mapping="identity_map", dtype=np.float64, **kw):
    mesh = small_mesh()
    sim = gravity.Simulation3DIntegral(
        mesh, survey=gravity_survey(components), rhoMap=make_mapping(mapping, mesh.n_cells),
        store_sensitivities=store, sensitivity_dtype=dtype, engine="cuda", **kw)
    rho = densities(mesh)
    model = rho if mapping == "identity_map" else np.log(rho)
    return sim, mesh, rho, model


@pytest.mark.parametrize("components", ["gz", ["gx", "gy", "gz"], ALL])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
def test_dpred_matches_oracle(components, store):
    sim, _, rho, model = make_sim(components, store)
    expected = oracle_for(sim).forward(rho)
    assert vec_rel(sim.dpred(model), expected) <= RTOL_NORMWISE


@pytest.mark.parametrize("components", ["gz", ALL])
def test_G_matches_oracle(components):
    sim, *_ = make_sim(components)
    G = sim.G.to_numpy()
    expected = oracle_for(sim).sensitivity_matrix()
    assert G.shape == expected.shape == (sim.survey.nD, sim.nC)
    assert rowwise_rel(G, expected).max() <= RTOL_NORMWISE


@pytest.mark.parametrize("mapping", ["identity_map", "exp_map"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
@pytest.mark.parametrize("transpose", [False, True], ids=["J @ m", "J.T @ v"])
def test_getJ_as_linear_operator(mapping, store, transpose):
    # test_forward_Grav_Linear.py:499-524: an operator for 'forward_only' (and for 'ram' once the
    # matrix is too large for the host: getJ_dense_max_bytes)
    sim, mesh, rho, model = make_sim("gz", store, mapping)
    if store == "ram":
        sim.getJ_dense_max_bytes = 0
    jac = sim.getJ(model)
    assert isinstance(jac, LinearOperator)
    G = oracle_for(sim).sensitivity_matrix()
    deriv = sim.rhoMap.deriv(model)
    if transpose:
        vector = np.random.default_rng(seed=42).uniform(size=sim.survey.nD)
        np.testing.assert_allclose(jac.T @ vector, deriv.T @ (G.T @ vector), rtol=1e-9)
    else:
        np.testing.assert_allclose(jac @ model, G @ (deriv.diagonal() * model), rtol=1e-9)


@pytest.mark.parametrize("mapping", ["identity_map", "exp_map"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_getJ_as_array(mapping, dtype):
    # test_forward_Grav_Linear.py:476-497: a dense ndarray for store_sensitivities="ram";
    # with an identity mapping it is G itself, otherwise G @ mapping.deriv(model)
    sim, mesh, rho, model = make_sim("gz", "ram", mapping, dtype=dtype)
    jac = sim.getJ(model)
    assert isinstance(jac, np.ndarray) and jac.shape == (sim.survey.nD, mesh.n_cells)
    G = oracle_for(sim).sensitivity_matrix()
    expected = G if mapping == "identity_map" else G @ sim.rhoMap.deriv(model)
    tol = RTOL_NORMWISE if dtype is np.float64 else 1e-6
    assert rowwise_rel(np.asarray(jac, dtype=np.float64), np.asarray(expected)).max() <= tol


@pytest.mark.parametrize("mapping", ["identity_map", "exp_map"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
def test_Jvec_Jtvec(mapping, store):
    # test_forward_Grav_Linear.py:565-632, atol = max|expected| * 1e-7
    sim, mesh, rho, model = make_sim(["gz", "gxx"], store, mapping)
    G = oracle_for(sim).sensitivity_matrix()
    deriv = sim.rhoMap.deriv(model)
    rng = np.random.default_rng(seed=42)
    v = rng.uniform(size=mesh.n_cells)
    expected = G @ (deriv @ v)
    np.testing.assert_allclose(sim.Jvec(model, v), expected, atol=np.max(np.abs(expected)) * 1e-7)
    w = rng.uniform(size=sim.survey.nD)
    expected = deriv.T @ (G.T @ w)
    np.testing.assert_allclose(sim.Jtvec(model, w), expected, atol=np.max(np.abs(expected)) * 1e-7)
    assert vec_rel(sim.Jtvec(model, w), expected) <= RTOL_NORMWISE


def test_ram_vs_forward_only():
    # test_forward_Grav_Linear.py:647-680
    a, _, _, model = make_sim(ALL, "ram")
    b, *_ = make_sim(ALL, "forward_only")
    assert vec_rel(a.dpred(model), b.dpred(model)) <= RTOL_NORMWISE
    v = np.random.default_rng(3).uniform(size=a.survey.nD)
    assert vec_rel(a.Jtvec(model, v), b.Jtvec(model, v)) <= RTOL_NORMWISE


@pytest.mark.parametrize("mapping", ["identity_map", "exp_map"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
@pytest.mark.parametrize("weights", [True, False])
def test_getJtJdiag(mapping, store, weights):
    # test_forward_Grav_Linear.py:684-755
    sim, mesh, rho, model = make_sim("gz", store, mapping)
    W = None
    if weights:
        w = np.random.default_rng(seed=42).normal(size=sim.survey.nD)
        W = sp.diags(np.sqrt(w**2))
    G = oracle_for(sim).sensitivity_matrix()
    J = G @ sim.rhoMap.deriv(model).toarray()
    Wd = np.eye(sim.survey.nD) if W is None else W.toarray()
    expected = np.diag(J.T @ Wd.T @ Wd @ J)
    got = sim.getJtJdiag(model, W=W)
    np.testing.assert_allclose(got, expected, atol=np.max(np.abs(expected)) * 1e-8)


def test_getJtJdiag_multicomponent_uses_row_weights():
    """The CUDA engine indexes weights by row for any component count (the
    reference's on-the-fly loop only does for one component)."""
    for store in ("ram", "forward_only"):
        sim, mesh, rho, model = make_sim(["gz", "gzz", "gxy"], store)
        w = np.random.default_rng(5).uniform(0.5, 1.5, size=sim.survey.nD)
        G = oracle_for(sim).sensitivity_matrix()
        expected = np.einsum("i,ij,ij->j", w**2, G, G)
        got = sim.getJtJdiag(model, W=sp.diags(w))
        assert vec_rel(got, expected) <= RTOL_NORMWISE


def test_getJtJdiag_caching():
    # test_forward_Grav_Linear.py:758-795
    sim, mesh, rho, model = make_sim("gz")
    assert not hasattr(sim, "_gtg_diagonal")
    w1 = sp.diags(np.random.default_rng(1).uniform(size=sim.survey.nD))
    d1 = sim.getJtJdiag(model, W=w1)
    cached = sim._gtg_diagonal
    sha = sim._weights_sha256.digest()
    d1b = sim.getJtJdiag(model, W=w1)
    assert sim._gtg_diagonal is cached and sim._weights_sha256.digest() == sha
    np.testing.assert_array_equal(d1, d1b)
    w2 = sp.diags(np.random.default_rng(2).uniform(size=sim.survey.nD))
    sim.getJtJdiag(model, W=w2)
    assert sim._gtg_diagonal is not cached and sim._weights_sha256.digest() != sha


def test_G_dot_m_and_GT_dot_v_operator():
    # test_forward_Grav_Linear.py:826-863
    sim, mesh, rho, model = make_sim("gz", "forward_only")
    G = oracle_for(sim).sensitivity_matrix()
    assert isinstance(sim.G, LinearOperator)
    np.testing.assert_allclose(sim.G @ rho, G @ rho, atol=np.max(np.abs(G @ rho)) * 1e-9)
    v = np.random.default_rng(seed=42).uniform(size=sim.survey.nD)
    np.testing.assert_allclose(sim.G.T @ v, G.T @ v, atol=np.max(np.abs(G.T @ v)) * 1e-9)


def test_float32_sensitivities_default_dtype():
    """reference default sensitivity_dtype=float32 (base.py:86): 1e-6 bar."""
    sim, mesh, rho, model = make_sim(["gz", "gzz"], dtype=np.float32)
    G = oracle_for(sim).sensitivity_matrix()
    assert sim.G.to_numpy().dtype == np.float32
    assert vec_rel(sim.dpred(model), G @ rho) <= 1e-6


def test_inactive_cells():
    mesh = small_mesh()
    active = mesh.cell_centers[:, 2] < -5.0
    sim = gravity.Simulation3DIntegral(
        mesh, survey=gravity_survey(["gz", "gzz"]), rhoMap=make_mapping("identity_map", int(active.sum())),
        active_cells=active, sensitivity_dtype=np.float64, engine="cuda")
    rho = densities(mesh)[active]
    ora = oracle_for(sim)
    assert vec_rel(sim.dpred(rho), ora.forward(rho)) <= RTOL_NORMWISE
    assert rowwise_rel(sim.G.to_numpy(), ora.sensitivity_matrix()).max() <= RTOL_NORMWISE


def test_user_supplied_G_is_uploaded_not_evaluated():
    # LinearSimulation(G=...) / ``sim.G = array`` (simpeg/simulation.py:664-665, 684-688): the matrix
    # handed in is what the products use
    sim, mesh, rho, model = make_sim(["gz", "gzz"], "ram")
    rng = np.random.default_rng(12)
    G = rng.normal(size=(sim.survey.nD, mesh.n_cells))
    other = gravity.Simulation3DIntegral(mesh, survey=sim.survey, rhoMap=sim.rhoMap, G=G,
                                         sensitivity_dtype=np.float64, counter=None, verbose=False)
    v = rng.normal(size=sim.survey.nD)
    assert vec_rel(other.dpred(model), G @ rho) <= 1e-13
    assert vec_rel(other.Jtvec(model, v), G.T @ v) <= 1e-13
    np.testing.assert_allclose(other.getJ(model), G)
    # assigning G afterwards replaces the resident matrix
    sim.dpred(model)
    sim.G = 2.0 * G
    assert vec_rel(sim.dpred(model), 2.0 * (G @ rho)) <= 1e-13
    with pytest.raises(ValueError, match="G must have shape"):
        bad = gravity.Simulation3DIntegral(mesh, survey=sim.survey, rhoMap=sim.rhoMap, G=G[:, :-1])
        bad.dpred(model)
