"""
GPU: magnetics.Simulation3DIntegral(engine="cuda") against the oracle,
mirroring tests/pf/test_forward_Mag_Linear.py:915-1754 (fixtures 915-975:
components tmi, [bx,by,bz], [tmi,bx], [tmi_x,tmi_y,tmi_z] x scalar/vector x
identity/exp map) and the self-contained finite-difference test 689-752.
"""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator

from helpers import RTOL_NORMWISE, rowwise_rel, vec_rel
from sim_fixtures import (magnetic_survey, make_mapping, oracle_for, receiver_locations,
                          small_mesh, susceptibilities)
from simpeg_b200.meshes import TensorMesh
from simpeg_b200.potential_fields import magnetics

pytestmark = pytest.mark.gpu

COMPONENTS = ["tmi", ["bx", "by", "bz"], ["tmi", "bx"], ["tmi_x", "tmi_y", "tmi_z"],
              ["bxx", "bxy", "bxz", "byy", "byz", "bzz"]]


def make_sim(components="tmi", model_type="scalar", store="ram", mapping="identity_map",
             dtype=np.float64, two_objects=False, **kw):
    mesh = small_mesh()
    n = mesh.n_cells * (1 if model_type == "scalar" else 3)
    sim = magnetics.Simulation3DIntegral(
        mesh, survey=magnetic_survey(components, two_objects), chiMap=make_mapping(mapping, n),
        model_type=model_type, store_sensitivities=store, sensitivity_dtype=dtype,
        engine="cuda", **kw)
    chi = susceptibilities(mesh, model_type)
    model = chi if mapping == "identity_map" else np.log(np.abs(chi) + 1e-12)
    chi = chi if mapping == "identity_map" else np.exp(model)
    return sim, mesh, chi, model


@pytest.mark.parametrize("components", COMPONENTS, ids=str)
@pytest.mark.parametrize("model_type", ["scalar", "vector"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
def test_dpred_matches_oracle(components, model_type, store):
    sim, _, chi, model = make_sim(components, model_type, store)
    expected = oracle_for(sim).forward(chi)
    assert vec_rel(sim.dpred(model), expected) <= RTOL_NORMWISE


@pytest.mark.parametrize("components", COMPONENTS, ids=str)
@pytest.mark.parametrize("model_type", ["scalar", "vector"])
def test_G_matches_oracle(components, model_type):
    sim, *_ = make_sim(components, model_type)
    G = sim.G.to_numpy()
    expected = oracle_for(sim).sensitivity_matrix()
    assert G.shape == expected.shape
    assert rowwise_rel(G, expected).max() <= RTOL_NORMWISE


def test_two_receiver_objects_mixed_components():
    # fixture with different component lists per object, test_forward_Mag_Linear.py:937-951
    sim, _, chi, model = make_sim(["tmi", "bx"], "vector", two_objects=True)
    ora = oracle_for(sim)
    assert rowwise_rel(sim.G.to_numpy(), ora.sensitivity_matrix()).max() <= RTOL_NORMWISE
    assert vec_rel(sim.dpred(model), ora.forward(chi)) <= RTOL_NORMWISE


@pytest.mark.parametrize("model_type", ["scalar", "vector"])
@pytest.mark.parametrize("mapping", ["identity_map", "exp_map"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
def test_Jvec_Jtvec_getJ(model_type, mapping, store):
    # test_forward_Mag_Linear.py:1106-1354
    sim, mesh, chi, model = make_sim(["tmi", "bx"], model_type, store, mapping)
    G = oracle_for(sim).sensitivity_matrix()
    deriv = sim.chiMap.deriv(model)
    rng = np.random.default_rng(seed=42)
    v = rng.uniform(size=model.size)
    expected = G @ (deriv @ v)
    np.testing.assert_allclose(sim.Jvec(model, v), expected, atol=np.max(np.abs(expected)) * 1e-7)
    w = rng.uniform(size=sim.survey.nD)
    expected_t = deriv.T @ (G.T @ w)
    np.testing.assert_allclose(sim.Jtvec(model, w), expected_t, atol=np.max(np.abs(expected_t)) * 1e-7)
    jac = sim.getJ(model)
    # magnetics/simulation.py:300-344: dense for 'ram', an operator for 'forward_only'
    assert isinstance(jac, np.ndarray if store == "ram" else LinearOperator)
    np.testing.assert_allclose(jac @ v, expected, atol=np.max(np.abs(expected)) * 1e-7)
    np.testing.assert_allclose(jac.T @ w, expected_t, atol=np.max(np.abs(expected_t)) * 1e-7)


@pytest.mark.parametrize("components", COMPONENTS, ids=str)
@pytest.mark.parametrize("model_type", ["scalar", "vector"])
def test_ram_vs_forward_only(components, model_type):
    # test_forward_Mag_Linear.py:1358-1393
    a, _, _, model = make_sim(components, model_type, "ram")
    b, *_ = make_sim(components, model_type, "forward_only")
    assert vec_rel(a.dpred(model), b.dpred(model)) <= RTOL_NORMWISE
    v = np.random.default_rng(3).uniform(size=a.survey.nD)
    assert vec_rel(a.Jtvec(model, v), b.Jtvec(model, v)) <= RTOL_NORMWISE


@pytest.mark.parametrize("model_type", ["scalar", "vector"])
@pytest.mark.parametrize("store", ["ram", "forward_only"])
@pytest.mark.parametrize("components", ["tmi", ["bx", "by", "bz"], ["tmi_x", "tmi_y", "tmi_z"]], ids=str)
def test_getJtJdiag(model_type, store, components):
    # test_forward_Mag_Linear.py:1397-1493, atol = max * 1e-8
    sim, mesh, chi, model = make_sim(components, model_type, store, "exp_map")
    w = np.random.default_rng(seed=42).uniform(0.5, 1.5, size=sim.survey.nD)
    W = sp.diags(w)
    G = oracle_for(sim).sensitivity_matrix()
    J = G @ sim.chiMap.deriv(model).toarray()
    expected = np.diag(J.T @ W.toarray().T @ W.toarray() @ J)
    got = sim.getJtJdiag(model, W=W)
    np.testing.assert_allclose(got, expected, atol=np.max(np.abs(expected)) * 1e-8)


@pytest.mark.parametrize("model_type", ["scalar", "vector"])
def test_amplitude_data(model_type):
    # test_forward_Mag_Linear.py:625-685 and 1543-1754: hand-derived amplitude algebra
    sim, mesh, chi, model = make_sim(["bx", "by", "bz"], model_type, is_amplitude_data=True)
    G = oracle_for(sim).sensitivity_matrix()
    b = (G @ chi).reshape((3, -1), order="F")
    amp = np.linalg.norm(b, axis=0)
    assert vec_rel(sim.dpred(model), amp) <= RTOL_NORMWISE
    amp_deriv = b / amp[None, :]
    np.testing.assert_allclose(sim.ampDeriv, amp_deriv, atol=1e-9)
    rng = np.random.default_rng(seed=42)
    v = rng.uniform(size=model.size)
    jv = (G @ v).reshape((-1, 3)).T
    expected = (amp_deriv * jv).sum(axis=0)
    np.testing.assert_allclose(sim.Jvec(model, v), expected, atol=np.max(np.abs(expected)) * 1e-7)
    w = rng.uniform(size=sim.nD)
    expected_t = G.T @ (amp_deriv * w).T.reshape(-1)
    np.testing.assert_allclose(sim.Jtvec(model, w), expected_t, atol=np.max(np.abs(expected_t)) * 1e-7)
    # magnetics/simulation.py:444-456
    weights = rng.uniform(0.5, 1.5, size=sim.nD)
    rows = amp_deriv[0][:, None] * G[::3] + amp_deriv[1][:, None] * G[1::3] + amp_deriv[2][:, None] * G[2::3]
    expected_d = np.einsum("i,ij,ij->j", weights**2, rows, rows)
    got = sim.getJtJdiag(model, W=sp.diags(weights))
    assert vec_rel(got, expected_d) <= RTOL_NORMWISE
    # cache is dropped on a model update for amplitude data (simulation.py:822-827)
    assert hasattr(sim, "_gtg_diagonal")
    sim.model = model * 1.1
    assert not hasattr(sim, "_gtg_diagonal") and getattr(sim, "_ampDeriv_dev", None) is None


def test_amplitude_forward_only_diag_not_implemented():
    sim, mesh, chi, model = make_sim(["bx", "by", "bz"], "scalar", "forward_only",
                                     is_amplitude_data=True)
    with pytest.raises(NotImplementedError, match="is_amplitude_data"):
        sim.getJtJdiag(model)


def test_tmi_derivatives_vs_finite_differences():
    """test_forward_Mag_Linear.py:689-752: tmi_x/y/z against central finite
    differences of tmi (delta = 1e-6 there; fp64 G)."""
    mesh = TensorMesh([[(5.0, 6)], [(5.0, 6)], [(5.0, 6)]], "CCN")
    chi = 1e-10 * np.ones(mesh.n_cells)
    chi[np.linalg.norm(mesh.cell_centers - np.array([0.0, 0.0, -15.0]), axis=1) < 8.0] = 0.2
    loc = receiver_locations() * 0.5 + np.array([0.3, -0.2, 0.0])

    def tmi_at(shift):
        rx = magnetics.receivers.Point(loc + shift, components="tmi")
        src = magnetics.sources.UniformBackgroundField([rx], 50000, 60, 250)
        sim = magnetics.Simulation3DIntegral(mesh, survey=magnetics.survey.Survey(src),
                                             chi=chi, store_sensitivities="forward_only")
        return sim.dpred(chi)

    rx = magnetics.receivers.Point(loc, components=["tmi_x", "tmi_y", "tmi_z"])
    src = magnetics.sources.UniformBackgroundField([rx], 50000, 60, 250)
    sim = magnetics.Simulation3DIntegral(mesh, survey=magnetics.survey.Survey(src), chi=chi,
                                         sensitivity_dtype=np.float64)
    d = sim.dpred(chi).reshape(-1, 3)
    delta = 1e-3
    for axis in range(3):
        shift = np.zeros(3)
        shift[axis] = delta
        fd = (tmi_at(shift) - tmi_at(-shift)) / (2 * delta)
        np.testing.assert_allclose(d[:, axis], fd, rtol=2e-5, atol=np.max(np.abs(fd)) * 1e-6)


def test_unsupported_component_set():
    # magnetics/simulation.py:862-866
    class Rx:
        components = ["tmi", "bogus"]
        locations = receiver_locations()

    src = magnetics.sources.UniformBackgroundField([Rx()], 50000, 60, 250)
    sim = magnetics.Simulation3DIntegral(small_mesh(), survey=magnetics.survey.Survey(src),
                                         chi=np.zeros(64))
    with pytest.raises(NotImplementedError, match="aren't implemented yet"):
        sim.dpred(np.zeros(64))


def test_jtvec_in_column_slabs_with_overlapped_d2h(monkeypatch):
    """The host-vector Jtvec runs the column kernel in slabs and fetches finished slabs on a copy
    stream while the next one computes (production sizes only by default): force it on a small
    problem and compare with the oracle, fp64 and fp32 storage."""
    from simpeg_b200.potential_fields import base

    monkeypatch.setattr(base, "JTVEC_SLABS", 3)
    monkeypatch.setattr(base, "JTVEC_SLAB_ALIGN", 16)
    monkeypatch.setattr(base, "JTVEC_SLAB_MIN", 32)
    for dtype, tol in ((np.float64, RTOL_NORMWISE), (np.float32, 1e-6)):
        sim, mesh, chi, model = make_sim(["tmi", "bx"], "vector", "ram", "identity_map", dtype=dtype)
        assert len(base.column_slabs(sim._n_columns(), 3)) == 3
        G = oracle_for(sim).sensitivity_matrix()
        w = np.random.default_rng(8).uniform(size=sim.survey.nD)
        assert vec_rel(sim.Jtvec(model, w), G.T @ w) <= tol
        assert vec_rel(sim.Jtvec(model, 2.0 * w), 2.0 * (G.T @ w)) <= tol     # buffers are reused safely
