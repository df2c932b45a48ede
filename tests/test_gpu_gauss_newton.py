"""
GPU: device-resident Gauss-Newton CG against a numpy restatement of the
reference's host loop -- ``ProjectedGNCG.findSearchDirection`` (simpeg/optimization.py:1625-1711)
with ``H v = 2 J^T W^T W J v + beta R v`` (simpeg/data_misfit.py:341-343) -- run on the golden
``G`` of the reference's own numba loops.

Bars: one Hessian-vector product 1e-9 norm-wise (same as Jvec / Jtvec); the CG step after k
iterations 1e-7 norm-wise (CG amplifies the 1e-15 differences of the summation order by the
condition number of the preconditioned Hessian; observed ~1e-12).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import RTOL_NORMWISE, load_golden, vec_rel

pytestmark = pytest.mark.gpu


def reference_search_direction(H, g, xc, approx_hinv, lower, upper, maxiter, rtol, atol_,
                               step_active_set=True, scale=1e-2):
    """optimization.py:1625-1711, numpy, H a dense matrix."""
    active = np.zeros(xc.shape, dtype=bool)
    active |= xc <= lower
    active |= xc >= upper
    inactive = ~active
    step = np.zeros(g.size)
    r = inactive * (-g)
    p = approx_hinv * r
    sold = np.dot(r, p)
    count = 0
    r_norm0 = np.linalg.norm(r)
    atol = max(rtol * r_norm0, atol_)
    r_norm = r_norm0
    while r_norm > atol and count < maxiter:
        count += 1
        q = inactive * (H @ p)
        alpha = sold / np.dot(p, q)
        step += alpha * p
        r -= alpha * q
        r_norm = np.linalg.norm(r)
        h = approx_hinv * r
        snew = np.dot(r, h)
        p = h + (snew / sold) * p
        sold = snew
    bound = ((xc <= lower) & (g >= 0)) | ((xc >= upper) & (g <= 0))
    active_not_bound = active & (~bound)
    if step_active_set and np.any(active_not_bound):
        rhs_a = active_not_bound * -g
        step += scale * (rhs_a * max(abs(step)) / max(abs(rhs_a)))
    step[bound] = 0
    return step, count, r_norm


def _simulation(case, mapping):
    """Simulation3DIntegral on the golden (nodes, cell_nodes) through a mesh stand-in."""
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity, magnetics
    from layer_helpers import _FixedField

    class GoldenMesh:
        dim = 3
        n_cells = case["cell_nodes"].shape[0]
        nodes = case["nodes"]
        cell_nodes = case["cell_nodes"]

    nC = GoldenMesh.n_cells
    if case["physics"] == "gravity":
        rxs = [gravity.receivers.Point(loc, components=c) for loc, c in case["rlist"]]
        survey = gravity.survey.Survey(gravity.sources.SourceField(receiver_list=rxs))
        mp = maps.IdentityMap(nP=nC) if mapping == "identity" else maps.ExpMap(nP=nC)
        return gravity.Simulation3DIntegral(GoldenMesh, survey=survey, rhoMap=mp,
                                            sensitivity_dtype=np.float64)
    rxs = [magnetics.receivers.Point(loc, components=c) for loc, c in case["rlist"]]
    src = _FixedField(receiver_list=rxs, amplitude=1.0, inclination=0.0, declination=0.0)
    src._b0 = np.asarray(case["b0"], dtype=np.float64)
    survey = magnetics.survey.Survey(src)
    nP = nC * (3 if case["model_type"] == "vector" else 1)
    mp = maps.IdentityMap(nP=nP) if mapping == "identity" else maps.ExpMap(nP=nP)
    return magnetics.Simulation3DIntegral(GoldenMesh, survey=survey, chiMap=mp,
                                          model_type=case["model_type"], sensitivity_dtype=np.float64)


def _smoothness(n, rng):
    """A sparse SPD regularisation Hessian: identity + 1-D second differences, random weights."""
    D = sp.diags([-np.ones(n - 1), np.ones(n)], [1, 0], shape=(n, n), format="csr")
    w = sp.diags(rng.uniform(0.5, 2.0, n))
    return (sp.identity(n) * 1e-2 + D.T @ w @ D).tocsr()


CASES = [("grav_gz_grid", "identity", "sparse"), ("grav_all10", "exp", "diag"),
         ("mag_tmi_vector", "identity", "sparse"), ("mag_tmi_bx_scalar", "exp", None),
         ("grav_gz_inactive", "identity", None)]


@pytest.mark.parametrize("name,mapping,reg", CASES)
def test_hessian_vector_and_search_direction(name, mapping, reg):
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

    case = load_golden(name)
    rng = np.random.default_rng(11)
    sim = _simulation(case, mapping)
    G = case["G"]
    nD, nP = G.shape
    wd = rng.uniform(0.5, 2.0, nD)
    scale = 1.0 / np.abs(G).max()
    wd = wd * scale                                   # data weights ~ 1 / standard deviation
    m = rng.uniform(-0.3, 0.3, nP)
    beta = 0.7
    R = None if reg is None else (_smoothness(nP, rng) if reg == "sparse" else rng.uniform(0.1, 1.0, nP))
    d = np.ones(nP) if mapping == "identity" else np.exp(m)
    J = G * d[None, :]
    H = 2.0 * J.T @ (wd[:, None] ** 2 * J)
    if R is not None:
        H = H + beta * (R.toarray() if sp.issparse(R) else np.diag(R))

    lower, upper = -0.25, 0.28
    solver = DeviceGaussNewtonCG(sim, data_weights=wd, reg_hessian=R, beta=beta, lower=lower,
                                 upper=upper, cg_maxiter=8, cg_rtol=1e-6, cg_atol=0.0)
    xc = np.clip(m, lower, upper)                     # some variables sit on their bounds
    solver.set_model(xc)
    d = np.ones(nP) if mapping == "identity" else np.exp(xc)
    J = G * d[None, :]
    H = 2.0 * J.T @ (wd[:, None] ** 2 * J)
    if R is not None:
        H = H + beta * (R.toarray() if sp.issparse(R) else np.diag(R))

    v = rng.uniform(-1.0, 1.0, nP)
    assert vec_rel(solver.hessian_vector(v), H @ v) <= RTOL_NORMWISE

    g = rng.uniform(-1.0, 1.0, nP)
    # UpdatePreconditioner: 1 / (JtJdiag + beta * diag R)
    pc = solver.jacobi_preconditioner()
    diagA = np.einsum("i,ij,ij->j", wd ** 2, J, J) + (0.0 if R is None else
                                                     beta * (R.diagonal() if sp.issparse(R) else R))
    assert vec_rel(pc, 1.0 / diagA) <= RTOL_NORMWISE
    step = solver.find_search_direction(g)
    ref, count, r_norm = reference_search_direction(H, g, xc, 1.0 / diagA, lower, upper, 8, 1e-6, 0.0)
    assert solver.cg_count == count
    assert vec_rel(step, ref) <= 1e-7
    assert solver.cg_abs_resid == pytest.approx(r_norm, rel=1e-6)
    assert np.all(step[(xc <= lower) & (g >= 0)] == 0)


def test_unpreconditioned_unbounded_converges_to_the_normal_equations():
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG

    case = load_golden("grav_gz_jitter")
    rng = np.random.default_rng(5)
    sim = _simulation(case, "identity")
    G = case["G"] / np.abs(case["G"]).max()
    nD, nP = G.shape
    solver = DeviceGaussNewtonCG(sim, data_weights=np.full(nD, 1.0 / np.abs(case["G"]).max()),
                                 reg_hessian=np.ones(nP), beta=1e-3, cg_maxiter=400, cg_rtol=1e-10)
    solver.set_model(np.zeros(nP))
    g = rng.uniform(-1.0, 1.0, nP)
    step = solver.find_search_direction(g)
    H = 2.0 * G.T @ G + 1e-3 * np.eye(nP)
    assert vec_rel(H @ step, -g) <= 1e-8
    assert solver.cg_rel_resid <= 1e-10


def test_requires_stored_sensitivities():
    from simpeg_b200.gauss_newton import DeviceGaussNewtonCG
    from sim_fixtures import gravity_survey, small_mesh
    from simpeg_b200 import maps
    from simpeg_b200.potential_fields import gravity

    mesh = small_mesh()
    sim = gravity.Simulation3DIntegral(mesh, survey=gravity_survey(), rhoMap=maps.IdentityMap(nP=mesh.n_cells),
                                       store_sensitivities="forward_only")
    with pytest.raises(ValueError, match="store_sensitivities='ram'"):
        DeviceGaussNewtonCG(sim)
