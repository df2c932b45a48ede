"""
GPU parity of the raw engine (C ABI through simpeg_b200.engine) against the
golden vectors generated from the reference's own numba loops
(tests/golden/*.npz, oracle/make_golden.py).
"""
import numpy as np
import pytest

from helpers import GOLDEN, RTOL_NORMWISE, load_golden, rowwise_rel, single_component, vec_rel

pytestmark = pytest.mark.gpu


def _engine(case, dtype=np.float64):
    from simpeg_b200.engine import CudaEngine

    return CudaEngine(case["physics"], case["nodes"], case["cell_nodes"], case["rlist"],
                      b0=case["b0"], model_type=case["model_type"], sensitivity_dtype=dtype)


@pytest.mark.parametrize("name", GOLDEN)
def test_build_G_matches_reference(name):
    case = load_golden(name)
    eng = _engine(case)
    eng.build_G()
    G = eng.G_to_numpy()
    assert G.shape == case["G"].shape
    assert rowwise_rel(G, case["G"]).max() <= RTOL_NORMWISE


@pytest.mark.parametrize("name", GOLDEN)
def test_forward_only_matches_reference(name):
    case = load_golden(name)
    eng = _engine(case)
    fwd = eng.to_host(eng.forward_dev(eng.to_device(case["model"])))
    assert vec_rel(fwd, case["fwd"]) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", GOLDEN)
def test_gt_dot_v_matches_reference(name):
    case = load_golden(name)
    eng = _engine(case)
    gtv = eng.to_host(eng.gt_dot_v_dev(eng.to_device(case["v"])))
    assert vec_rel(gtv, case["gtv"]) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", GOLDEN)
def test_gtg_diag_matches_reference(name):
    case = load_golden(name)
    eng = _engine(case)
    diag = eng.to_host(eng.gtg_diag_dev(eng.to_device(case["w"])))
    # stored-G definition (gravity/simulation.py:313-315) -- always
    expected = np.einsum("i,ij,ij->j", case["w"], case["G"], case["G"])
    assert vec_rel(diag, expected) <= RTOL_NORMWISE
    # the reference's on-the-fly loop indexes weights by receiver: identical
    # to the definition only for one-component receiver objects
    if single_component(case):
        assert vec_rel(diag, case["diag_ref"]) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", GOLDEN)
def test_stored_G_products(name):
    case = load_golden(name)
    eng = _engine(case)
    eng.build_G()
    G = case["G"]
    y = eng.to_host(eng.gemv_dev(eng.to_device(case["model"])))
    assert vec_rel(y, G @ case["model"]) <= RTOL_NORMWISE
    z = eng.to_host(eng.gemv_t_dev(eng.to_device(case["v"])))
    assert vec_rel(z, G.T @ case["v"]) <= RTOL_NORMWISE
    d = eng.to_host(eng.colsumsq_dev(eng.to_device(case["w"])))
    assert vec_rel(d, np.einsum("i,ij,ij->j", case["w"], G, G)) <= RTOL_NORMWISE


@pytest.mark.parametrize("name", ["grav_gz_grid", "mag_tmi_vector"])
def test_float32_storage(name):
    """sensitivity_dtype=float32 (reference default, base.py:86): 1e-6 bar."""
    case = load_golden(name)
    eng = _engine(case, np.float32)
    eng.build_G()
    G = eng.G_to_numpy()
    assert G.dtype == np.float32
    assert rowwise_rel(G, case["G"]).max() <= 1e-6
    y = eng.to_host(eng.gemv_dev(eng.to_device(case["model"])))
    assert vec_rel(y, case["G"] @ case["model"]) <= 1e-6
    z = eng.to_host(eng.gemv_t_dev(eng.to_device(case["v"])))
    assert vec_rel(z, case["G"].T @ case["v"]) <= 1e-6
