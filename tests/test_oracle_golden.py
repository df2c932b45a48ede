"""
CPU: the C oracle (oracle/pf_oracle.c) against the golden vectors produced by
the reference's own numba loops (oracle/make_golden.py).  G and the forward
sums follow the reference's operation order exactly -> bit-identical; the
threaded column reductions may differ in the last bits.
"""
import numpy as np
import pytest

from helpers import GOLDEN, load_golden, vec_rel
from oracle.oracle import OracleSimulation


def _sim(case, **kw):
    return OracleSimulation(case["physics"], case["nodes"], case["cell_nodes"], case["rlist"],
                            b0=case["b0"], model_type=case["model_type"], backend="c", **kw)


@pytest.mark.parametrize("name", GOLDEN)
def test_c_oracle_matches_reference_golden(name):
    case = load_golden(name)
    sim = _sim(case)
    assert np.array_equal(sim.sensitivity_matrix(), case["G"])
    assert np.array_equal(sim.forward(case["model"]), case["fwd"])
    assert vec_rel(sim.gt_dot_v(case["v"]), case["gtv"]) < 1e-14
    assert vec_rel(sim.diag_gtg(case["w"]), case["diag_ref"]) < 1e-14


@pytest.mark.parametrize("name", ["grav_gz_grid", "mag_tmi_bx_vector"])
def test_c_oracle_serial_equals_parallel(name):
    case = load_golden(name)
    a, b = _sim(case, parallel=True), _sim(case, parallel=False)
    assert np.array_equal(a.sensitivity_matrix(), b.sensitivity_matrix())
    assert vec_rel(a.gt_dot_v(case["v"]), b.gt_dot_v(case["v"])) < 1e-14


def test_c_oracle_float32_storage():
    case = load_golden("grav_gz_grid")
    sim = _sim(case, sensitivity_dtype=np.float32)
    G32 = sim.sensitivity_matrix()
    assert G32.dtype == np.float32
    assert np.array_equal(G32, case["G"].astype(np.float32))


def test_golden_consistency():
    """fixtures are self-consistent: fwd = G m, gtv = G.T v (reference order noise only)."""
    for name in GOLDEN:
        case = load_golden(name)
        assert vec_rel(case["G"] @ case["model"], case["fwd"]) < 1e-12
        assert vec_rel(case["G"].T @ case["v"], case["gtv"]) < 1e-12
