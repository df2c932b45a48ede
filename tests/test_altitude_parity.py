"""
Parity margin at airborne receiver heights.

Receivers 200 / 500 / 1000 m above 10 m cells: the signed 8-vertex sum cancels 5-7 digits, so the
REFERENCE'S OWN fp64 result sits 3e-11 / 3e-10 / 1.5e-9 (norm-wise per row) away from the exactly
evaluated formulas (oracle/longdouble_oracle.py: the same algorithm in x87 extended precision;
numbers pinned in tests/golden/altitude.json from the reference's unmodified numba loops by
oracle/make_golden_altitude.py).  Any other fp64 evaluation order shares that floor, so the
north-star tolerance "1e-9 against the reference" can only be asked where the reference itself is
better than 1e-9 -- up to 500 m over 10 m cells here -- and above that the meaningful statement
is the one tested below: the CUDA engine is as close to the exact value as the reference (gravity:
equal; TMI: within 1.55 x, the price of combining K f at the nodes before the vertex sum).
Measured deviation from the reference (profiles/r2_altitude_report.jsonl): TMI 1.5e-11 / 2.6e-10 /
2.4e-9 and gz 1.9e-11 / 1.2e-10 / 5.7e-10 at 200 / 500 / 1000 m, i.e. the 1e-9 bar holds up to
about 750 m (75 cell sizes) above the mesh.

CPU tests: the C port reproduces the reference bit for bit on this geometry (sha256), and the
extended-precision oracle reproduces the stored floor.  GPU test: CUDA vs exact vs port.
"""
import hashlib
import json

import numpy as np
import pytest

from helpers import ROOT, RTOL_NORMWISE, rowwise_rel
from oracle import longdouble_oracle as ld
from oracle.make_golden_altitude import B0, CASES, geometry, truth
from oracle.oracle import OracleSimulation

FIXTURE = json.loads((ROOT / "tests" / "golden" / "altitude.json").read_text())
EXTENDED = np.finfo(np.longdouble).eps < 1e-18      # x87 80-bit long double


def port_G(name, nodes, cell_nodes, rx):
    physics, comps, model_type = CASES[name]
    return OracleSimulation(physics, nodes, cell_nodes, [(rx, comps)], b0=np.array(B0),
                            model_type=model_type, backend="c").sensitivity_matrix()


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_is_the_reference_at_altitude(name):
    nodes, cell_nodes, rx = geometry()
    G = port_G(name, nodes, cell_nodes, rx)
    digest = hashlib.sha256(np.ascontiguousarray(G).tobytes()).hexdigest()
    assert digest == FIXTURE["cases"][name]["sha256_reference_G"]


@pytest.mark.skipif(not EXTENDED, reason="needs x87 80-bit long double")
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_noise_floor_is_reproduced(name):
    nodes, cell_nodes, rx = geometry()
    dev = ld.rowwise_deviation(port_G(name, nodes, cell_nodes, rx), truth(name, nodes, cell_nodes, rx))
    stored = FIXTURE["cases"][name]["reference_vs_extended_precision_by_height"]
    for i, ref in enumerate(stored):
        assert dev[3 * i: 3 * i + 3].max() == pytest.approx(ref, rel=1e-6)
    # the floor grows with height and crosses the 1e-9 bar between 500 m and 1000 m for the
    # first-order (gz) and TMI kernels: the reference cannot be matched to 1e-9 up there by
    # anything but itself
    if name != "grav_gzz":
        assert stored[1] < RTOL_NORMWISE < stored[2]


@pytest.mark.gpu
@pytest.mark.skipif(not EXTENDED, reason="needs x87 80-bit long double")
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_is_no_further_from_exact_than_the_reference(name):
    from simpeg_b200.engine import CudaEngine
    from simpeg_b200.tiling import build_tile_plan_tensor

    nodes, cell_nodes, rx = geometry()
    physics, comps, model_type = CASES[name]
    axes = tuple(-80.0 + 10.0 * np.arange(n + 1) for n in (16, 16, 8))
    axes = (axes[0], axes[1], -80.0 + 10.0 * np.arange(9))
    exact = truth(name, nodes, cell_nodes, rx)
    G_port = port_G(name, nodes, cell_nodes, rx)
    stored = FIXTURE["cases"][name]["reference_vs_extended_precision_by_height"]
    for use_marching in (True, False):          # z-marching kernel and generic tile kernel
        eng = CudaEngine(physics, None, None, [(rx, comps)], b0=np.array(B0), model_type=model_type,
                         sensitivity_dtype=np.float64, plan=build_tile_plan_tensor(axes, None),
                         use_marching=use_marching)
        eng.build_G()
        G = eng.G_to_numpy()
        dev_exact = ld.rowwise_deviation(G, exact)
        dev_ref = rowwise_rel(G, G_port)
        for i, height in enumerate(FIXTURE["heights"]):
            rows = slice(3 * i, 3 * i + 3)
            # (1) no further from the exact value than the reference's own distance, up to the
            #     factor the node-level hoisting of the TMI combination costs (measured on B200,
            #     profiles/r2_altitude_report.jsonl: gz 1.0-1.2 x, gzz 1.0 x, TMI 1.4-1.55 x)
            assert dev_exact[rows].max() <= 2.0 * stored[i], (use_marching, height, dev_exact[rows].max(), stored[i])
            # (2) the 1e-9 bar against the reference holds wherever the reference's own floor
            #     leaves room for it (<= 500 m over 10 m cells)
            if stored[i] < 0.5 * RTOL_NORMWISE:
                assert dev_ref[rows].max() <= RTOL_NORMWISE, (use_marching, height, dev_ref[rows].max())
            # (3) everywhere: the distance to the reference is explained by the two distances to
            #     the exact value (triangle inequality, nothing else is going on)
            assert dev_ref[rows].max() <= 1.05 * (dev_exact[rows].max() + stored[i]) + 1e-15
