"""
Airborne-height fixture from the REFERENCE's own numba loops (TEST INFRASTRUCTURE; run in the
build container where /root/reference exists):

    python -m oracle.make_golden_altitude

Geometry: TensorMesh 16 x 16 x 8 cells of 10 m, receivers 200 / 500 / 1000 m above the top
(three off-grid positions per height).  Cases: gravity gz, gravity gzz, magnetic TMI (vector
model).  At these heights the vertex sum cancels 5-7 digits, so the reference's fp64 result
carries a visible rounding floor.  The fixture stores, per case, the sha256
of the reference's G (bit-level pin for the C port, which then stands in for the reference at
test time) and the reference's own deviation from the extended-precision evaluation of the same
algorithm (oracle/longdouble_oracle.py) -- the yardstick tests/test_altitude_parity.py holds the
CUDA engine to.  Output: tests/golden/altitude.json (a few hundred bytes).
"""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np
import scipy.constants as constants

from . import longdouble_oracle as ld
from .oracle import OracleSimulation, tensor_mesh_arrays

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "altitude.json"
HEIGHTS = (200.0, 500.0, 1000.0)
B0 = (50000.0 * np.array([0.3, -0.25, -0.92]) / np.linalg.norm([0.3, -0.25, -0.92])).tolist()


def geometry():
    nodes, cell_nodes = tensor_mesh_arrays([10.0] * 16, [10.0] * 16, [10.0] * 8, (-80.0, -80.0, -80.0))
    xy = np.array([[3.3, -2.2], [41.7, 57.1], [-66.4, 12.9]])
    rx = np.concatenate([np.c_[xy, np.full(3, h)] for h in HEIGHTS])
    return nodes, cell_nodes, np.ascontiguousarray(rx)


CASES = {
    "grav_gz": ("gravity", ["gz"], "scalar"),
    "grav_gzz": ("gravity", ["gzz"], "scalar"),
    "mag_tmi_vector": ("magnetics", ["tmi"], "vector"),
}


def truth(name, nodes, cell_nodes, rx):
    physics, comps, model_type = CASES[name]
    if physics == "gravity":
        kernel, factor = {"gz": ("u", 1e8), "gzz": ("uu", 1e12)}[comps[0]]
        return ld.gravity_rows(nodes, cell_nodes, rx, kernel, constants.G * factor)
    return ld.tmi_rows(nodes, cell_nodes, rx, B0, scalar=(model_type == "scalar"))


def main():
    nodes, cell_nodes, rx = geometry()
    record = {"heights": list(HEIGHTS), "receivers_per_height": 3, "cases": {}}
    for name, (physics, comps, model_type) in CASES.items():
        G = {}
        for backend in ("ref", "c"):
            G[backend] = OracleSimulation(physics, nodes, cell_nodes, [(rx, comps)], b0=np.array(B0),
                                          model_type=model_type, backend=backend).sensitivity_matrix()
        assert np.array_equal(G["ref"], G["c"]), f"{name}: the C port differs from the reference"
        dev = ld.rowwise_deviation(G["ref"], truth(name, nodes, cell_nodes, rx))
        record["cases"][name] = {
            "sha256_reference_G": hashlib.sha256(np.ascontiguousarray(G["ref"]).tobytes()).hexdigest(),
            "reference_vs_extended_precision_by_height": [float(dev[3 * i: 3 * i + 3].max()) for i in range(3)],
        }
        print(name, record["cases"][name])
    OUT.write_text(json.dumps(record, indent=1) + "\n")


if __name__ == "__main__":
    main()
