"""
Generate tests/golden/*.npz from the REFERENCE's own numba loops
(TEST INFRASTRUCTURE; run in the build container where /root/reference exists):

    python -m oracle.make_golden

Each file holds the inputs (nodes, cell_nodes, receivers, components, b0,
model, v, w) and the outputs of the reference loop functions
(``NUMBA_FUNCTIONS_3D[...]`` of gravity/_numba/_3d_mesh.py and
magnetics/_numba/_3d_mesh.py, executed unmodified by oracle/ref_loader.py on
the choclo shim) dispatched exactly like the reference's simulation classes:

    G        _sensitivity_matrix()                      (nD, nP) float64
    fwd      _forward(model)                            (nD,)
    gtv      _sensitivity_matrix_transpose_dot_vec(v)   (nP,)
    diag_ref _gtg_diagonal_without_building_g(w)        (nP,)  [weights[i] quirk, SURVEY a11]

The fixtures are small on purpose; they pin the C oracle (CPU tests) and the
CUDA engine (GPU tests) to the reference without needing numba or the
reference tree at test time.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

from .oracle import OracleSimulation, tensor_mesh_arrays

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _receivers(seed, n_side=5, half=20.0, z=1.0, jitter=0.0):
    x = np.linspace(-half, half, n_side)
    X, Y = np.meshgrid(x, x)
    rx = np.c_[X.ravel(), Y.ravel(), np.full(X.size, z)]
    if jitter:
        rx = rx + np.random.default_rng(seed).uniform(-jitter, jitter, rx.shape)
    return np.ascontiguousarray(rx)


def compact_active(nodes, cell_nodes, active):
    """base.py:371-399 _get_active_nodes for a boolean mask."""
    cn = cell_nodes[active]
    unique_nodes, inv = np.unique(cn, return_inverse=True)
    return np.ascontiguousarray(nodes[unique_nodes]), np.ascontiguousarray(
        inv.reshape(cn.shape).astype(np.int64))


def cases():
    rng = np.random.default_rng(20261017)
    # uniform 8x7x6 mesh, h=5, top at z=0; several receivers sit exactly above
    # nodes / cell edges (x, y multiples of 5) -> exercises the safe_* branches.
    nodes, cn = tensor_mesh_arrays([5.0] * 8, [5.0] * 7, [5.0] * 6, (-20.0, -17.5, -30.0))
    # non-uniform mesh
    nodes2, cn2 = tensor_mesh_arrays([8.0, 4.0, 2.0, 2.0, 4.0, 8.0], [6.0, 3.0, 3.0, 6.0, 12.0],
                                     [10.0, 5.0, 2.5, 2.5], (-14.0, -12.0, -20.0))
    rx_grid = _receivers(0)                       # on-grid, z = +1
    rx_jit = _receivers(1, z=2.5, jitter=1.0)     # generic positions
    rx_b = _receivers(2, n_side=3, half=12.0, z=4.0, jitter=0.5)
    b0 = np.array([1.0e4, -2.0e4, -4.0e4])
    # b0 of the reference tests: 50000 nT, I=60, D=250 (test_forward_Mag_Linear.py:214)
    inc, dec = np.deg2rad(60.0), np.deg2rad((450.0 - 250.0) % 360.0)
    b0_t = 50000.0 * np.array([np.cos(-inc) * np.cos(dec), np.cos(-inc) * np.sin(dec), np.sin(-inc)])

    out = []

    def add(name, physics, nodes_, cn_, rlist, b0_=None, model_type="scalar"):
        out.append(dict(name=name, physics=physics, nodes=nodes_, cell_nodes=cn_, rlist=rlist,
                        b0=b0_, model_type=model_type))

    all10 = ["gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv"]
    add("grav_gz_grid", "gravity", nodes, cn, [(rx_grid, ["gz"])])
    add("grav_gz_jitter", "gravity", nodes, cn, [(rx_jit, ["gz"])])
    add("grav_all10", "gravity", nodes, cn, [(rx_grid, all10)])
    add("grav_tensor7_nonuniform", "gravity", nodes2, cn2,
        [(rx_b, ["gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "gz"])])
    add("grav_two_objects", "gravity", nodes, cn, [(rx_jit, ["gz"]), (rx_b, ["gxx", "gzz", "gx"])])
    active = rng.uniform(size=cn.shape[0]) < 0.6
    an, acn = compact_active(nodes, cn, active)
    add("grav_gz_inactive", "gravity", an, acn, [(rx_grid, ["gz"])])
    for mt in ("scalar", "vector"):
        add(f"mag_tmi_{mt}", "magnetics", nodes, cn, [(rx_grid, ["tmi"])], b0_t, mt)
        add(f"mag_bxyz_{mt}", "magnetics", nodes, cn, [(rx_jit, ["bx", "by", "bz"])], b0, mt)
        add(f"mag_tmi_bx_{mt}", "magnetics", nodes2, cn2, [(rx_b, ["tmi", "bx"])], b0_t, mt)
        add(f"mag_tmideriv_{mt}", "magnetics", nodes, cn,
            [(rx_grid, ["tmi_x", "tmi_y", "tmi_z"])], b0_t, mt)
        add(f"mag_grad_{mt}", "magnetics", nodes, cn,
            [(rx_jit, ["bxx", "bxy", "bxz", "byy", "byz", "bzz"])], b0, mt)
    add("mag_tmi_inactive_vector", "magnetics", an, acn, [(rx_jit, ["tmi"])], b0_t, "vector")
    add("mag_mixed_families_scalar", "magnetics", nodes2, cn2,
        [(rx_b, ["tmi", "tmi_z", "bzz", "bz"])], b0, "scalar")
    return out


def run_case(case, backend):
    sim = OracleSimulation(case["physics"], case["nodes"], case["cell_nodes"], case["rlist"],
                           b0=case["b0"], model_type=case["model_type"], backend=backend)
    # deterministic regardless of PYTHONHASHSEED: derive the seed from the name bytes
    rng = np.random.default_rng(int.from_bytes(case["name"].encode(), "little") % (2**32))
    lo, hi = (-0.5, 0.5) if case["physics"] == "gravity" else (0.0, 0.05)
    model = rng.uniform(lo, hi, sim.nP)
    v = rng.uniform(0.0, 1.0, sim.nD)
    w = rng.uniform(0.5, 1.5, sim.nD)
    return dict(model=model, v=v, w=w, G=sim.sensitivity_matrix(), fwd=sim.forward(model),
                gtv=sim.gt_dot_v(v), diag_ref=sim.diag_gtg(w))


def save_case(case, res):
    comps = ["|".join(c) for _, c in case["rlist"]]
    arrays = dict(
        physics=np.array(case["physics"]), model_type=np.array(case["model_type"]),
        nodes=case["nodes"], cell_nodes=case["cell_nodes"],
        n_objects=np.array(len(case["rlist"])), components=np.array(comps),
        b0=np.zeros(3) if case["b0"] is None else case["b0"], **res)
    for i, (rx, _) in enumerate(case["rlist"]):
        arrays[f"receivers_{i}"] = rx
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN_DIR / f"{case['name']}.npz", **arrays)


def load_case(path):
    z = np.load(path)
    rlist = [(z[f"receivers_{i}"], str(z["components"][i]).split("|"))
             for i in range(int(z["n_objects"]))]
    d = {k: z[k] for k in ("nodes", "cell_nodes", "model", "v", "w", "G", "fwd", "gtv", "diag_ref")}
    d.update(name=Path(path).stem, physics=str(z["physics"]), model_type=str(z["model_type"]),
             rlist=rlist, b0=z["b0"] if str(z["physics"]) == "magnetics" else None)
    return d


def main():
    t0 = time.time()
    for case in cases():
        res = run_case(case, backend="ref")
        save_case(case, res)
        print(f"{case['name']:32s} G{res['G'].shape}  {time.time() - t0:7.1f}s", flush=True)


if __name__ == "__main__":
    sys.exit(main())
