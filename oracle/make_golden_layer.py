"""
Generate tests/golden/layer/*.npz from the REFERENCE's own equivalent-source layer loops
(TEST INFRASTRUCTURE; run in the build container where /root/reference exists):

    python -m oracle.make_golden_layer

The loops are ``NUMBA_FUNCTIONS_2D[...]`` of gravity/_numba/_2d_mesh.py and
magnetics/_numba/_2d_mesh.py, executed unmodified by oracle/ref_loader.py on the choclo shim and
dispatched like ``SimulationEquivalentSourceLayer`` (gravity/simulation.py:727-939,
magnetics/simulation.py:1235-1653).  Each file holds the inputs (cell bounds, top, bottom,
receivers, components, b0, model, v, w) and G, fwd, gtv, diag_ref as in oracle/make_golden.py.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

from .oracle import OracleLayerSimulation

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden" / "layer"


def tensor_cell_bounds(xs, ys):
    """``mesh.cell_bounds`` of a 2D tensor mesh: (x_min, x_max, y_min, y_max), x fastest."""
    X0, Y0 = np.meshgrid(xs[:-1], ys[:-1], indexing="ij")
    X1, Y1 = np.meshgrid(xs[1:], ys[1:], indexing="ij")
    return np.ascontiguousarray(np.c_[X0.ravel(order="F"), X1.ravel(order="F"),
                                      Y0.ravel(order="F"), Y1.ravel(order="F")])


def quadtree_cell_bounds():
    """A small quadtree: 4x4 base cells of 8 m, the centre 2x2 refined once, one of those twice."""
    cells = []
    for i in range(4):
        for j in range(4):
            x0, y0 = -16.0 + 8 * i, -16.0 + 8 * j
            if i in (1, 2) and j in (1, 2):
                for a in range(2):
                    for b in range(2):
                        xa, ya = x0 + 4 * a, y0 + 4 * b
                        if (i, j, a, b) == (1, 1, 1, 1):
                            for c in range(2):
                                for d in range(2):
                                    cells.append((xa + 2 * c, xa + 2 * c + 2, ya + 2 * d, ya + 2 * d + 2))
                        else:
                            cells.append((xa, xa + 4, ya, ya + 4))
            else:
                cells.append((x0, x0 + 8, y0, y0 + 8))
    return np.ascontiguousarray(np.array(cells, dtype=np.float64))


def cases():
    rng = np.random.default_rng(20261018)
    xs = -20.0 + 5.0 * np.arange(9)
    ys = -17.5 + 5.0 * np.arange(8)
    bounds = tensor_cell_bounds(xs, ys)                      # 8 x 7 cells of 5 m
    bounds2 = tensor_cell_bounds(np.r_[-14.0, -6.0, -2.0, 0.0, 2.0, 6.0, 14.0],
                                 np.r_[-12.0, -6.0, -3.0, 0.0, 6.0, 18.0])
    nC, nC2 = bounds.shape[0], bounds2.shape[0]
    cx, cy = 0.5 * (bounds[:, 0] + bounds[:, 1]), 0.5 * (bounds[:, 2] + bounds[:, 3])
    topo = -2.0 + 1.5 * np.exp(-0.5 * ((cx / 10.0) ** 2 + (cy / 10.0) ** 2))   # draped top
    thick = rng.uniform(2.0, 6.0, nC)
    qb = quadtree_cell_bounds()

    def grid_rx(n_side, half, z, jitter, seed):
        x = np.linspace(-half, half, n_side)
        X, Y = np.meshgrid(x, x)
        rx = np.c_[X.ravel(), Y.ravel(), np.full(X.size, z)]
        if jitter:
            rx = rx + np.random.default_rng(seed).uniform(-jitter, jitter, rx.shape)
        return np.ascontiguousarray(rx)

    rx_grid = grid_rx(5, 20.0, 1.0, 0.0, 0)        # above cell edges / corners (x, y multiples of 5)
    rx_jit = grid_rx(5, 20.0, 2.5, 1.0, 1)
    rx_b = grid_rx(3, 12.0, 4.0, 0.5, 2)
    b0 = np.array([1.0e4, -2.0e4, -4.0e4])
    inc, dec = np.deg2rad(60.0), np.deg2rad((450.0 - 250.0) % 360.0)
    b0_t = 50000.0 * np.array([np.cos(-inc) * np.cos(dec), np.cos(-inc) * np.sin(dec), np.sin(-inc)])

    out = []

    def add(name, physics, bounds_, top, bottom, rlist, b0_=None, model_type="scalar"):
        n = bounds_.shape[0]
        out.append(dict(name=name, physics=physics, bounds=bounds_,
                        top=np.ascontiguousarray(np.broadcast_to(top, (n,)), dtype=np.float64),
                        bottom=np.ascontiguousarray(np.broadcast_to(bottom, (n,)), dtype=np.float64),
                        rlist=rlist, b0=b0_, model_type=model_type))

    all10 = ["gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "guv"]
    add("layer_grav_flat_all10", "gravity", bounds, -1.0, -6.0, [(rx_grid, all10)])
    add("layer_grav_topo_gz", "gravity", bounds, topo, topo - thick, [(rx_jit, ["gz"])])
    add("layer_grav_two_objects", "gravity", bounds2, -0.5, -0.5 - rng.uniform(1.0, 3.0, nC2),
        [(rx_jit, ["gz"]), (rx_b, ["gxx", "gzz", "gx"])])
    active = rng.uniform(size=nC) < 0.6
    add("layer_grav_inactive_gz", "gravity", bounds[active], topo[active], (topo - thick)[active],
        [(rx_grid, ["gz"])])
    add("layer_grav_quadtree_gz", "gravity", qb, -1.0, -3.0, [(rx_jit, ["gz", "gzz"])])
    for mt in ("scalar", "vector"):
        add(f"layer_mag_tmi_flat_{mt}", "magnetics", bounds, -1.0, -6.0, [(rx_grid, ["tmi"])], b0_t, mt)
        add(f"layer_mag_tmi_topo_{mt}", "magnetics", bounds, topo, topo - thick, [(rx_jit, ["tmi"])],
            b0_t, mt)
        add(f"layer_mag_bxyz_{mt}", "magnetics", bounds2, -0.5, -4.0, [(rx_b, ["bx", "by", "bz"])], b0, mt)
        add(f"layer_mag_tmideriv_{mt}", "magnetics", bounds, topo, topo - thick,
            [(rx_grid, ["tmi_x", "tmi_y", "tmi_z"])], b0_t, mt)
        add(f"layer_mag_grad_{mt}", "magnetics", qb, -1.0, -3.0,
            [(rx_jit, ["bxx", "bxy", "bxz", "byy", "byz", "bzz"])], b0, mt)
    add("layer_mag_mixed_scalar", "magnetics", bounds2, -0.5, -4.0,
        [(rx_b, ["tmi", "tmi_z", "bzz", "bz"])], b0, "scalar")
    return out


def run_case(case, backend):
    sim = OracleLayerSimulation(case["physics"], case["bounds"], case["top"], case["bottom"],
                                case["rlist"], b0=case["b0"], model_type=case["model_type"],
                                backend=backend)
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
        bounds=case["bounds"], top=case["top"], bottom=case["bottom"],
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
    d = {k: z[k] for k in ("bounds", "top", "bottom", "model", "v", "w", "G", "fwd", "gtv", "diag_ref")}
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
