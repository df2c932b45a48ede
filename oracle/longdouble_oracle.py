"""
Extended-precision restatement of the reference's vertex-sum algorithm
(TEST INFRASTRUCTURE -- tests/ only).

Same algorithm as the reference -- choclo's closed-form prism kernels evaluated on every node
(``_evaluate_kernel``, gravity/_numba/_3d_mesh.py:494-498), the signed 8-vertex sum per cell
(``kernels_in_nodes_to_cell``, _numba_utils.py:18-45), the TMI combination
(magnetics/_numba/_3d_mesh.py:344-372) -- but carried out in ``np.longdouble`` (x87 80-bit,
64-bit mantissa) with vectorised numpy.  It answers one question: how far is the REFERENCE'S OWN
fp64 result from the exactly evaluated formulas?  That distance is the noise floor any other fp64
evaluation order (ours) shares; SURVEY.md App. D.4 measured it once, tests/test_altitude_parity.py
pins it for airborne receiver heights.

Only ordinary (non-singular) geometries are supported: receivers strictly outside the mesh and
not aligned with any node in two coordinates.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble
SIGNS = np.array([-1, 1, 1, -1, 1, -1, -1, 1], dtype=LD)      # _numba_utils.py:18-45


def _safe_log(x, y, z, r):
    """choclo._utils._safe_log for r > 0 and no on-axis points."""
    neg = x < 0
    pos_branch = np.log(np.where(neg, LD(1), x + r))
    neg_branch = np.log(np.where(neg, (y * y + z * z) / (r - x), LD(1)))
    return np.where(neg, neg_branch, pos_branch)


def _safe_atan2(y, x):
    if np.any(x == 0):
        raise ValueError("singular geometry (x == 0) is outside the extended-precision oracle")
    return np.arctan(y / x)


def node_kernels(e, n, u):
    """The six second-order kernels and kernel_e/n/u on arrays of shifted coordinates."""
    r = np.sqrt(e * e + n * n + u * u)
    k = {
        "ee": -_safe_atan2(n * u, e * r),
        "nn": -_safe_atan2(e * u, n * r),
        "uu": -_safe_atan2(e * n, u * r),
        "en": _safe_log(u, e, n, r),
        "eu": _safe_log(n, e, u, r),
        "nu": _safe_log(e, n, u, r),
    }
    k["e"] = -(n * k["en"] + u * k["eu"] + e * k["ee"])
    k["n"] = -(u * k["nu"] + e * k["en"] + n * k["nn"])
    k["u"] = -(e * k["eu"] + n * k["nu"] + u * k["uu"])
    return k


def _vertex_sum(values, cell_nodes):
    """values (nN,) -> (nC,): -k0+k1+k2-k3+k4-k5-k6+k7, left to right."""
    v = values[cell_nodes]                       # (nC, 8)
    out = SIGNS[0] * v[:, 0]
    for j in range(1, 8):
        out = out + SIGNS[j] * v[:, j]
    return out


def gravity_rows(nodes, cell_nodes, receivers, kernel, factor):
    """Rows of G (one per receiver) for one gravity kernel name ("u", "uu", ...), longdouble."""
    nodes = np.asarray(nodes, dtype=LD)
    rows = []
    for rx in np.asarray(receivers, dtype=LD):
        k = node_kernels(nodes[:, 0] - rx[0], nodes[:, 1] - rx[1], nodes[:, 2] - rx[2])
        rows.append(LD(factor) * _vertex_sum(k[kernel], cell_nodes))
    return np.stack(rows)


def tmi_rows(nodes, cell_nodes, receivers, b0, scalar):
    """Rows of the TMI sensitivity (magnetics/_numba/_3d_mesh.py:208-372), longdouble:
    (R, nC) for a scalar model, (R, 3 nC) for a vector model."""
    nodes = np.asarray(nodes, dtype=LD)
    b0 = np.asarray(b0, dtype=LD)
    amp = np.sqrt(np.sum(b0 * b0))
    fx, fy, fz = b0 / amp
    cf = LD(1) / LD(4) / LD(np.pi) * amp          # the reference's fp64 pi, promoted
    rows = []
    for rx in np.asarray(receivers, dtype=LD):
        k = node_kernels(nodes[:, 0] - rx[0], nodes[:, 1] - rx[1], nodes[:, 2] - rx[2])
        u = {q: _vertex_sum(k[q], cell_nodes) for q in ("ee", "nn", "uu", "en", "eu", "nu")}
        bx = u["ee"] * fx + u["en"] * fy + u["eu"] * fz
        by = u["en"] * fx + u["nn"] * fy + u["nu"] * fz
        bz = u["eu"] * fx + u["nu"] * fy + u["uu"] * fz
        if scalar:
            rows.append(cf * (bx * fx + by * fy + bz * fz))
        else:
            rows.append(np.concatenate([cf * bx, cf * by, cf * bz]))
    return np.stack(rows)


def rowwise_deviation(A, truth):
    """max_k |A - truth| / max_k |truth| per row, as float64."""
    A = np.asarray(A, dtype=LD)
    d = np.max(np.abs(A - truth), axis=1) / np.max(np.abs(truth), axis=1)
    return np.asarray(d, dtype=np.float64)
