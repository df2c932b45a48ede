"""Shared helpers of the test-suite."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = sorted(p.stem for p in (ROOT / "tests" / "golden").glob("*.npz"))

# parity bar of BASELINE.json: relative 1e-9 in fp64, stated norm-wise
# (SURVEY.md 8(d)): max|delta| / max|reference| per row of G and per vector.
RTOL_NORMWISE = 1e-9


def load_golden(name):
    from oracle.make_golden import load_case

    return load_case(ROOT / "tests" / "golden" / f"{name}.npz")


def rowwise_rel(a, b):
    """max_j |a_ij - b_ij| / max_j |b_ij| for every row i (rows of zeros -> abs)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b), axis=1)
    scale[scale == 0] = 1.0
    return np.max(np.abs(a - b), axis=1) / scale


def vec_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / (scale if scale else 1.0)


def single_component(case):
    return all(len(c) == 1 for _, c in case["rlist"])
