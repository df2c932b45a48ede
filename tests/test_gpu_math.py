"""
GPU: accuracy of the branch-free fp64 building blocks of the prism kernels (csrc/pf_math.cuh)
against extended precision (numpy longdouble): division, atan(y / x) with choclo's _safe_atan2
conventions, log(N / D) and _safe_log.  Bars in ulp of the result, stated per function.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def probe(kind, a, b, c=None):
    import torch

    from simpeg_b200 import _lib

    lib = _lib.load_library()
    ta, tb = (torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda() for v in (a, b))
    tc = None if c is None else torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)).cuda()
    out = torch.empty_like(ta)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pf_math_probe(kind, ta.data_ptr(), tb.data_ptr(), None if tc is None else tc.data_ptr(),
                                 out.data_ptr(), ta.numel(), stream), "pf_math_probe")
    return out.cpu().numpy()


def ulp_error(value, exact):
    exact = np.asarray(exact, dtype=np.longdouble)
    ulp = np.spacing(np.abs(exact.astype(np.float64)))
    return np.abs((np.asarray(value, dtype=np.longdouble) - exact) / ulp).astype(np.float64)


def test_division():
    rng = np.random.default_rng(0)
    n = 200_000
    a = rng.uniform(-1.0, 1.0, n) * 10.0 ** rng.uniform(-8, 8, n)
    b = rng.uniform(0.1, 1.0, n) * 10.0 ** rng.uniform(-8, 8, n) * rng.choice([-1.0, 1.0], n)
    q = probe(0, a, b)
    err = ulp_error(q, a.astype(np.longdouble) / b.astype(np.longdouble))
    assert err.max() <= 1.0                       # faithfully rounded
    assert np.mean(q == a / b) > 0.99             # and almost always the IEEE quotient


def test_atan_ratio_accuracy_and_special_values():
    rng = np.random.default_rng(1)
    n = 200_000
    y = rng.uniform(-1.0, 1.0, n) * 10.0 ** rng.uniform(-6, 6, n)
    x = rng.uniform(-1.0, 1.0, n) * 10.0 ** rng.uniform(-6, 6, n)
    r = probe(1, y, x)
    exact = np.arctan(y.astype(np.longdouble) / x.astype(np.longdouble))
    assert ulp_error(r, exact).max() <= 2.5        # observed 2.25 (numpy arctan(y / x): 1.4)
    # choclo _safe_atan2: x == 0 -> +-pi/2 by the sign of y, 0 when y == 0 as well
    y0 = np.array([1.0, -2.0, 0.0, 0.0, 3.0, -3.0])
    x0 = np.array([0.0, 0.0, 0.0, 5.0, -0.0, 1e-300])
    r0 = probe(1, y0, x0)
    assert r0[0] == np.pi / 2 and r0[1] == -np.pi / 2 and r0[2] == 0.0 and r0[3] == 0.0
    assert abs(r0[4]) == np.pi / 2 and r0[5] == pytest.approx(-np.pi / 2, abs=1e-15)


def test_log_ratio_and_safe_log():
    rng = np.random.default_rng(2)
    n = 200_000
    N = rng.uniform(0.5, 2.0, n) * 10.0 ** rng.uniform(-8, 8, n)
    D = rng.uniform(0.5, 2.0, n) * 10.0 ** rng.uniform(-8, 8, n)
    r = probe(2, N, D)
    exact = np.log(N.astype(np.longdouble)) - np.log(D.astype(np.longdouble))
    # relative to the result, except next to N == D where the result is tiny: there the error
    # is bounded by the rounding of the inputs' difference (absolute 2^-53 of the operands)
    err = np.abs(r.astype(np.longdouble) - exact).astype(np.float64)
    assert np.all(err <= 2.0 * np.spacing(np.abs(r)) + 2.3e-16)
    # _safe_log(x, y, z, r) = log(x + r), all branches
    x = np.r_[rng.uniform(-50, 50, 1000), [-7.0, 0.0, 3.0, -1e-3]]
    y = np.r_[rng.uniform(-50, 50, 1000), [0.0, 0.0, 0.0, 40.0]]
    z = np.r_[rng.uniform(-50, 50, 1000), [0.0, 0.0, 4.0, 30.0]]
    got = probe(3, x, y, z)
    xl, yl, zl = (v.astype(np.longdouble) for v in (x, y, z))
    rl = np.sqrt(xl * xl + yl * yl + zl * zl)
    with np.errstate(divide="ignore", invalid="ignore"):
        exact = np.where(xl < 0, np.log((yl * yl + zl * zl) / (rl - xl)), np.log(xl + rl))
    exact[1000] = -np.log(np.longdouble(14.0))       # on the axis: -log(-2x)
    exact[1001] = 0.0                                # r == 0
    assert np.all(np.abs(got - exact.astype(np.float64)) <= 4e-16 * np.maximum(1.0, np.abs(got)))


def test_sqrt_is_the_ieee_root():
    # pf_sqrt = the fast path of sqrt.rn.f64 without its range check: bit-identical to the correctly
    # rounded root on squared distances (0 and normal doubles), which the `r == -x` axis test of
    # _safe_log relies on
    rng = np.random.default_rng(3)
    n = 2_000_000
    x = np.r_[rng.uniform(0.5, 2.0, n) * 10.0 ** rng.uniform(-12, 14, n), [0.0, 1.0, 4.0, 2.0, 1e-200, 1e200]]
    e, nn, u = rng.uniform(-500, 500, (3, 100_000))
    x = np.r_[x, e * e + nn * nn + u * u, (e * e)]      # exact squares: the axis case r == |e|
    got = probe(4, x, x)
    assert np.array_equal(got, np.sqrt(x))
