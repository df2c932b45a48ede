"""
CPU: pin the restated choclo formulas (third-party, not in the reference tree)
against the physics they stand for:

* first / second order kernels summed over the 8 vertices of a prism equal
  Gauss-Legendre quadrature of d(1/R) and d2(1/R) over the prism;
* third order kernels equal central finite differences of the second order
  prism sums;
* Laplace: ee + nn + uu = 0 outside the prism;
* the singular branches of safe_log / safe_atan2 take their documented values.
"""
import numpy as np
import pytest

from oracle import oracle

PRISM = (-1.0, 2.0, -3.0, 1.5, -4.0, -1.0)   # west, east, south, north, bottom, top
POINT = np.array([3.3, -2.2, 2.7])


def prism_sum(name, p=POINT, prism=PRISM):
    w, e, s, n, b, t = prism
    total = 0.0
    for i, x in enumerate((w, e)):
        for j, y in enumerate((s, n)):
            for k, z in enumerate((b, t)):
                sign = (-1.0) ** (i + j + k + 1)   # +1 at (east, north, top)
                total += sign * oracle.kernel(name, x - p[0], y - p[1], z - p[2])
    return total


def quadrature(fun, p=POINT, prism=PRISM, order=24):
    xg, wg = np.polynomial.legendre.leggauss(order)
    w, e, s, n, b, t = prism
    X = 0.5 * (e - w) * xg + 0.5 * (e + w)
    Y = 0.5 * (n - s) * xg + 0.5 * (n + s)
    Z = 0.5 * (t - b) * xg + 0.5 * (t + b)
    XX, YY, ZZ = np.meshgrid(X, Y, Z, indexing="ij")
    W = wg[:, None, None] * wg[None, :, None] * wg[None, None, :]
    jac = 0.125 * (e - w) * (n - s) * (t - b)
    dx, dy, dz = XX - p[0], YY - p[1], ZZ - p[2]
    R = np.sqrt(dx**2 + dy**2 + dz**2)
    return jac * np.sum(W * fun(dx, dy, dz, R))


FIRST = {"e": lambda x, y, z, R: x / R**3, "n": lambda x, y, z, R: y / R**3,
         "u": lambda x, y, z, R: z / R**3}
SECOND = {
    "ee": lambda x, y, z, R: 3 * x * x / R**5 - 1 / R**3,
    "nn": lambda x, y, z, R: 3 * y * y / R**5 - 1 / R**3,
    "uu": lambda x, y, z, R: 3 * z * z / R**5 - 1 / R**3,
    "en": lambda x, y, z, R: 3 * x * y / R**5,
    "eu": lambda x, y, z, R: 3 * x * z / R**5,
    "nu": lambda x, y, z, R: 3 * y * z / R**5,
}


@pytest.mark.parametrize("name", sorted(FIRST))
def test_first_order_vs_quadrature(name):
    assert prism_sum(name) == pytest.approx(quadrature(FIRST[name]), rel=1e-12)


@pytest.mark.parametrize("name", sorted(SECOND))
def test_second_order_vs_quadrature(name):
    assert prism_sum(name) == pytest.approx(quadrature(SECOND[name]), rel=1e-11)


def test_laplace_trace():
    assert abs(prism_sum("ee") + prism_sum("nn") + prism_sum("uu")) < 1e-14


def test_kernel_uv_definition():
    assert prism_sum("uv") == pytest.approx(0.5 * (prism_sum("nn") - prism_sum("ee")), rel=1e-13)


THIRD = {  # name -> (second order kernel, axis to differentiate along)
    "eee": ("ee", 0), "nnn": ("nn", 1), "uuu": ("uu", 2),
    "een": ("ee", 1), "eeu": ("ee", 2), "enn": ("nn", 0), "nnu": ("nn", 2),
    "euu": ("uu", 0), "nuu": ("uu", 1), "enu": ("en", 2),
}


@pytest.mark.parametrize("name", sorted(THIRD))
def test_third_order_vs_finite_difference(name):
    second, axis = THIRD[name]
    h = 1e-4
    dp = np.zeros(3)
    dp[axis] = h
    fd = (prism_sum(second, POINT + dp) - prism_sum(second, POINT - dp)) / (2 * h)
    assert prism_sum(name) == pytest.approx(fd, rel=2e-7)


def test_safe_branches():
    lib = oracle._library()
    # safe_atan2
    assert lib.pfo_safe_atan2(1.0, 0.0) == np.pi / 2
    assert lib.pfo_safe_atan2(-1.0, 0.0) == -np.pi / 2
    assert lib.pfo_safe_atan2(0.0, 0.0) == 0.0
    assert lib.pfo_safe_atan2(1.0, -2.0) == np.arctan(-0.5)   # NOT the 4-quadrant atan2
    # safe_log
    assert lib.pfo_safe_log(0.0, 0.0, 0.0, 0.0) == 0.0
    assert lib.pfo_safe_log(-3.0, 0.0, 0.0, 3.0) == -np.log(6.0)
    x, y, z = -2.0, 0.5, 0.25
    r = np.sqrt(x * x + y * y + z * z)
    assert lib.pfo_safe_log(x, y, z, r) == pytest.approx(np.log(x + r), rel=1e-13)
    assert lib.pfo_safe_log(2.0, y, z, np.sqrt(4 + y * y + z * z)) == np.log(2.0 + np.sqrt(4 + y * y + z * z))
    # third order zero branches
    assert oracle.kernel("eee", 0.0, 0.0, -1.0) == 0.0
    assert oracle.kernel("een", 0.0, 0.0, -1.0) == 0.0
    assert oracle.kernel("enu", 0.0, 0.0, 0.0) == 0.0


def test_gravity_point_mass_limit():
    """A small prism far away acts like a point mass: gz ~ G m z / r^3 (mGal, g/cc)."""
    import scipy.constants as constants

    prism = (-0.5, 0.5, -0.5, 0.5, -100.5, -99.5)
    p = np.array([30.0, -20.0, 0.0])
    gz = constants.G * 1e8 * prism_sum("u", p, prism)           # density 1 g/cc
    d = np.array([0.0, 0.0, -100.0]) - p
    expected = constants.G * 1e8 * d[2] / np.linalg.norm(d) ** 3
    assert gz == pytest.approx(expected, rel=1e-4)


@pytest.mark.parametrize("name", ["e", "n", "u", "ee", "nn", "uu", "en", "eu", "nu",
                                  "eee", "nnn", "uuu", "een", "eeu", "enn", "nnu", "euu", "nuu", "enu"])
def test_singular_branches_are_the_continuous_limits(name):
    """
    The branches of safe_log / safe_atan2 / kernel_iii / kernel_iij that fire when the
    observation point is aligned with prism vertices, edges or faces are restated from memory of
    choclo (parity unpinned against the real package).  Physics pins them instead: outside the
    prism every component of the field is continuous, so the prism sum AT an aligned point must
    equal its limit from nearby points.  The limit is taken from symmetric pairs 1e-2 m and
    5e-3 m away with one Richardson step (error O(h^4)); closer points are useless because the
    reference formulas themselves lose all digits there (r + x cancels below 1e-7 m).
    """
    w, e, s, n, b, t = PRISM
    aligned = [
        np.array([w, 0.2, 1.0]),          # above the plane of the west face (e == 0 for 4 vertices)
        np.array([0.3, n, 1.0]),          # plane of the north face
        np.array([w, s, 2.0]),            # straight above a vertical edge (e == n == 0)
        np.array([e, n, 0.5]),            # above another vertical edge
        np.array([4.0, 0.1, t]),          # level with the top face, outside (u == 0)
        np.array([4.0, n, t]),            # on the line of a horizontal edge, outside (n == u == 0)
        np.array([w, -5.0, b]),           # on the line of another edge, outside (e == u == 0)
    ]
    shift = np.array([1.0, -0.7, 0.4])

    def mean_at(p, h):
        return 0.5 * (prism_sum(name, p + h * shift) + prism_sum(name, p - h * shift))

    for p in aligned:
        at = prism_sum(name, p)
        limit = (4.0 * mean_at(p, 5e-3) - mean_at(p, 1e-2)) / 3.0
        assert np.isfinite(at)
        assert at == pytest.approx(limit, abs=5e-7), (name, p)
