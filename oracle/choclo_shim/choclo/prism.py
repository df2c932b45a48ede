"""
choclo.prism kernel functions restated (shim, TEST INFRASTRUCTURE).

Shifted coordinates are (vertex - observation point); ``radius`` is their
Euclidean norm.  Sign conventions: kernel_e = -d/de kernel_pot,
kernel_ee = +d2/de2 kernel_pot, kernel_eee = -d3/de3 kernel_pot.
"""
from numba import jit

from ._utils import _safe_atan2, _safe_log


# ---- first order -----------------------------------------------------------
@jit(nopython=True)
def kernel_e(easting, northing, upward, radius):
    kernel = -(
        northing * _safe_log(upward, easting, northing, radius)
        + upward * _safe_log(northing, easting, upward, radius)
        - easting * _safe_atan2(northing * upward, easting * radius)
    )
    return kernel


@jit(nopython=True)
def kernel_n(easting, northing, upward, radius):
    kernel = -(
        upward * _safe_log(easting, northing, upward, radius)
        + easting * _safe_log(upward, easting, northing, radius)
        - northing * _safe_atan2(upward * easting, northing * radius)
    )
    return kernel


@jit(nopython=True)
def kernel_u(easting, northing, upward, radius):
    kernel = -(
        easting * _safe_log(northing, easting, upward, radius)
        + northing * _safe_log(easting, northing, upward, radius)
        - upward * _safe_atan2(easting * northing, upward * radius)
    )
    return kernel


# ---- second order ----------------------------------------------------------
@jit(nopython=True)
def kernel_ee(easting, northing, upward, radius):
    return -_safe_atan2(northing * upward, easting * radius)


@jit(nopython=True)
def kernel_nn(easting, northing, upward, radius):
    return -_safe_atan2(easting * upward, northing * radius)


@jit(nopython=True)
def kernel_uu(easting, northing, upward, radius):
    return -_safe_atan2(easting * northing, upward * radius)


@jit(nopython=True)
def kernel_en(easting, northing, upward, radius):
    return _safe_log(upward, easting, northing, radius)


@jit(nopython=True)
def kernel_eu(easting, northing, upward, radius):
    return _safe_log(northing, easting, upward, radius)


@jit(nopython=True)
def kernel_nu(easting, northing, upward, radius):
    return _safe_log(easting, northing, upward, radius)


# ---- third order -----------------------------------------------------------
@jit(nopython=True)
def _kernel_iii(x_i, x_j, x_k, radius):
    if (x_i == 0 and x_j == 0) or (x_i == 0 and x_k == 0):
        return 0.0
    x_i_sq, x_j_sq, x_k_sq = x_i**2, x_j**2, x_k**2
    numerator = -x_j * x_k * (2 * x_i_sq + x_j_sq + x_k_sq)
    denominator = (x_i_sq + x_j_sq) * (x_i_sq + x_k_sq) * radius
    return numerator / denominator


@jit(nopython=True)
def _kernel_iij(x_i, x_j, x_k, radius):
    if x_i == 0 and x_j == 0 and x_k < 0:
        return 0.0
    return -x_i / ((x_k + radius) * radius)


@jit(nopython=True)
def kernel_eee(easting, northing, upward, radius):
    return _kernel_iii(easting, northing, upward, radius)


@jit(nopython=True)
def kernel_nnn(easting, northing, upward, radius):
    return _kernel_iii(northing, upward, easting, radius)


@jit(nopython=True)
def kernel_uuu(easting, northing, upward, radius):
    return _kernel_iii(upward, easting, northing, radius)


@jit(nopython=True)
def kernel_een(easting, northing, upward, radius):
    return _kernel_iij(easting, northing, upward, radius)


@jit(nopython=True)
def kernel_eeu(easting, northing, upward, radius):
    return _kernel_iij(easting, upward, northing, radius)


@jit(nopython=True)
def kernel_enn(easting, northing, upward, radius):
    return _kernel_iij(northing, easting, upward, radius)


@jit(nopython=True)
def kernel_nnu(easting, northing, upward, radius):
    return _kernel_iij(northing, upward, easting, radius)


@jit(nopython=True)
def kernel_euu(easting, northing, upward, radius):
    return _kernel_iij(upward, easting, northing, radius)


@jit(nopython=True)
def kernel_nuu(easting, northing, upward, radius):
    return _kernel_iij(upward, northing, easting, radius)


@jit(nopython=True)
def kernel_enu(easting, northing, upward, radius):
    if radius == 0:
        return 0.0
    return -1 / radius


# ---------------------------------------------------------------------------
# Forward functions (choclo/prism/_gravity.py, _magnetic.py), used by the reference's
# equivalent-source layer loops (gravity/_numba/_2d_mesh.py, magnetics/_numba/_2d_mesh.py).
# Restated for observation points OUTSIDE the prism (the only case an equivalent-source layer
# meets: data lie above the sources).  Real choclo additionally returns nan on singular
# edges/vertices and corrects the diagonal components inside the prism: parity UNPINNED there.
# ---------------------------------------------------------------------------
from .constants import GRAVITATIONAL_CONST, VACUUM_MAGNETIC_PERMEABILITY  # noqa: E402
import numpy as np  # noqa: E402


@jit(nopython=True)
def _evaluate_kernel(easting, northing, upward, prism_west, prism_east, prism_south,
                     prism_north, prism_bottom, prism_top, kernel):
    """Signed sum of ``kernel`` over the 8 shifted vertices: east, north, top first."""
    result = 0.0
    for i in range(2):
        if i == 0:
            shift_east = prism_east - easting
        else:
            shift_east = prism_west - easting
        shift_east_sq = shift_east**2
        for j in range(2):
            if j == 0:
                shift_north = prism_north - northing
            else:
                shift_north = prism_south - northing
            shift_north_sq = shift_north**2
            for k in range(2):
                if k == 0:
                    shift_upward = prism_top - upward
                else:
                    shift_upward = prism_bottom - upward
                shift_upward_sq = shift_upward**2
                radius = np.sqrt(shift_east_sq + shift_north_sq + shift_upward_sq)
                result += (-1) ** (i + j + k) * kernel(
                    shift_east, shift_north, shift_upward, radius)
    return result


def _make_gravity(kernel):
    @jit(nopython=True)
    def gravity(easting, northing, upward, prism_west, prism_east, prism_south, prism_north,
                prism_bottom, prism_top, density):
        return (
            GRAVITATIONAL_CONST
            * density
            * _evaluate_kernel(easting, northing, upward, prism_west, prism_east, prism_south,
                               prism_north, prism_bottom, prism_top, kernel)
        )

    return gravity


gravity_e = _make_gravity(kernel_e)
gravity_n = _make_gravity(kernel_n)
gravity_u = _make_gravity(kernel_u)
gravity_ee = _make_gravity(kernel_ee)
gravity_nn = _make_gravity(kernel_nn)
gravity_uu = _make_gravity(kernel_uu)
gravity_en = _make_gravity(kernel_en)
gravity_eu = _make_gravity(kernel_eu)
gravity_nu = _make_gravity(kernel_nu)


def _make_magnetic(kernel_a, kernel_b, kernel_c):
    """One component of B (or of its gradient): per vertex M . (k_a, k_b, k_c), times mu0/4pi."""

    @jit(nopython=True)
    def magnetic(easting, northing, upward, prism_west, prism_east, prism_south, prism_north,
                 prism_bottom, prism_top, magnetization_east, magnetization_north,
                 magnetization_up):
        result = 0.0
        for i in range(2):
            if i == 0:
                shift_east = prism_east - easting
            else:
                shift_east = prism_west - easting
            shift_east_sq = shift_east**2
            for j in range(2):
                if j == 0:
                    shift_north = prism_north - northing
                else:
                    shift_north = prism_south - northing
                shift_north_sq = shift_north**2
                for k in range(2):
                    if k == 0:
                        shift_upward = prism_top - upward
                    else:
                        shift_upward = prism_bottom - upward
                    shift_upward_sq = shift_upward**2
                    radius = np.sqrt(shift_east_sq + shift_north_sq + shift_upward_sq)
                    result += (-1) ** (i + j + k) * (
                        magnetization_east * kernel_a(shift_east, shift_north, shift_upward, radius)
                        + magnetization_north * kernel_b(shift_east, shift_north, shift_upward, radius)
                        + magnetization_up * kernel_c(shift_east, shift_north, shift_upward, radius)
                    )
        return VACUUM_MAGNETIC_PERMEABILITY / 4 / np.pi * result

    return magnetic


magnetic_e = _make_magnetic(kernel_ee, kernel_en, kernel_eu)
magnetic_n = _make_magnetic(kernel_en, kernel_nn, kernel_nu)
magnetic_u = _make_magnetic(kernel_eu, kernel_nu, kernel_uu)
magnetic_ee = _make_magnetic(kernel_eee, kernel_een, kernel_eeu)
magnetic_nn = _make_magnetic(kernel_enn, kernel_nnn, kernel_nnu)
magnetic_uu = _make_magnetic(kernel_euu, kernel_nuu, kernel_uuu)
magnetic_en = _make_magnetic(kernel_een, kernel_enn, kernel_enu)
magnetic_eu = _make_magnetic(kernel_eeu, kernel_enu, kernel_euu)
magnetic_nu = _make_magnetic(kernel_enu, kernel_nnu, kernel_nuu)


@jit(nopython=True)
def magnetic_field(easting, northing, upward, prism_west, prism_east, prism_south, prism_north,
                   prism_bottom, prism_top, magnetization_east, magnetization_north,
                   magnetization_up):
    """(b_e, b_n, b_u) of a uniformly magnetised prism, in T for M in A/m."""
    b_e, b_n, b_u = 0.0, 0.0, 0.0
    for i in range(2):
        if i == 0:
            shift_east = prism_east - easting
        else:
            shift_east = prism_west - easting
        shift_east_sq = shift_east**2
        for j in range(2):
            if j == 0:
                shift_north = prism_north - northing
            else:
                shift_north = prism_south - northing
            shift_north_sq = shift_north**2
            for k in range(2):
                if k == 0:
                    shift_upward = prism_top - upward
                else:
                    shift_upward = prism_bottom - upward
                shift_upward_sq = shift_upward**2
                radius = np.sqrt(shift_east_sq + shift_north_sq + shift_upward_sq)
                k_ee = kernel_ee(shift_east, shift_north, shift_upward, radius)
                k_nn = kernel_nn(shift_east, shift_north, shift_upward, radius)
                k_uu = kernel_uu(shift_east, shift_north, shift_upward, radius)
                k_en = kernel_en(shift_east, shift_north, shift_upward, radius)
                k_eu = kernel_eu(shift_east, shift_north, shift_upward, radius)
                k_nu = kernel_nu(shift_east, shift_north, shift_upward, radius)
                b_e += (-1) ** (i + j + k) * (
                    magnetization_east * k_ee + magnetization_north * k_en + magnetization_up * k_eu)
                b_n += (-1) ** (i + j + k) * (
                    magnetization_east * k_en + magnetization_north * k_nn + magnetization_up * k_nu)
                b_u += (-1) ** (i + j + k) * (
                    magnetization_east * k_eu + magnetization_north * k_nu + magnetization_up * k_uu)
    c_m = VACUUM_MAGNETIC_PERMEABILITY / 4 / np.pi
    return c_m * b_e, c_m * b_n, c_m * b_u
