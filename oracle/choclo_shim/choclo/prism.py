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
