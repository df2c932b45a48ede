"""choclo._utils restated (shim): numerically safe log / arctan."""
import numpy as np
from numba import jit


@jit(nopython=True)
def _safe_atan2(y, x):
    """arctan(y/x) restricted to [-pi/2, pi/2]; +-pi/2 (or 0) when x == 0."""
    if x != 0:
        result = np.arctan(y / x)
    else:
        if y > 0:
            result = np.pi / 2
        elif y < 0:
            result = -np.pi / 2
        else:
            result = 0.0
    return result


@jit(nopython=True)
def _safe_log(x, y, z, r):
    """log(x + r) with r = sqrt(x^2+y^2+z^2), stable for x < 0; 0 when r == 0."""
    if r == 0:
        return 0.0
    if x < 0:
        if r == -x:
            return -np.log(-2 * x)
        result = np.log((y**2 + z**2) / (r - x))
    else:
        result = np.log(x + r)
    return result
