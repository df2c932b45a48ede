"""
Minimal model mappings (the two the reference's hot-path tests use:
``maps.IdentityMap`` and ``maps.ExpMap``).  Mappings stay on the host; their
(sparse) derivatives are applied before / after the device calls, exactly as
in the reference (gravity/simulation.py:386-387, 418-419).  Any object with
``__mul__(m)`` and ``deriv(m)`` (e.g. a real ``simpeg.maps`` mapping) works.
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.sparse as sp


def is_identity_deriv(d):
    """True only for the very objects ``IdentityMap.deriv`` hands out.  The flag is an INSTANCE
    attribute of a plain ``csr_matrix``: scipy builds the results of ``2 * d``, ``-d``,
    ``d @ other``, ``d[:k]`` and ``d.copy()`` as new matrices without it, so a scaled or chained
    derivative is never mistaken for the identity."""
    return getattr(d, "is_identity", False) is True


@functools.lru_cache(maxsize=4)
def _identity_deriv(n):
    d = sp.identity(n, format="csr")
    d.is_identity = True
    # shared between calls: make the buffers read-only so that in-place edits fail loudly
    for a in (d.data, d.indices, d.indptr):
        a.flags.writeable = False
    return d


class IdentityMap:
    def __init__(self, mesh=None, nP=None):
        if nP is None and mesh is not None:
            nP = mesh.n_cells
        self.nP = nP

    @property
    def shape(self):
        return (self.nP, self.nP)

    def _transform(self, m):
        return np.asarray(m, dtype=np.float64)

    def __mul__(self, m):
        return self._transform(m)

    def deriv(self, m, v=None):
        if v is not None:
            return np.asarray(v)
        n = self.nP if self.nP is not None else np.asarray(m).size
        return _identity_deriv(int(n))


class ExpMap(IdentityMap):
    def _transform(self, m):
        return np.exp(np.asarray(m, dtype=np.float64))

    def deriv(self, m, v=None):
        d = sp.diags(np.exp(np.asarray(m, dtype=np.float64)), format="csr")
        return d if v is None else d @ v


class LinearMap(IdentityMap):
    """``m -> A m`` for a (sparse) matrix ``A`` (simpeg.maps.LinearMap): the stand-in for the
    mappings of tiled workflows -- ``maps.TileMap`` (global octree -> local mesh, a sparse
    volume-averaging matrix) and ``maps.Wires`` / ``InjectActiveCells`` (selections) all act on
    this path as a fixed sparse matrix with ``deriv(m) = A``."""

    def __init__(self, A):
        self.A = sp.csr_matrix(A)
        self.nP = self.A.shape[1]

    @property
    def shape(self):
        return self.A.shape

    def _transform(self, m):
        return np.asarray(self.A @ np.asarray(m, dtype=np.float64))

    def deriv(self, m, v=None):
        return self.A if v is None else self.A @ v
