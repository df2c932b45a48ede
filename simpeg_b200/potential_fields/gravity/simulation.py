"""
``gravity.Simulation3DIntegral`` with ``engine="cuda"`` -- mirror of
simpeg/potential_fields/gravity/simulation.py:114-724.

Public API kept: constructor keywords, ``fields/dpred/getJ/Jvec/Jtvec/
getJtJdiag`` and the ``G`` property, data ordering (location-major,
component-minor) and units (mGal for gx/gy/gz, Eotvos for the gradient
components with densities in g/cc).  Model mappings stay on the host.
"""
from __future__ import annotations

import hashlib

import numpy as np
from scipy.sparse.linalg import LinearOperator

from ..base import BaseEquivalentSourceLayerSimulation, BasePFSimulation


class Simulation3DIntegral(BasePFSimulation):
    _physics = "gravity"
    _prop_name = "rho"

    def __init__(self, mesh, rho=None, rhoMap=None, engine="cuda", numba_parallel=True, **kwargs):
        super().__init__(mesh, engine=engine, numba_parallel=numba_parallel,
                         **{k: v for k, v in kwargs.items()})
        self._rho = rho
        self.rhoMap = rhoMap
        self.modelMap = self.rhoMap
        # gravity/simulation.py:186-194 (n_processes is meaningless here as well)
        self._warn_n_processes(kwargs, "gravity")

    # props.Invertible("Density") ------------------------------------------------
    @property
    def rho(self):
        return self._physical_property(self._rho, self.rhoMap)

    @rho.setter
    def rho(self, value):
        self._rho = None if value is None else np.asarray(value, dtype=np.float64)

    @property
    def rhoDeriv(self):
        return self._property_deriv(self.rhoMap)

    # API -------------------------------------------------------------------------
    def fields(self, m):
        """gravity/simulation.py:196-221"""
        self.model = m
        if self.store_sensitivities == "forward_only":
            fields = self._forward(self.rho)
        else:
            fields = self.G @ self.rho
        return np.asarray(fields)

    def getJtJdiag(self, m, W=None, f=None):
        """gravity/simulation.py:223-283 (sha256-keyed cache of diag(G.T W.T W G))."""
        self.model = m
        weights = (W.diagonal() ** 2 if W is not None
                   else np.ones(self.survey.nD, dtype=np.float64))
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        weights_sha256 = hashlib.sha256(weights)
        use_cached = (hasattr(self, "_gtg_diagonal") and hasattr(self, "_weights_sha256")
                      and self._weights_sha256.digest() == weights_sha256.digest())
        if not use_cached:
            self._gtg_diagonal = self._get_gtg_diagonal(weights)
            self._weights_sha256 = weights_sha256
        diagonal = self._jtj_diagonal(self._gtg_diagonal, self.rhoDeriv)
        return diagonal

    def _get_gtg_diagonal(self, weights):
        """gravity/simulation.py:285-316"""
        if self.store_sensitivities == "forward_only":
            return self._gtg_diagonal_without_building_g(weights)
        dev = self._dev
        self.G  # make sure the matrix exists
        return self._cols_to_host(dev.colsumsq_dev(self._local_rows_of(weights)))

    def getJ(self, m, f=None):
        """gravity/simulation.py:318-355: ndarray for 'ram' / 'disk' (small enough for the host),
        LinearOperator otherwise (see ``BasePFSimulation._jacobian``)."""
        self.model = m
        return self._jacobian(self.rhoDeriv)

    def Jvec(self, m, v, f=None):
        """gravity/simulation.py:357-387"""
        self.model = m
        dmu_dm_v = self._deriv_dot(self.rhoDeriv, v)
        return np.asarray(self.G @ dmu_dm_v)

    def Jtvec(self, m, v, f=None):
        """gravity/simulation.py:389-419"""
        self.model = m
        Jtvec = self.G.T @ np.asarray(v, dtype=np.float64)
        return self._deriv_t_dot(self.rhoDeriv, Jtvec)


__all__ = ["Simulation3DIntegral", "LinearOperator"]


class SimulationEquivalentSourceLayer(BaseEquivalentSourceLayerSimulation, Simulation3DIntegral):
    """
    Equivalent-source layer gravity simulation with ``engine="cuda"`` -- mirror of
    simpeg/potential_fields/gravity/simulation.py:727-939.

    The reference evaluates ``choclo.prism.gravity_*`` per (receiver, cell) on the 8 vertices of
    every prism (``gravity/_numba/_2d_mesh.py``); here the layer's prisms are handed to the same
    device kernels as the 3D mesh (vertices shared where they coincide), with the same units:
    ``G_newton * 1e8`` (mGal) / ``* 1e12`` (Eotvos) for densities in g/cc.

    Parameters
    ----------
    mesh : 2D tensor or tree mesh (x, y discretisation of the layer)
    cell_z_top, cell_z_bottom : float or (n_active_cells,) array
    """

    def __init__(self, mesh, cell_z_top, cell_z_bottom, engine="cuda", numba_parallel=True, **kwargs):
        super().__init__(mesh, cell_z_top, cell_z_bottom, engine=engine,
                         numba_parallel=numba_parallel, **kwargs)


__all__ += ["SimulationEquivalentSourceLayer"]
