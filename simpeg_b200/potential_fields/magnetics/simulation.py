"""
``magnetics.Simulation3DIntegral`` with ``engine="cuda"`` -- mirror of
simpeg/potential_fields/magnetics/simulation.py:117-1232.

Public API kept: constructor keywords (``model_type`` scalar/vector,
``is_amplitude_data``), ``fields/dpred/getJ/Jvec/Jtvec/getJtJdiag``, ``G``,
``nD``, ``tmi_projection``, ``ampDeriv``, ``M``; supported components are the
reference's CHOCLO_SUPPORTED_COMPONENTS (simulation.py:29-44); the vector model
uses three column blocks [x | y | z] (``_numba/_3d_mesh.py:131-160``).
"""
from __future__ import annotations

import hashlib

import numpy as np


from ...utils import dip_azimuth2cartesian, mkvc, sdiag, validate_string, validate_type
from ..base import BaseEquivalentSourceLayerSimulation, BasePFSimulation

CUDA_SUPPORTED_COMPONENTS = {
    "tmi", "bx", "by", "bz", "bxx", "byy", "bzz", "bxy", "bxz", "byz", "tmi_x", "tmi_y", "tmi_z",
}


class Simulation3DIntegral(BasePFSimulation):
    _physics = "magnetics"
    _prop_name = "chi"

    def __init__(self, mesh, chi=None, chiMap=None, model_type="scalar", is_amplitude_data=False,
                 engine="cuda", numba_parallel=True, **kwargs):
        self.model_type = model_type
        super().__init__(mesh, engine=engine, numba_parallel=numba_parallel, **kwargs)
        self._chi = chi
        self.chiMap = chiMap
        self._M = None
        self.is_amplitude_data = is_amplitude_data
        self.modelMap = self.chiMap
        self._warn_n_processes(kwargs, "magnetic")

    # ------------------------------------------------------------ properties
    @property
    def model_type(self):
        return self._model_type

    @model_type.setter
    def model_type(self, value):
        self._model_type = validate_string("model_type", value, ["scalar", "vector"])

    @property
    def is_amplitude_data(self):
        return self._is_amplitude_data

    @is_amplitude_data.setter
    def is_amplitude_data(self, value):
        self._is_amplitude_data = validate_type("is_amplitude_data", value, bool)

    @property
    def chi(self):
        return self._physical_property(self._chi, self.chiMap)

    @chi.setter
    def chi(self, value):
        self._chi = None if value is None else np.asarray(value, dtype=np.float64)

    @property
    def chiDeriv(self):
        return self._property_deriv(self.chiMap)

    @property
    def M(self):
        """magnetics/simulation.py:212-223"""
        if self.model_type == "vector":
            return None
        if getattr(self, "_M", None) is None:
            self._M = np.ones((self.nC, 3)) * self.survey.source_field.b0
        return self._M

    @M.setter
    def M(self, M):
        if self.model_type == "vector":
            self._M = sdiag(mkvc(M) * self.survey.source_field.amplitude)
        else:
            self._M = np.asarray(M).reshape((self.nC, 3))

    @property
    def nD(self):
        """magnetics/simulation.py:281-287"""
        self._nD = self.survey.nD if not self.is_amplitude_data else self.survey.nD // 3
        return self._nD

    @property
    def tmi_projection(self):
        """magnetics/simulation.py:289-298"""
        if getattr(self, "_tmi_projection", None) is None:
            self._tmi_projection = dip_azimuth2cartesian(
                self.survey.source_field.inclination, self.survey.source_field.declination
            ).squeeze()
        return self._tmi_projection

    @property
    def _delete_on_model_update(self):
        """magnetics/simulation.py:822-827"""
        deletes = super()._delete_on_model_update
        if self.is_amplitude_data:
            deletes = deletes + ["_gtg_diagonal", "_ampDeriv", "_ampDeriv_dev"]
        return deletes

    def _n_columns(self):
        return self.nC if self.model_type == "scalar" else 3 * self.nC

    def _engine_kwargs(self):
        for components, _ in self._get_components_and_receivers():
            if not CUDA_SUPPORTED_COMPONENTS.issuperset(components):
                raise NotImplementedError(
                    f"Other components besides {CUDA_SUPPORTED_COMPONENTS} "
                    "aren't implemented yet.")
        return dict(b0=np.asarray(self.survey.source_field.b0, dtype=np.float64),
                    model_type=self.model_type)

    # ------------------------------------------------------------------- API
    def fields(self, model):
        """magnetics/simulation.py:242-258"""
        self.model = model
        if self.store_sensitivities == "forward_only":
            fields = self._forward(self.chi)
        else:
            fields = np.asarray(self.G @ self.chi)
        if self.is_amplitude_data:
            fields = self.compute_amplitude(fields)
        return fields

    def getJ(self, m, f=None):
        """magnetics/simulation.py:300-344"""
        if self.is_amplitude_data:
            raise NotImplementedError(
                "The `getJ` method is not yet implemented to work with `is_amplitude_data`.")
        self.model = m
        return self._jacobian(self.chiDeriv)

    def getJtJdiag(self, m, W=None, f=None):
        """magnetics/simulation.py:346-406"""
        self.model = m
        # one weight per datum: per row, or per location for amplitude data (the
        # reference's default of survey.nD ones overruns its own loop at 446 then)
        weights = (W.diagonal() ** 2 if W is not None
                   else np.ones(self.nD, dtype=np.float64))
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        weights_sha256 = hashlib.sha256(weights)
        use_cached = (hasattr(self, "_gtg_diagonal") and hasattr(self, "_weights_sha256")
                      and self._weights_sha256.digest() == weights_sha256.digest())
        if not use_cached:
            self._gtg_diagonal = self._get_gtg_diagonal(weights)
            self._weights_sha256 = weights_sha256
        diagonal = self._jtj_diagonal(self._gtg_diagonal, self.chiDeriv)
        return diagonal

    def _get_gtg_diagonal(self, weights):
        """magnetics/simulation.py:408-457"""
        if self.store_sensitivities == "forward_only":
            if self.is_amplitude_data:
                raise NotImplementedError(
                    "Computing the diagonal of `G.T @ G` using `'forward_only'` and "
                    "`is_amplitude_data` hasn't been implemented yet.")
            return self._gtg_diagonal_without_building_g(weights)
        dev = self._dev
        self.G
        if not self.is_amplitude_data:
            return self._cols_to_host(dev.colsumsq_dev(self._local_rows_of(weights)))
        # amplitude data: one weight per location (444-456)
        ad = self._amp_deriv_device()
        w_loc = self._local_locations_of(weights)
        return self._cols_to_host(dev.colsumsq_amp_dev(w_loc, ad))

    def Jvec(self, m, v, f=None):
        """magnetics/simulation.py:459-497"""
        self.model = m
        dmu_dm_v = self._deriv_dot(self.chiDeriv, v)
        if not self.is_amplitude_data:
            return np.asarray(self.G @ dmu_dm_v)
        dev = self._dev
        self.G
        jv = dev.gemv_dev(dev.to_device(dmu_dm_v, n=self._n_columns()))
        out = dev.amp_combine_dev(self._amp_deriv_device(), jv)
        return self._locations_to_host(out)

    def Jtvec(self, m, v, f=None):
        """magnetics/simulation.py:499-534"""
        self.model = m
        v = np.asarray(v, dtype=np.float64)
        if not self.is_amplitude_data:
            Jtvec = self.G.T @ v
        else:
            dev = self._dev
            self.G
            spread = dev.amp_spread_dev(self._amp_deriv_device(), self._local_locations_of(v))
            Jtvec = self._cols_to_host(dev.gemv_t_dev(spread))
        return self._deriv_t_dot(self.chiDeriv, Jtvec)

    # ------------------------------------------------------- amplitude helpers
    def _check_amplitude_layout(self):
        for components, _ in self._get_components_and_receivers():
            if list(components) != ["bx", "by", "bz"]:
                raise NotImplementedError(
                    "is_amplitude_data needs receivers with components ['bx', 'by', 'bz']")

    def _amp_deriv_device(self):
        """Normalised local fields (3, n_loc_local) on the device; cached per model."""
        if getattr(self, "_ampDeriv_dev", None) is None:
            self._check_amplitude_layout()
            dev = self._dev
            self.G
            b = dev.gemv_dev(dev.to_device(self.chi, n=self._n_columns()))
            _, self._ampDeriv_dev = dev.amp_normalize_dev(b)
        return self._ampDeriv_dev

    def _local_locations_of(self, per_location):
        """full (n_loc,) host vector -> local locations on the device."""
        dev = self._dev
        n_loc = self.survey.nD // 3
        v = dev.to_device(per_location, n=n_loc)
        if self._coll.world_size == 1:
            return v
        return v[self._local_locations_index()].contiguous()

    def _local_locations_index(self):
        if getattr(self, "_loc_index_t", None) is None:
            rows = self._sharding.local_rows
            self._loc_index_t = self._dev.torch.from_numpy(rows[::3] // 3).to(self._dev.device)
        return self._loc_index_t

    def _locations_to_host(self, local_t):
        """local per-location device vector -> full (n_loc,) numpy vector."""
        dev = self._dev
        torch = dev.torch
        if self._coll.world_size == 1:
            return dev.to_host(local_t)
        # reuse the row all-gather on a 3x interleaved copy (tiny vectors)
        rep = local_t.repeat_interleave(3)
        full = self._coll.all_gather_rows(rep, self._sharding)
        return dev.to_host(full[::3].contiguous())

    @property
    def ampDeriv(self):
        """magnetics/simulation.py:536-544: (3, n_loc) normalised fields, on the host."""
        if getattr(self, "_ampDeriv", None) is None:
            fields = np.asarray(self.G @ self.chi)
            self._ampDeriv = self.normalized_fields(fields)
        return self._ampDeriv

    @classmethod
    def normalized_fields(cls, fields):
        """magnetics/simulation.py:546-555"""
        amp = cls.compute_amplitude(fields.astype(np.float64))
        return fields.reshape((3, -1), order="F") / amp[None, :]

    @classmethod
    def compute_amplitude(cls, b_xyz):
        """magnetics/simulation.py:557-565"""
        return np.linalg.norm(b_xyz.reshape((3, -1), order="F"), axis=0)


class SimulationEquivalentSourceLayer(BaseEquivalentSourceLayerSimulation, Simulation3DIntegral):
    """
    Equivalent-source layer magnetic simulation with ``engine="cuda"`` -- mirror of
    simpeg/potential_fields/magnetics/simulation.py:1235-1653.

    The reference evaluates, per (receiver, cell), the kernels on the 8 vertices of the prism
    (``magnetics/_numba/_2d_mesh.py``; ``choclo.prism.magnetic_field`` / ``magnetic_*`` divided by
    mu_0 in the forward loops, ``1/(4 pi)`` times the kernel sums elsewhere -- the same number).
    Here the prisms go through the device kernels of the 3D simulation; scalar / vector models,
    components and column layout are those of :class:`Simulation3DIntegral`.

    Parameters
    ----------
    mesh : 2D tensor or tree mesh (x, y discretisation of the layer)
    cell_z_top, cell_z_bottom : float or (n_active_cells,) array
    """

    def __init__(self, mesh, cell_z_top, cell_z_bottom, engine="cuda", numba_parallel=True, **kwargs):
        super().__init__(mesh, cell_z_top, cell_z_bottom, engine=engine,
                         numba_parallel=numba_parallel, **kwargs)
