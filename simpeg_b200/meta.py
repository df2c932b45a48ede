"""
Tiled simulations on top of the ``engine="cuda"`` potential-field simulations -- the two
containers of ``simpeg/meta/simulation.py`` that the reference's tiled inversions are built
from:

* :class:`MetaSimulation` (meta/simulation.py:24-326): every tile has its OWN receivers and a
  mapping from the global model to the tile's mesh (``maps.TileMap`` in
  ``examples/02-gravity/plot_inv_grav_tiled.py:155-168``); data are concatenated tile after tile.
* :class:`SumMetaSimulation` (meta/simulation.py:329-424): every tile sees ALL receivers and a
  part of the model; data are summed.

The tiles are ordinary ``Simulation3DIntegral`` objects, so each one builds / streams its own
block of G on the GPU; the containers only move (nD,) / (n_model,) host vectors and apply the
mappings' sparse derivatives, exactly like the reference.  A tile simulation whose ``rhoMap`` /
``chiMap`` already is the tile mapping (the reference example's construction) needs no container
mapping: pass ``maps.IdentityMap``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class MetaSimulation:
    """
    Parameters
    ----------
    simulations : list of simulations (``dpred/Jvec/Jtvec/getJtJdiag``, ``survey.nD``, ``model``)
    mappings : list of mappings (``mapping * m``, ``mapping.deriv(m)``, ``shape``), one per
        simulation, all with the same input length
    """

    def __init__(self, simulations, mappings):
        self.simulations = simulations
        self.mappings = mappings
        self._model = None
        self._jtjdiag = None
        self._data_offsets = np.cumsum(np.r_[0, [sim.survey.nD for sim in self.simulations]])

    @property
    def survey(self):
        """Stand-in for the ``BaseSurvey`` the reference gives itself (meta/simulation.py:99-104)."""
        import types

        return types.SimpleNamespace(nD=self.nD, vnD=np.diff(self._data_offsets))

    # -- validation (meta/simulation.py:107-160) ----------------------------------------
    @property
    def simulations(self):
        return self._simulations

    @simulations.setter
    def simulations(self, value):
        value = list(value)
        if len(set(map(id, value))) != len(value):
            raise ValueError("The 'simulations' list must be unique.")
        self._simulations = value

    @property
    def mappings(self):
        return self._mappings

    @mappings.setter
    def mappings(self, value):
        value = list(value)
        if len(value) != len(self.simulations):
            raise ValueError("Must provide the same number of mappings and simulations.")
        model_len = value[0].shape[1]
        for i, (mapping, sim) in enumerate(zip(value, self.simulations)):
            if mapping.shape[1] != model_len:
                raise ValueError("All mappings must have the same input length")
            sim_map = getattr(sim, "modelMap", None)
            sim_in = getattr(sim_map, "shape", (None, None))[1] if sim_map is not None else None
            if sim_in not in (None, "*") and mapping.shape[0] not in (None, "*") and sim_in != mapping.shape[0]:
                raise ValueError(
                    f"Simulation and mapping at index {i} inconsistent. "
                    f"Simulation mapping shape {sim_in} incompatible with "
                    f"input mapping shape {mapping.shape[0]}.")
        self._mappings = value

    @property
    def nD(self):
        return int(self._data_offsets[-1])

    # -- model (meta/simulation.py:175-192): pushed to every tile through its mapping ----
    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, value):
        if value is not None:
            value = np.asarray(value, dtype=np.float64).reshape(-1)
        changed = value is None or self._model is None or self._model.shape != value.shape \
            or not np.array_equal(self._model, value)
        self._model = None if value is None else value.copy()
        if changed:
            self._jtjdiag = None
            for mapping, sim in zip(self.mappings, self.simulations):
                sim.model = None if value is None else mapping * self._model

    # -- API -----------------------------------------------------------------------------
    def fields(self, m):
        """One field object (= predicted data of the linear tile) per simulation."""
        self.model = m
        return [sim.fields(sim.model) for sim in self.simulations]

    def dpred(self, m=None, f=None):
        if f is None:
            f = self.fields(self.model if m is None else m)
        return np.concatenate([np.asarray(sim.dpred(m=sim.model, f=field)).reshape(-1)
                               for sim, field in zip(self.simulations, f)])

    def Jvec(self, m, v, f=None):
        self.model = m
        out = []
        for mapping, sim in zip(self.mappings, self.simulations):
            out.append(np.asarray(sim.Jvec(sim.model, mapping.deriv(self.model) @ v)).reshape(-1))
        return np.concatenate(out)

    def Jtvec(self, m, v, f=None):
        self.model = m
        v = np.asarray(v, dtype=np.float64).reshape(-1)
        jt_vec = 0
        for i, (mapping, sim) in enumerate(zip(self.mappings, self.simulations)):
            local = v[self._data_offsets[i]: self._data_offsets[i + 1]]
            jt_vec = jt_vec + mapping.deriv(self.model).T @ sim.Jtvec(sim.model, local)
        return np.asarray(jt_vec)

    def getJtJdiag(self, m, W=None, f=None):
        self.model = m
        if self._jtjdiag is None:
            if W is None:
                W = np.ones(self.nD)
            else:
                W = W.diagonal()
            jtj_diag = 0.0
            for i, (mapping, sim) in enumerate(zip(self.mappings, self.simulations)):
                sim_w = sp.diags(W[self._data_offsets[i]: self._data_offsets[i + 1]])
                sim_jtj = sp.diags(np.sqrt(sim.getJtJdiag(sim.model, sim_w)))
                m_deriv = sp.csr_matrix(mapping.deriv(self.model))
                jtj_diag = jtj_diag + np.asarray((sim_jtj @ m_deriv).power(2).sum(axis=0)).flatten()
            self._jtjdiag = jtj_diag
        return self._jtjdiag


class SumMetaSimulation(MetaSimulation):
    """Tiles over the MODEL: every simulation predicts all the data from its part of the model and
    the predictions are summed (meta/simulation.py:329-424)."""

    def __init__(self, simulations, mappings):
        super().__init__(simulations, mappings)
        nD = self.simulations[0].survey.nD
        for sim in self.simulations:
            if sim.survey.nD != nD:
                raise ValueError("All simulations must have the same number of data.")

    @property
    def nD(self):
        return int(self.simulations[0].survey.nD)

    def dpred(self, m=None, f=None):
        if f is None:
            f = self.fields(self.model if m is None else m)
        d_pred = 0
        for sim, field in zip(self.simulations, f):
            d_pred = d_pred + np.asarray(sim.dpred(m=sim.model, f=field))
        return d_pred

    def Jvec(self, m, v, f=None):
        self.model = m
        j_vec = 0
        for mapping, sim in zip(self.mappings, self.simulations):
            j_vec = j_vec + np.asarray(sim.Jvec(sim.model, mapping.deriv(self.model) @ v))
        return j_vec

    def Jtvec(self, m, v, f=None):
        self.model = m
        jt_vec = 0
        for mapping, sim in zip(self.mappings, self.simulations):
            jt_vec = jt_vec + mapping.deriv(self.model).T @ sim.Jtvec(sim.model, v)
        return np.asarray(jt_vec)

    def getJtJdiag(self, m, W=None, f=None):
        self.model = m
        if self._jtjdiag is None:
            jtj_diag = 0.0
            for mapping, sim in zip(self.mappings, self.simulations):
                sim_jtj = sp.diags(np.sqrt(sim.getJtJdiag(sim.model, W)))
                m_deriv = sp.csr_matrix(mapping.deriv(self.model))
                jtj_diag = jtj_diag + np.asarray((sim_jtj @ m_deriv).power(2).sum(axis=0)).flatten()
            self._jtjdiag = jtj_diag
        return self._jtjdiag
