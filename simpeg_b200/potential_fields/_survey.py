"""
Survey bookkeeping shared by gravity and magnetics (boundary inputs of the hot
path).  Data are ordered location-major,
component-minor inside each receiver object (gravity/receivers.py:7-12).
"""
from __future__ import annotations

import numpy as np

from ..utils import validate_string


class BasePoint:
    _allowed = ()
    _default = None

    def __init__(self, locations, components=None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        locations = np.atleast_2d(np.asarray(locations, dtype=np.float64))
        if locations.shape[1] != 3:
            raise ValueError("locations must be an (n, 3) array")
        self.locations = locations
        if components is None:
            components = self._default
        if isinstance(components, str):
            components = [components]
        for component in components:
            validate_string("component", component, list(self._allowed))
        self.components = list(components)

    @property
    def nD(self):
        return self.locations.shape[0] * len(self.components)


class BaseSurvey:
    def __init__(self, source_field, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.source_field = source_field

    @property
    def nRx(self):
        return sum(rx.locations.shape[0] for rx in self.source_field.receiver_list)

    @property
    def receiver_locations(self):
        return np.concatenate([rx.locations for rx in self.source_field.receiver_list])

    @property
    def nD(self):
        return sum(rx.nD for rx in self.source_field.receiver_list)

    @property
    def vnD(self):
        return np.array([rx.nD for rx in self.source_field.receiver_list])

    @property
    def components(self):
        comps = []
        for rx in self.source_field.receiver_list:
            comps += rx.components
        return comps
