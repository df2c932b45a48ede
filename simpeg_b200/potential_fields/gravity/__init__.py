"""Gravity: mirrors ``simpeg.potential_fields.gravity`` for the integral hot path."""
from . import receivers, sources, survey, simulation  # noqa: F401
from .receivers import Point  # noqa: F401
from .sources import SourceField  # noqa: F401
from .survey import Survey  # noqa: F401
from .simulation import Simulation3DIntegral, SimulationEquivalentSourceLayer  # noqa: F401
