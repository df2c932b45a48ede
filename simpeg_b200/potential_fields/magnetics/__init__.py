"""Magnetics: mirrors ``simpeg.potential_fields.magnetics`` for the integral hot path."""
from . import receivers, sources, survey, simulation  # noqa: F401
from .receivers import Point  # noqa: F401
from .sources import SourceField, UniformBackgroundField  # noqa: F401
from .survey import Survey  # noqa: F401
from .simulation import Simulation3DIntegral, SimulationEquivalentSourceLayer  # noqa: F401
