"""
simpeg_b200 -- B200-native engine for SimPEG's integral-equation potential-field
hot path (G build, forward_only, Jvec / Jtvec / getJtJdiag for gravity and
magnetics ``Simulation3DIntegral``), exposed as ``engine="cuda"``.
"""
from . import maps, meshes, utils  # noqa: F401
from . import potential_fields  # noqa: F401

__version__ = "0.1.0"
