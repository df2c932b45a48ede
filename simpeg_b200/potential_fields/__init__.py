"""Potential-field simulations with ``engine="cuda"``."""
from . import gravity, magnetics  # noqa: F401
