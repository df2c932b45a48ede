"""Physical constants exposed by choclo.constants (shim)."""
GRAVITATIONAL_CONST = 6.6743e-11
VACUUM_MAGNETIC_PERMEABILITY = 1.25663706212e-6
