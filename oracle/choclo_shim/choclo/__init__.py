"""
Stand-in for the third-party package ``choclo`` (TEST INFRASTRUCTURE).

The reference pins ``choclo>=0.3.0`` (pyproject.toml:54) but does not vendor
it, and it is not installable in this image (no network).  This shim restates
the published closed-form prism kernels of ``choclo.prism`` so that the
reference's own numba loop files can be imported unmodified (see
``oracle/ref_loader.py``).  Formulas: SURVEY.md section 8 rows a1-a5.

Singular branches of ``_safe_log`` / ``_safe_atan2`` are restated from memory
of ``choclo/_utils.py``: parity UNPINNED against real choclo for those.
"""
from . import prism  # noqa: F401
from . import constants  # noqa: F401

__version__ = "0.3.0+shim"
