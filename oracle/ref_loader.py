"""
Load the reference's own numba loop files, UNMODIFIED and BY PATH
(TEST INFRASTRUCTURE -- only usable where ``/root/reference`` exists).

``import simpeg`` is impossible in this image (discretize, geoana, pymatsolver
and choclo are absent), but the three files that hold the hot loops only need
numpy, numba and a module called ``choclo``:

    simpeg/potential_fields/_numba_utils.py
    simpeg/potential_fields/gravity/_numba/_3d_mesh.py
    simpeg/potential_fields/magnetics/_numba/_3d_mesh.py

They are executed from where they lie under the reference root, inside fake
parent packages, with ``oracle/choclo_shim`` providing ``choclo``.  Nothing is
copied into this repository.  The result is used to (1) validate the C
restatement in ``oracle/pf_oracle.c`` and (2) generate ``tests/golden/*.npz``
(``oracle/make_golden.py``).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("PF_REFERENCE_ROOT", "/root/reference"))
_SHIM_DIR = Path(__file__).resolve().parent / "choclo_shim"

_loaded = None


def reference_available() -> bool:
    return (REFERENCE_ROOT / "simpeg" / "potential_fields" / "_numba_utils.py").is_file()


def _exec(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, str(path))
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module


def load_reference():
    """
    Returns ``(choclo, gravity_3d, magnetics_3d)`` where the last two are the
    reference modules ``gravity/_numba/_3d_mesh.py`` and
    ``magnetics/_numba/_3d_mesh.py`` (each exposing ``NUMBA_FUNCTIONS_3D``).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise FileNotFoundError(
            f"reference tree not found under {REFERENCE_ROOT} "
            "(it only exists in the build container, never on the GPU box)"
        )
    if str(_SHIM_DIR) not in sys.path:
        sys.path.insert(0, str(_SHIM_DIR))
    import choclo  # the shim

    pf = REFERENCE_ROOT / "simpeg" / "potential_fields"
    for pkg in (
        "refpf",
        "refpf.gravity",
        "refpf.gravity._numba",
        "refpf.magnetics",
        "refpf.magnetics._numba",
    ):
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = []
            sys.modules[pkg] = mod
    _exec("refpf._numba_utils", pf / "_numba_utils.py")
    grav = _exec("refpf.gravity._numba._3d_mesh", pf / "gravity" / "_numba" / "_3d_mesh.py")
    mag = _exec("refpf.magnetics._numba._3d_mesh", pf / "magnetics" / "_numba" / "_3d_mesh.py")
    _loaded = (choclo, grav, mag)
    return _loaded


_loaded_2d = None


def load_reference_2d():
    """
    ``(gravity_2d, magnetics_2d)``: the reference's equivalent-source layer loop files
    ``gravity/_numba/_2d_mesh.py`` and ``magnetics/_numba/_2d_mesh.py`` (each exposing
    ``NUMBA_FUNCTIONS_2D``), executed unmodified on the choclo shim.
    """
    global _loaded_2d
    if _loaded_2d is None:
        load_reference()
        pf = REFERENCE_ROOT / "simpeg" / "potential_fields"
        grav = _exec("refpf.gravity._numba._2d_mesh", pf / "gravity" / "_numba" / "_2d_mesh.py")
        mag = _exec("refpf.magnetics._numba._2d_mesh", pf / "magnetics" / "_numba" / "_2d_mesh.py")
        _loaded_2d = (grav, mag)
    return _loaded_2d


def reference_forward_tables():
    """
    component -> choclo forward function, as gravity/simulation.py:87-98 (with ``gravity_uv``
    of lines 33-70 rebuilt on the shim) and magnetics/simulation.py:104-114.
    """
    from numba import jit

    choclo, _, _ = load_reference()
    p = choclo.prism
    gravity_nn, gravity_ee = p.gravity_nn, p.gravity_ee

    @jit(nopython=True)
    def gravity_uv(easting, northing, upward, prism_west, prism_east, prism_south, prism_north,
                   prism_bottom, prism_top, density):
        result = 0.5 * (
            gravity_nn(easting, northing, upward, prism_west, prism_east, prism_south,
                       prism_north, prism_bottom, prism_top, density)
            - gravity_ee(easting, northing, upward, prism_west, prism_east, prism_south,
                         prism_north, prism_bottom, prism_top, density)
        )
        return result

    gravity = {
        "gx": p.gravity_e, "gy": p.gravity_n, "gz": p.gravity_u,
        "gxx": p.gravity_ee, "gyy": p.gravity_nn, "gzz": p.gravity_uu,
        "gxy": p.gravity_en, "gxz": p.gravity_eu, "gyz": p.gravity_nu,
        "guv": gravity_uv,
    }
    magnetics = {
        "bx": p.magnetic_e, "by": p.magnetic_n, "bz": p.magnetic_u,
        "bxx": p.magnetic_ee, "byy": p.magnetic_nn, "bzz": p.magnetic_uu,
        "bxy": p.magnetic_en, "bxz": p.magnetic_eu, "byz": p.magnetic_nu,
    }
    return gravity, magnetics


def reference_kernel_tables():
    """
    The component -> choclo kernel tables of the reference's simulation.py
    files (gravity/simulation.py:74-85 and 23-30; magnetics/simulation.py:
    45-103), rebuilt on the shim because those files cannot be imported.
    """
    from numba import jit

    choclo, _, _ = load_reference()
    p = choclo.prism

    @jit(nopython=True)
    def kernel_uv(easting, northing, upward, radius):
        return 0.5 * (
            p.kernel_nn(easting, northing, upward, radius)
            - p.kernel_ee(easting, northing, upward, radius)
        )

    gravity = {
        "gx": p.kernel_e, "gy": p.kernel_n, "gz": p.kernel_u,
        "gxx": p.kernel_ee, "gyy": p.kernel_nn, "gzz": p.kernel_uu,
        "gxy": p.kernel_en, "gxz": p.kernel_eu, "gyz": p.kernel_nu,
        "guv": kernel_uv,
    }
    magnetics = {
        "bx": (p.kernel_ee, p.kernel_en, p.kernel_eu),
        "by": (p.kernel_en, p.kernel_nn, p.kernel_nu),
        "bz": (p.kernel_eu, p.kernel_nu, p.kernel_uu),
        "bxx": (p.kernel_eee, p.kernel_een, p.kernel_eeu),
        "byy": (p.kernel_enn, p.kernel_nnn, p.kernel_nnu),
        "bzz": (p.kernel_euu, p.kernel_nuu, p.kernel_uuu),
        "bxy": (p.kernel_een, p.kernel_enn, p.kernel_enu),
        "bxz": (p.kernel_eeu, p.kernel_enu, p.kernel_euu),
        "byz": (p.kernel_enu, p.kernel_nnu, p.kernel_nuu),
        "tmi_x": (p.kernel_eee, p.kernel_enn, p.kernel_euu,
                  p.kernel_een, p.kernel_eeu, p.kernel_enu),
        "tmi_y": (p.kernel_een, p.kernel_nnn, p.kernel_nuu,
                  p.kernel_enn, p.kernel_enu, p.kernel_nnu),
        "tmi_z": (p.kernel_eeu, p.kernel_nnu, p.kernel_uuu,
                  p.kernel_enu, p.kernel_euu, p.kernel_nuu),
    }
    return gravity, magnetics
