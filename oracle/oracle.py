"""
CPU oracle for the potential-field hot path (TEST INFRASTRUCTURE).

``OracleSimulation`` reproduces the per-(receiver object, component) Python
dispatch of the reference's ``Simulation3DIntegral`` private methods

    gravity/simulation.py:541-724   (_forward, _sensitivity_matrix,
                                     _sensitivity_matrix_transpose_dot_vec,
                                     _gtg_diagonal_without_building_g)
    magnetics/simulation.py:829-1232

on top of one of two loop back-ends:

* ``backend="c"``   -- ``oracle/pf_oracle.c`` (plain C + OpenMP restatement; it
  travels to the GPU box and is the CPU baseline that ``bench.py`` times);
* ``backend="ref"`` -- the reference's own numba files loaded by path
  (``oracle/ref_loader.py``; only where ``/root/reference`` exists).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import scipy.constants as constants

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libpf_oracle.so"

# kernel ids of pf_oracle.c
KERNEL_ID = {
    "e": 0, "n": 1, "u": 2,
    "ee": 3, "nn": 4, "uu": 5, "en": 6, "eu": 7, "nu": 8,
    "uv": 9,
    "eee": 10, "nnn": 11, "uuu": 12, "een": 13, "eeu": 14, "enn": 15,
    "nnu": 16, "euu": 17, "nuu": 18, "enu": 19,
}

# gravity/simulation.py:74-85
GRAVITY_KERNELS = {
    "gx": "e", "gy": "n", "gz": "u",
    "gxx": "ee", "gyy": "nn", "gzz": "uu",
    "gxy": "en", "gxz": "eu", "gyz": "nu",
    "guv": "uv",
}
# magnetics/simulation.py:45-103 ("tmi" hard-wired in _3d_mesh.py:337-342)
MAGNETIC_KERNELS = {
    "tmi": ("ee", "nn", "uu", "en", "eu", "nu"),
    "bx": ("ee", "en", "eu"),
    "by": ("en", "nn", "nu"),
    "bz": ("eu", "nu", "uu"),
    "bxx": ("eee", "een", "eeu"),
    "byy": ("enn", "nnn", "nnu"),
    "bzz": ("euu", "nuu", "uuu"),
    "bxy": ("een", "enn", "enu"),
    "bxz": ("eeu", "enu", "euu"),
    "byz": ("enu", "nnu", "nuu"),
    "tmi_x": ("eee", "enn", "euu", "een", "eeu", "enu"),
    "tmi_y": ("een", "nnn", "nuu", "enn", "enu", "nnu"),
    "tmi_z": ("eeu", "nnu", "uuu", "enu", "euu", "nuu"),
}


def gravity_conversion_factor(component: str) -> float:
    """gravity/simulation.py:101-111 _get_conversion_factor"""
    if component in ("gx", "gy", "gz"):
        return 1e8
    if component in ("gxx", "gyy", "gzz", "gxy", "gxz", "gyz", "guv"):
        return 1e12
    raise ValueError(f"Invalid component '{component}'.")


def build_library(force: bool = False) -> Path:
    """Compile oracle/pf_oracle.c -> oracle/libpf_oracle.so (gcc, OpenMP)."""
    src = _HERE / "pf_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "libpf_oracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def _library():
    global _lib
    if _lib is None:
        build_library()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.pfo_kernel.restype = ctypes.c_double
        _lib.pfo_kernel.argtypes = [ctypes.c_int] + [ctypes.c_double] * 4
        _lib.pfo_safe_log.restype = ctypes.c_double
        _lib.pfo_safe_log.argtypes = [ctypes.c_double] * 4
        _lib.pfo_safe_atan2.restype = ctypes.c_double
        _lib.pfo_safe_atan2.argtypes = [ctypes.c_double] * 2
        _lib.pfo_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(_library().pfo_num_threads())


def set_num_threads(n: int) -> None:
    _library().pfo_set_num_threads(ctypes.c_int(int(n)))


def kernel(name: str, e: float, n: float, u: float) -> float:
    """One choclo kernel value from the C restatement (for formula tests)."""
    r = float(np.sqrt(e * e + n * n + u * u))
    return float(_library().pfo_kernel(KERNEL_ID[name], e, n, u, r))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(x):
    return ctypes.c_int64(int(x))


class _CBackend:
    """Loop back-end on pf_oracle.c.  Strided views are passed as (ptr, stride)."""

    def __init__(self, parallel=True):
        self.lib = _library()
        self.parallel = int(bool(parallel))

    @staticmethod
    def _view(arr):
        # arr is a strided 1-D / 2-D view of a C-contiguous parent.
        itemsize = arr.itemsize
        assert arr.strides[0] % itemsize == 0
        if arr.ndim == 2:
            assert arr.strides[1] == itemsize
        return _p(arr), _i64(arr.strides[0] // itemsize)

    def _common(self, receivers, nodes, cell_nodes):
        return (_p(receivers), _i64(receivers.shape[0]), _p(nodes), _i64(nodes.shape[0]),
                _p(cell_nodes), _i64(cell_nodes.shape[0]))

    # gravity ---------------------------------------------------------------
    def grav_sensitivity(self, receivers, nodes, G_view, cell_nodes, kname, factor):
        ptr, stride = self._view(G_view)
        self.lib.pfo_gravity_sensitivity(
            *self._common(receivers, nodes, cell_nodes), ptr, stride,
            ctypes.c_int(int(G_view.dtype == np.float32)),
            ctypes.c_int(KERNEL_ID[kname]), ctypes.c_double(factor), ctypes.c_int(self.parallel))

    def grav_forward(self, receivers, nodes, densities, fields_view, cell_nodes, kname, factor):
        ptr, stride = self._view(fields_view)
        self.lib.pfo_gravity_forward(
            *self._common(receivers, nodes, cell_nodes), _p(densities), ptr, stride,
            ctypes.c_int(KERNEL_ID[kname]), ctypes.c_double(factor), ctypes.c_int(self.parallel))

    def grav_gt_dot_v(self, receivers, nodes, cell_nodes, kname, factor, vector_view, result):
        ptr, stride = self._view(vector_view)
        self.lib.pfo_gravity_gt_dot_v(
            *self._common(receivers, nodes, cell_nodes), ptr, stride, _p(result),
            ctypes.c_int(KERNEL_ID[kname]), ctypes.c_double(factor), ctypes.c_int(self.parallel))

    def grav_diag(self, receivers, nodes, cell_nodes, kname, factor, weights, diagonal):
        self.lib.pfo_gravity_diag_gtg(
            *self._common(receivers, nodes, cell_nodes), _p(weights), _p(diagonal),
            ctypes.c_int(KERNEL_ID[kname]), ctypes.c_double(factor), ctypes.c_int(self.parallel))

    # magnetics ---------------------------------------------------------------
    @staticmethod
    def _kids(knames):
        return (ctypes.c_int * len(knames))(*[KERNEL_ID[k] for k in knames]), ctypes.c_int(len(knames))

    def mag_sensitivity(self, receivers, nodes, G_view, cell_nodes, b0, knames, factor, scalar):
        ptr, stride = self._view(G_view)
        kids, nk = self._kids(knames)
        self.lib.pfo_mag_sensitivity(
            *self._common(receivers, nodes, cell_nodes), ptr, stride,
            ctypes.c_int(int(G_view.dtype == np.float32)), _p(b0), kids, nk,
            ctypes.c_double(factor), ctypes.c_int(int(scalar)), ctypes.c_int(self.parallel))

    def mag_forward(self, receivers, nodes, model, fields_view, cell_nodes, b0, knames, factor, scalar):
        ptr, stride = self._view(fields_view)
        kids, nk = self._kids(knames)
        self.lib.pfo_mag_forward(
            *self._common(receivers, nodes, cell_nodes), _p(model), ptr, stride, _p(b0), kids, nk,
            ctypes.c_double(factor), ctypes.c_int(int(scalar)), ctypes.c_int(self.parallel))

    def mag_gt_dot_v(self, receivers, nodes, cell_nodes, b0, knames, factor, scalar, vector_view, result):
        ptr, stride = self._view(vector_view)
        kids, nk = self._kids(knames)
        self.lib.pfo_mag_gt_dot_v(
            *self._common(receivers, nodes, cell_nodes), ptr, stride, _p(result), _p(b0), kids, nk,
            ctypes.c_double(factor), ctypes.c_int(int(scalar)), ctypes.c_int(self.parallel))

    def mag_diag(self, receivers, nodes, cell_nodes, b0, knames, factor, scalar, weights, diagonal):
        kids, nk = self._kids(knames)
        self.lib.pfo_mag_diag_gtg(
            *self._common(receivers, nodes, cell_nodes), _p(weights), _p(diagonal), _p(b0), kids, nk,
            ctypes.c_double(factor), ctypes.c_int(int(scalar)), ctypes.c_int(self.parallel))


class _RefBackend:
    """Loop back-end on the reference's numba files (argument orders exactly as
    the reference call sites: gravity simulation.py:573-581, 620-627, 661-669,
    715-723; magnetics simulation.py:877-927, 977-1024, 1089-1139, 1181-1231)."""

    def __init__(self, parallel=True):
        from .ref_loader import load_reference, reference_kernel_tables

        _, self.g3, self.m3 = load_reference()
        self.gk, self.mk = reference_kernel_tables()
        self.parallel = bool(parallel)
        # kernel-name -> function (for the tuples of magnetic kernels)
        import choclo

        self._by_name = {k: getattr(choclo.prism, "kernel_" + k) for k in KERNEL_ID if k != "uv"}
        self._by_name["uv"] = self.gk["guv"]

    def grav_sensitivity(self, receivers, nodes, G_view, cell_nodes, kname, factor):
        self.g3.NUMBA_FUNCTIONS_3D["sensitivity"][self.parallel](
            receivers, nodes, G_view, cell_nodes, self._by_name[kname], factor)

    def grav_forward(self, receivers, nodes, densities, fields_view, cell_nodes, kname, factor):
        self.g3.NUMBA_FUNCTIONS_3D["forward"][self.parallel](
            receivers, nodes, densities, fields_view, cell_nodes, self._by_name[kname], factor)

    def grav_gt_dot_v(self, receivers, nodes, cell_nodes, kname, factor, vector_view, result):
        self.g3.NUMBA_FUNCTIONS_3D["gt_dot_v"][self.parallel](
            receivers, nodes, cell_nodes, self._by_name[kname], factor, vector_view, result)

    def grav_diag(self, receivers, nodes, cell_nodes, kname, factor, weights, diagonal):
        self.g3.NUMBA_FUNCTIONS_3D["diagonal_gtg"][self.parallel](
            receivers, nodes, cell_nodes, self._by_name[kname], factor, weights, diagonal)

    def _family(self, knames):
        if len(knames) == 3:
            return "magnetic_component"
        if tuple(knames) == MAGNETIC_KERNELS["tmi"]:
            return "tmi"
        return "tmi_derivative"

    def _kargs(self, knames):
        fam = self._family(knames)
        return fam, ([] if fam == "tmi" else [self._by_name[k] for k in knames])

    def mag_sensitivity(self, receivers, nodes, G_view, cell_nodes, b0, knames, factor, scalar):
        fam, kargs = self._kargs(knames)
        self.m3.NUMBA_FUNCTIONS_3D["sensitivity"][fam][self.parallel](
            receivers, nodes, G_view, cell_nodes, b0.copy(), *kargs, factor, scalar)

    def mag_forward(self, receivers, nodes, model, fields_view, cell_nodes, b0, knames, factor, scalar):
        fam, kargs = self._kargs(knames)
        self.m3.NUMBA_FUNCTIONS_3D["forward"][fam][self.parallel](
            receivers, nodes, model, fields_view, cell_nodes, b0.copy(), *kargs, factor, scalar)

    def mag_gt_dot_v(self, receivers, nodes, cell_nodes, b0, knames, factor, scalar, vector_view, result):
        fam, kargs = self._kargs(knames)
        self.m3.NUMBA_FUNCTIONS_3D["gt_dot_v"][fam][self.parallel](
            receivers, nodes, cell_nodes, b0.copy(), *kargs, factor, scalar, vector_view, result)

    def mag_diag(self, receivers, nodes, cell_nodes, b0, knames, factor, scalar, weights, diagonal):
        fam, kargs = self._kargs(knames)
        self.m3.NUMBA_FUNCTIONS_3D["diagonal_gtg"][fam][self.parallel](
            receivers, nodes, cell_nodes, b0.copy(), *kargs, factor, scalar, weights, diagonal)


class OracleSimulation:
    """
    Reference-order CPU evaluation of G, forward, G.T@v and diag(G.T W G).

    Parameters
    ----------
    physics : "gravity" | "magnetics"
    nodes : (nN, 3) float64, cell_nodes : (nC, 8) int64 (F-order vertices)
    receiver_list : list of (locations (R,3), components list[str])
    b0 : (3,) inducing field (magnetics), model_type : "scalar" | "vector"
    """

    def __init__(self, physics, nodes, cell_nodes, receiver_list, b0=None,
                 model_type="scalar", sensitivity_dtype=np.float64,
                 backend="c", parallel=True):
        self.physics = physics
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int64)
        self.receiver_list = [
            (np.ascontiguousarray(loc, dtype=np.float64),
             [comps] if isinstance(comps, str) else list(comps))
            for loc, comps in receiver_list
        ]
        self.b0 = None if b0 is None else np.ascontiguousarray(b0, dtype=np.float64)
        self.scalar = model_type == "scalar"
        self.dtype = sensitivity_dtype
        self.nC = self.cell_nodes.shape[0]
        self.nP = self.nC if (physics == "gravity" or self.scalar) else 3 * self.nC
        self.nD = sum(loc.shape[0] * len(c) for loc, c in self.receiver_list)
        self.backend = _CBackend(parallel) if backend == "c" else _RefBackend(parallel)

    def _iter(self):
        offset = 0
        for receivers, components in self.receiver_list:
            n_comp = len(components)
            n_rows = n_comp * receivers.shape[0]
            for i, component in enumerate(components):
                yield receivers, component, slice(offset + i, offset + n_rows, n_comp)
            offset += n_rows

    def sensitivity_matrix(self):
        G = np.empty((self.nD, self.nP), dtype=self.dtype)
        be = self.backend
        for receivers, component, sl in self._iter():
            if self.physics == "gravity":
                be.grav_sensitivity(receivers, self.nodes, G[sl, :], self.cell_nodes,
                                    GRAVITY_KERNELS[component],
                                    constants.G * gravity_conversion_factor(component))
            else:
                be.mag_sensitivity(receivers, self.nodes, G[sl, :], self.cell_nodes, self.b0,
                                   MAGNETIC_KERNELS[component], 1 / 4 / np.pi, self.scalar)
        return G

    def forward(self, model):
        model = np.ascontiguousarray(model, dtype=np.float64)
        fields = np.zeros(self.nD, dtype=np.float64)
        be = self.backend
        for receivers, component, sl in self._iter():
            if self.physics == "gravity":
                be.grav_forward(receivers, self.nodes, model, fields[sl], self.cell_nodes,
                                GRAVITY_KERNELS[component],
                                constants.G * gravity_conversion_factor(component))
            else:
                be.mag_forward(receivers, self.nodes, model, fields[sl], self.cell_nodes, self.b0,
                               MAGNETIC_KERNELS[component], 1 / 4 / np.pi, self.scalar)
        return fields

    def gt_dot_v(self, vector):
        vector = np.ascontiguousarray(vector, dtype=np.float64)
        result = np.zeros(self.nP, dtype=np.float64)
        be = self.backend
        for receivers, component, sl in self._iter():
            if self.physics == "gravity":
                be.grav_gt_dot_v(receivers, self.nodes, self.cell_nodes,
                                 GRAVITY_KERNELS[component],
                                 constants.G * gravity_conversion_factor(component),
                                 vector[sl], result)
            else:
                be.mag_gt_dot_v(receivers, self.nodes, self.cell_nodes, self.b0,
                                MAGNETIC_KERNELS[component], 1 / 4 / np.pi, self.scalar,
                                vector[sl], result)
        return result

    def diag_gtg(self, weights):
        """Mirrors the reference verbatim: the WHOLE weights array is handed to
        every per-component call, which indexes it with the receiver index."""
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        diagonal = np.zeros(self.nP, dtype=np.float64)
        be = self.backend
        for receivers, component, _ in self._iter():
            if self.physics == "gravity":
                be.grav_diag(receivers, self.nodes, self.cell_nodes,
                             GRAVITY_KERNELS[component],
                             constants.G * gravity_conversion_factor(component),
                             weights, diagonal)
            else:
                be.mag_diag(receivers, self.nodes, self.cell_nodes, self.b0,
                            MAGNETIC_KERNELS[component], 1 / 4 / np.pi, self.scalar,
                            weights, diagonal)
        return diagonal


# ---------------------------------------------------------------------------
# Equivalent-source layer (SimulationEquivalentSourceLayer)
# ---------------------------------------------------------------------------
class _CLayerBackend:
    """Layer loops of pf_oracle.c (pfo_layer_gravity / pfo_layer_mag)."""

    def __init__(self, parallel=True):
        self.lib = _library()
        self.parallel = int(bool(parallel))

    def gravity(self, mode, receivers, bounds, top, bottom, component, factor, G_view=None,
                vec=None, out=None):
        gp, gs, f32 = (None, _i64(0), 0) if G_view is None else \
            (*_CBackend._view(G_view), int(G_view.dtype == np.float32))
        vp, vs = (None, _i64(0)) if vec is None else _CBackend._view(vec)
        op, os_ = (None, _i64(0)) if out is None else _CBackend._view(out)
        self.lib.pfo_layer_gravity(
            ctypes.c_int(mode), _p(receivers), _i64(receivers.shape[0]), _p(bounds), _p(top),
            _p(bottom), _i64(bounds.shape[0]), gp, gs, ctypes.c_int(f32), vp, vs, op, os_,
            ctypes.c_int(KERNEL_ID[GRAVITY_KERNELS[component]]), ctypes.c_double(factor),
            ctypes.c_int(self.parallel))

    def mag(self, mode, receivers, bounds, top, bottom, component, b0, scalar, G_view=None,
            vec=None, out=None):
        gp, gs, f32 = (None, _i64(0), 0) if G_view is None else \
            (*_CBackend._view(G_view), int(G_view.dtype == np.float32))
        vp, vs = (None, _i64(0)) if vec is None else _CBackend._view(vec)
        op, os_ = (None, _i64(0)) if out is None else _CBackend._view(out)
        kids, nk = _CBackend._kids(MAGNETIC_KERNELS[component])
        # forward of "tmi" and of the b components goes through choclo's own forward functions
        choclo_forward = int(component == "tmi" or len(MAGNETIC_KERNELS[component]) == 3)
        self.lib.pfo_layer_mag(
            ctypes.c_int(mode), _p(receivers), _i64(receivers.shape[0]), _p(bounds), _p(top),
            _p(bottom), _i64(bounds.shape[0]), gp, gs, ctypes.c_int(f32), vp, vs, op, os_,
            _p(b0), kids, nk, ctypes.c_int(int(scalar)), ctypes.c_int(choclo_forward),
            ctypes.c_int(self.parallel))


class _RefLayerBackend:
    """The reference's own 2D-mesh numba loops; argument orders as the call sites
    gravity/simulation.py:801-810, 849-857, 888-897, 929-938 and
    magnetics/simulation.py:1320-1372, 1416-1466, 1505-1559, 1584-1651."""

    def __init__(self, parallel=True):
        from .ref_loader import load_reference_2d, reference_forward_tables, reference_kernel_tables

        self.g2, self.m2 = load_reference_2d()
        self.gf, self.mf = reference_forward_tables()
        _, self.mk = reference_kernel_tables()
        self.parallel = bool(parallel)

    def gravity(self, mode, receivers, bounds, top, bottom, component, factor, G_view=None,
                vec=None, out=None):
        F = self.g2.NUMBA_FUNCTIONS_2D
        fn = self.gf[component]
        if mode == 0:
            F["sensitivity"][self.parallel](receivers, bounds, top, bottom, G_view, fn, factor)
        elif mode == 1:
            F["forward"][self.parallel](receivers, bounds, top, bottom, vec, out, fn, factor)
        elif mode == 2:
            F["gt_dot_v"][self.parallel](receivers, bounds, top, bottom, fn, factor, vec, out)
        else:
            F["diagonal_gtg"][self.parallel](receivers, bounds, top, bottom, fn, factor, vec, out)

    def mag(self, mode, receivers, bounds, top, bottom, component, b0, scalar, G_view=None,
            vec=None, out=None):
        F = self.m2.NUMBA_FUNCTIONS_2D
        key = ["sensitivity", "forward", "gt_dot_v", "diagonal_gtg"][mode]
        if component == "tmi":
            fam, kargs = "tmi", []
        elif component in ("tmi_x", "tmi_y", "tmi_z"):
            fam, kargs = "tmi_derivative", list(self.mk[component])
        else:
            fam, kargs = "magnetic_component", list(self.mk[component])
        fn = F[key][fam][self.parallel]
        head = (receivers, bounds, top, bottom)
        cf = 1 / 4 / np.pi
        if mode == 0:
            fn(*head, G_view, b0.copy(), *kargs, scalar)
        elif mode == 1:
            if fam == "magnetic_component":
                kargs = [self.mf[component]]
            fn(*head, vec, out, b0.copy(), *kargs, scalar)
        elif mode == 2:
            fn(*head, b0.copy(), *kargs, scalar, vec, out)
        else:
            fn(*head, b0.copy(), *kargs, cf, scalar, vec, out)


class OracleLayerSimulation:
    """
    Reference-order CPU evaluation of the equivalent-source layer simulations
    (gravity/simulation.py:727-939, magnetics/simulation.py:1235-1653).

    Parameters
    ----------
    physics : "gravity" | "magnetics"
    cell_bounds : (nC, 4) x_min, x_max, y_min, y_max of the active cells of the 2D mesh
    top, bottom : (nC,) elevation of the top / bottom face of every cell
    receiver_list, b0, model_type : as OracleSimulation
    """

    def __init__(self, physics, cell_bounds, top, bottom, receiver_list, b0=None,
                 model_type="scalar", sensitivity_dtype=np.float64, backend="c", parallel=True):
        self.physics = physics
        self.bounds = np.ascontiguousarray(cell_bounds, dtype=np.float64).reshape(-1, 4)
        self.nC = self.bounds.shape[0]
        self.top = np.ascontiguousarray(np.broadcast_to(top, (self.nC,)), dtype=np.float64)
        self.bottom = np.ascontiguousarray(np.broadcast_to(bottom, (self.nC,)), dtype=np.float64)
        self.receiver_list = [
            (np.ascontiguousarray(loc, dtype=np.float64),
             [comps] if isinstance(comps, str) else list(comps))
            for loc, comps in receiver_list
        ]
        self.b0 = None if b0 is None else np.ascontiguousarray(b0, dtype=np.float64)
        self.scalar = model_type == "scalar"
        self.dtype = sensitivity_dtype
        self.nP = self.nC if (physics == "gravity" or self.scalar) else 3 * self.nC
        self.nD = sum(loc.shape[0] * len(c) for loc, c in self.receiver_list)
        self.backend = _CLayerBackend(parallel) if backend == "c" else _RefLayerBackend(parallel)

    _iter = OracleSimulation._iter

    def _call(self, mode, receivers, component, **kw):
        head = (mode, receivers, self.bounds, self.top, self.bottom, component)
        if self.physics == "gravity":
            self.backend.gravity(*head, gravity_conversion_factor(component), **kw)
        else:
            self.backend.mag(*head, self.b0, self.scalar, **kw)

    def sensitivity_matrix(self):
        G = np.empty((self.nD, self.nP), dtype=self.dtype)
        for receivers, component, sl in self._iter():
            self._call(0, receivers, component, G_view=G[sl, :])
        return G

    def forward(self, model):
        model = np.ascontiguousarray(model, dtype=np.float64)
        fields = np.zeros(self.nD, dtype=np.float64)
        for receivers, component, sl in self._iter():
            self._call(1, receivers, component, vec=model, out=fields[sl])
        return fields

    def gt_dot_v(self, vector):
        vector = np.ascontiguousarray(vector, dtype=np.float64)
        result = np.zeros(self.nP, dtype=np.float64)
        for receivers, component, sl in self._iter():
            self._call(2, receivers, component, vec=vector[sl], out=result)
        return result

    def diag_gtg(self, weights):
        """Whole weights array per component, indexed by receiver (reference quirk)."""
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        diagonal = np.zeros(self.nP, dtype=np.float64)
        for receivers, component, _ in self._iter():
            self._call(3, receivers, component, vec=weights, out=diagonal)
        return diagonal


# ---------------------------------------------------------------------------
# numpy tensor-mesh (nodes, cell_nodes) exactly as base.py:120-141 / discretize
# ---------------------------------------------------------------------------
def tensor_mesh_arrays(hx, hy, hz, origin):
    """nodes (nN,3) F-ordered and cell_nodes (nC,8) for a tensor mesh."""
    xs = origin[0] + np.r_[0.0, np.cumsum(hx)]
    ys = origin[1] + np.r_[0.0, np.cumsum(hy)]
    zs = origin[2] + np.r_[0.0, np.cumsum(hz)]
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    nodes = np.c_[X.reshape(-1, order="F"), Y.reshape(-1, order="F"), Z.reshape(-1, order="F")]
    shape_nodes = (len(xs), len(ys), len(zs))
    inds = np.arange(nodes.shape[0]).reshape(shape_nodes, order="F")
    cell_nodes = np.stack([
        inds[:-1, :-1, :-1].reshape(-1, order="F"),
        inds[1:, :-1, :-1].reshape(-1, order="F"),
        inds[:-1, 1:, :-1].reshape(-1, order="F"),
        inds[1:, 1:, :-1].reshape(-1, order="F"),
        inds[:-1, :-1, 1:].reshape(-1, order="F"),
        inds[1:, :-1, 1:].reshape(-1, order="F"),
        inds[:-1, 1:, 1:].reshape(-1, order="F"),
        inds[1:, 1:, 1:].reshape(-1, order="F"),
    ], axis=-1)
    return np.ascontiguousarray(nodes), np.ascontiguousarray(cell_nodes.astype(np.int64))
