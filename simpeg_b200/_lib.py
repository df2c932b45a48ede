"""
ctypes binding of ``libpf_cuda.so`` (C ABI in ``include/pf_cuda.h``).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, a ``RuntimeError`` is raised (``engine="cuda"`` has no CPU path).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import os

# PF_CUDA_LIB: development override (A/B builds of the same ABI)
LIB_PATH = Path(os.environ.get("PF_CUDA_LIB", Path(__file__).resolve().parent / "libpf_cuda.so"))

PF_MAX_COMPONENTS = 16
PF_ABI_VERSION = 6
PF_F64, PF_F32 = 0, 1
PF_GRAVITY, PF_MAGNETICS = 0, 1

# pf_component (include/pf_cuda.h); names are the reference's component strings
# (gravity/receivers.py:56-71, magnetics/receivers.py:46-62).
GRAVITY_COMPONENTS = {
    "gx": 0, "gy": 1, "gz": 2,
    "gxx": 3, "gyy": 4, "gzz": 5, "gxy": 6, "gxz": 7, "gyz": 8,
    "guv": 9,
}
MAGNETIC_COMPONENTS = {
    "tmi": 16, "bx": 17, "by": 18, "bz": 19,
    "bxx": 20, "byy": 21, "bzz": 22, "bxy": 23, "bxz": 24, "byz": 25,
    "tmi_x": 26, "tmi_y": 27, "tmi_z": 28,
}

# every symbol declared in include/pf_cuda.h
EXPORTED_SYMBOLS = (
    "pf_abi_version", "pf_last_error", "pf_launch_count", "pf_job_uses_marching", "pf_job_columns",
    "pf_describe",
    "pf_build_G",
    "pf_forward_workspace_bytes", "pf_forward",
    "pf_gt_dot_v_workspace_bytes", "pf_gt_dot_v",
    "pf_gtg_diag_workspace_bytes", "pf_gtg_diag",
    "pf_gemv_workspace_bytes", "pf_gemv",
    "pf_gemv_t_workspace_bytes", "pf_gemv_t",
    "pf_colsumsq", "pf_colsumsq_acc", "pf_colsumsq_amp",
    "pf_amp_normalize", "pf_amp_combine", "pf_amp_spread",
    "pf_plan_tensor_count", "pf_plan_tensor_fill",
    "pf_probe_fp64_peak",
    "pf_cg_workspace_bytes", "pf_cg_init", "pf_cg_hvp_combine", "pf_cg_update", "pf_cg_direction",
    "pf_vec_mul", "pf_csr_spmv", "pf_host_copy", "pf_math_probe",
    "pf_plan_generic_assign", "pf_plan_generic_fill",
)


class PfPlan(ctypes.Structure):
    _fields_ = [
        ("n_tiles", ctypes.c_int64),
        ("n_cells", ctypes.c_int64),
        ("tile_cells", ctypes.c_int32),
        ("tile_nodes", ctypes.c_int32),
        ("tile_xyz", ctypes.c_void_p),
        ("tile_cell_nodes", ctypes.c_void_p),
        ("tile_cell_col", ctypes.c_void_p),
        ("tile_n_nodes", ctypes.c_void_p),
        ("grid_nx", ctypes.c_int32), ("grid_ny", ctypes.c_int32), ("grid_nz", ctypes.c_int32),
        ("n_footprints4", ctypes.c_int32), ("n_footprints2", ctypes.c_int32),
        ("n_footprints4h", ctypes.c_int32),
        ("grid_xs", ctypes.c_void_p), ("grid_ys", ctypes.c_void_p), ("grid_zs", ctypes.c_void_p),
        ("grid_colmap", ctypes.c_void_p),
        ("footprints4", ctypes.c_void_p), ("footprints2", ctypes.c_void_p),
        ("n_nodes", ctypes.c_int64), ("node_xyz", ctypes.c_void_p), ("cell_nodes32", ctypes.c_void_p),
        ("footprints4h", ctypes.c_void_p),
    ]


class PfJob(ctypes.Structure):
    _fields_ = [
        ("physics", ctypes.c_int32),
        ("n_components", ctypes.c_int32),
        ("components", ctypes.c_int32 * PF_MAX_COMPONENTS),
        ("scalar_model", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("b0", ctypes.c_double * 3),
        ("receivers", ctypes.c_void_p),
        ("n_receivers", ctypes.c_int64),
        ("row_offset", ctypes.c_int64),
    ]


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libpf_cuda.so; raise RuntimeError (never fall back) when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: engine='cuda' needs the compiled sm_100a library "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C simpeg_b200/csrc`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    PP, JP = ctypes.POINTER(PfPlan), ctypes.POINTER(PfJob)
    sigs = {
        "pf_abi_version": (i32, []),
        "pf_last_error": (ctypes.c_char_p, []),
        "pf_launch_count": (i64, []),
        "pf_job_uses_marching": (i32, [PP, JP, i32]),
        "pf_job_columns": (i64, [PP, JP]),
        "pf_describe": (i32, [PP, JP, i32, i32, ctypes.c_char_p, i64]),
        "pf_build_G": (i32, [PP, JP, vp, i64, i32, vp]),
        "pf_forward_workspace_bytes": (i64, [PP, JP]),
        "pf_forward": (i32, [PP, JP, vp, vp, vp, i64, vp]),
        "pf_gt_dot_v_workspace_bytes": (i64, [PP, JP]),
        "pf_gt_dot_v": (i32, [PP, JP, vp, vp, vp, i64, vp]),
        "pf_gtg_diag_workspace_bytes": (i64, [PP, JP]),
        "pf_gtg_diag": (i32, [PP, JP, vp, vp, vp, i64, vp]),
        "pf_gemv_workspace_bytes": (i64, [i64, i64]),
        "pf_gemv": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, i64, vp]),
        "pf_gemv_t_workspace_bytes": (i64, [i64, i64]),
        "pf_gemv_t": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, i64, vp]),
        "pf_colsumsq": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, i64, vp]),
        "pf_colsumsq_acc": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, i64, vp]),
        "pf_colsumsq_amp": (i32, [vp, i64, i64, i64, i32, vp, vp, vp, vp, i64, vp]),
        "pf_amp_normalize": (i32, [vp, i64, vp, vp, vp]),
        "pf_amp_combine": (i32, [vp, vp, i64, vp, vp]),
        "pf_amp_spread": (i32, [vp, vp, i64, vp, vp]),
        "pf_plan_tensor_count": (i64, [i64] * 6 + [vp, i64, vp]),
        "pf_plan_tensor_fill": (i32, [i64] * 3 + [vp] * 3 + [i64] * 3 + [vp, i64, vp, i64] + [vp] * 4),
        "pf_probe_fp64_peak": (i32, [ctypes.POINTER(dbl), vp]),
        "pf_cg_workspace_bytes": (i64, [i64]),
        "pf_cg_init": (i32, [vp, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp]),
        "pf_cg_hvp_combine": (i32, [vp] * 7 + [dbl, dbl, vp, i64, vp, vp, vp, i64, vp]),
        "pf_cg_update": (i32, [vp, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp]),
        "pf_cg_direction": (i32, [vp, i64, vp, vp, vp]),
        "pf_vec_mul": (i32, [vp, vp, i64, vp, vp]),
        "pf_csr_spmv": (i32, [vp, vp, vp, i64, vp, vp, vp]),
        "pf_host_copy": (i32, [vp, vp, i64]),
        "pf_math_probe": (i32, [i32, vp, vp, vp, vp, i64, vp]),
        "pf_plan_generic_assign": (i64, [vp, i64, vp, i64, vp]),
        "pf_plan_generic_fill": (i32, [vp, i64, vp, i64, vp, i64, vp, vp, vp, vp]),
    }
    for name, (restype, argtypes) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pf_abi_version() != PF_ABI_VERSION:
        raise RuntimeError(
            f"libpf_cuda.so ABI version {lib.pf_abi_version()} != expected {PF_ABI_VERSION}"
        )
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    """Turn a negative pf_status into an exception carrying pf_last_error()."""
    if status >= 0:
        return
    msg = load_library().pf_last_error().decode(errors="replace")
    if status == -2:
        raise NotImplementedError(f"{what}: {msg}")
    if status == -1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed (status {status}): {msg}")
