"""CPU: libpf_cuda.so loads without a GPU and exports exactly what
include/pf_cuda.h declares; argument errors come back as status codes."""
import ctypes
import re
from pathlib import Path

import pytest

from simpeg_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pf_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(pf_[A-Za-z0-9_]+)\s*\(", text))


def test_header_and_loader_agree():
    assert declared_symbols() == set(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_struct_layout():
    lib = _lib.load_library()
    assert lib.pf_abi_version() == _lib.PF_ABI_VERSION
    assert ctypes.sizeof(_lib.PfPlan) == 160
    assert ctypes.sizeof(_lib.PfJob) == 128


def test_argument_errors_are_status_codes():
    lib = _lib.load_library()
    plan = _lib.PfPlan(n_tiles=0, n_cells=0, tile_cells=512, tile_nodes=768)
    job = _lib.PfJob()
    job.physics = 7
    job.n_components = 1
    st = lib.pf_build_G(ctypes.byref(plan), ctypes.byref(job), None, 0, 0, None)
    assert st == -1 and b"physics" in lib.pf_last_error()
    job.physics = _lib.PF_GRAVITY
    job.components[0] = _lib.MAGNETIC_COMPONENTS["tmi"]
    st = lib.pf_build_G(ctypes.byref(plan), ctypes.byref(job), None, 0, 0, None)
    assert st == -2
    with pytest.raises(NotImplementedError):
        _lib.check(st, "pf_build_G")
    bad = _lib.PfPlan(n_tiles=1, n_cells=1, tile_cells=256, tile_nodes=768)
    job.components[0] = _lib.GRAVITY_COMPONENTS["gz"]
    assert lib.pf_build_G(ctypes.byref(bad), ctypes.byref(job), None, 0, 0, None) == -1
    # empty problems are fine without touching the device
    assert lib.pf_build_G(ctypes.byref(plan), ctypes.byref(job), None, 0, 0, None) == 0
    assert lib.pf_gemv(None, 0, 0, 0, 0, None, None, None, 0, None) == 0
    assert lib.pf_gemv(None, 4, 4, 2, 0, None, None, None, 0, None) == -1


def test_host_staging_copy_and_cg_argument_checks():
    import numpy as np

    lib = _lib.load_library()
    # pf_host_copy is a host-only entry point: usable without a GPU
    for n in (0, 7, 131072, 3_000_001):
        src = np.random.default_rng(n).uniform(size=n)
        dst = np.full(n, -1.0)
        assert lib.pf_host_copy(dst.ctypes.data, src.ctypes.data, src.nbytes) == 0
        assert np.array_equal(src, dst)
    assert lib.pf_host_copy(None, None, 16) == -1
    # the CG entry points validate before touching the device
    assert lib.pf_cg_workspace_bytes(10) > 0
    assert lib.pf_cg_init(None, None, None, 5, None, None, None, None, None, 0, None) == -1
    assert lib.pf_vec_mul(None, None, 0, None, None) == 0
    assert lib.pf_vec_mul(None, None, 3, None, None) == -1
