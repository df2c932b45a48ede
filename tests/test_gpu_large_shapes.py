"""GPU: the two size limits the round-1 advisor found -- pf_gemv beyond 65535 row blocks per launch
and pf_forward's per-tile partial sums beyond the workspace cap (receiver chunks) -- plus the
accumulating column norms and the CSR product of the ABI, called the way a binding would."""
import ctypes

import numpy as np
import pytest

from simpeg_b200 import _lib

pytestmark = pytest.mark.gpu


def _stream(torch):
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_gemv_with_more_than_65535_row_blocks():
    import torch

    lib = _lib.load_library()
    rows, cols, ld = 65535 * 32 + 1000, 24, 32          # one launch covers at most 65535 x 32 rows
    g = torch.Generator(device="cuda").manual_seed(0)
    G = torch.randn((rows, ld), dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(cols, dtype=torch.float64, device="cuda", generator=g)
    y = torch.empty(rows, dtype=torch.float64, device="cuda")
    nbytes = lib.pf_gemv_workspace_bytes(rows, cols)
    ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device="cuda")
    _lib.check(lib.pf_gemv(G.data_ptr(), rows, cols, ld, _lib.PF_F64, x.data_ptr(), y.data_ptr(), ws.data_ptr(),
                           ws.numel() * 8, _stream(torch)), "pf_gemv")
    ref = G[:, :cols] @ x
    assert float((y - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # wide enough for several column chunks as well (partials indexed with the full row count)
    rows2, cols2 = 65535 * 32 + 64, 5000
    ld2 = 5008
    G2 = torch.randn((rows2, ld2), dtype=torch.float32, device="cuda", generator=g)
    x2 = torch.randn(cols2, dtype=torch.float64, device="cuda", generator=g)
    y2 = torch.empty(rows2, dtype=torch.float64, device="cuda")
    ws2 = torch.empty(lib.pf_gemv_workspace_bytes(rows2, cols2) // 8, dtype=torch.float64, device="cuda")
    _lib.check(lib.pf_gemv(G2.data_ptr(), rows2, cols2, ld2, _lib.PF_F32, x2.data_ptr(), y2.data_ptr(),
                           ws2.data_ptr(), ws2.numel() * 8, _stream(torch)), "pf_gemv")
    ref2 = G2[:, :cols2].double() @ x2
    assert float((y2 - ref2).abs().max()) <= 1e-5 * float(ref2.abs().max())


def test_forward_in_receiver_chunks():
    """Tree meshes keep one partial sum per tile and receiver; with enough receivers the partials
    exceed the 256 MB cap and pf_forward walks the receivers in chunks."""
    import torch

    from simpeg_b200.engine import CudaEngine
    from simpeg_b200.meshes import synthetic_octree

    mesh = synthetic_octree((24, 24, 6), 10.0, levels=3)
    rng = np.random.default_rng(1)
    n_rx = 150_000
    loc = np.c_[rng.uniform(-400, 400, (n_rx, 2)), rng.uniform(5.0, 30.0, n_rx)]
    eng = CudaEngine("gravity", mesh.total_nodes, mesh.cell_nodes, [(loc, ["gz"])])
    eng._upload_tiles()
    lib = eng.lib
    full = eng.plan.n_tiles * n_rx * 8
    asked = lib.pf_forward_workspace_bytes(*eng._jp(eng.jobs[0]))
    assert asked < full and asked <= (256 << 20) + eng.plan.n_tiles * 64 * 8      # chunked
    m = rng.uniform(-0.5, 0.5, eng.nP)
    d = eng.to_host(eng.forward_dev(eng.to_device(m)))
    for lo, hi in ((0, 300), (n_rx // 2, n_rx // 2 + 300), (n_rx - 300, n_rx)):
        part = CudaEngine("gravity", mesh.total_nodes, mesh.cell_nodes, [(loc[lo:hi], ["gz"])], plan=eng.host_plan)
        ref = part.to_host(part.gemv_dev(part.to_device(m)))
        assert np.max(np.abs(d[lo:hi] - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_colsumsq_acc_and_csr_spmv():
    import scipy.sparse as sp
    import torch

    lib = _lib.load_library()
    rng = np.random.default_rng(2)
    G = rng.normal(size=(700, 3000))
    w = rng.uniform(0.5, 1.5, 700)
    d0 = rng.normal(size=3000)
    Gt = torch.from_numpy(G).cuda()
    wt, dt = torch.from_numpy(w).cuda(), torch.from_numpy(d0.copy()).cuda()
    ws = torch.empty(lib.pf_gemv_t_workspace_bytes(700, 3000) // 8 + 1, dtype=torch.float64, device="cuda")
    _lib.check(lib.pf_colsumsq_acc(Gt.data_ptr(), 700, 3000, 3000, _lib.PF_F64, wt.data_ptr(), dt.data_ptr(),
                                   ws.data_ptr(), ws.numel() * 8, _stream(torch)), "pf_colsumsq_acc")
    ref = d0 + np.einsum("i,ij,ij->j", w, G, G)
    assert np.max(np.abs(dt.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    A = sp.random(900, 400, density=0.02, random_state=3, format="csr")
    x = rng.normal(size=400)
    y = torch.empty(900, dtype=torch.float64, device="cuda")
    ip, ix = torch.from_numpy(A.indptr.astype(np.int32)).cuda(), torch.from_numpy(A.indices.astype(np.int32)).cuda()
    va, xt = torch.from_numpy(A.data).cuda(), torch.from_numpy(x).cuda()
    _lib.check(lib.pf_csr_spmv(ip.data_ptr(), ix.data_ptr(), va.data_ptr(), 900, xt.data_ptr(), y.data_ptr(),
                               _stream(torch)), "pf_csr_spmv")
    assert np.max(np.abs(y.cpu().numpy() - A @ x)) <= 1e-13 * max(np.max(np.abs(A @ x)), 1.0)
