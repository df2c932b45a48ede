"""
Device-side engine: owns the torch buffers (tile plan, receivers, G, scratch)
and issues the C-ABI calls of ``libpf_cuda.so`` on the current CUDA stream.

PyTorch is used for device memory and streams only; all arithmetic happens in
the hand-written sm_100a kernels behind ``include/pf_cuda.h``.

The engine works on the LOCAL receiver shard of this process (see
``simpeg_b200/sharding.py``); rows are local rows.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .tiling import TILE_CELLS, TILE_NODES, TilePlan, build_tile_plan

_G_ALIGN = 16  # leading dimension of G in elements (128 B rows for fp64)
_PINNED = {"turn": 0, "bufs": [None, None], "events": [None, None]}


def _pieces(n, min_piece=1 << 19, max_pieces=4):
    """Split [0, n) into at most ``max_pieces`` ranges of at least ``min_piece`` doubles (4 MB)."""
    k = max(1, min(max_pieces, n // min_piece))
    bounds = [n * j // k for j in range(k + 1)]
    return [(bounds[j], bounds[j + 1]) for j in range(k) if bounds[j + 1] > bounds[j]] or [(0, 0)]


def _require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "engine='cuda' requires a CUDA device (no CPU fallback): "
            "torch.cuda.is_available() is False"
        )
    return torch


class CudaEngine:
    """
    Parameters
    ----------
    physics : "gravity" | "magnetics"
    nodes, cell_nodes : active nodes (nN, 3) and cell -> node table (nC, 8)
    receiver_list : list of (locations (R, 3), components list[str]) -- LOCAL shard
    b0 : (3,) inducing field for magnetics
    model_type : "scalar" | "vector"
    sensitivity_dtype : np.float64 | np.float32 (storage of G)
    device : torch device / index (default: current device)
    """

    def __init__(self, physics, nodes, cell_nodes, receiver_list, b0=None,
                 model_type="scalar", sensitivity_dtype=np.float64, device=None,
                 plan: TilePlan | None = None, use_marching=True, use_node_space=True):
        torch = _require_cuda()
        self.torch = torch
        self.lib = _lib.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        if physics not in ("gravity", "magnetics"):
            raise ValueError(f"unknown physics {physics!r}")
        self.physics = physics
        self.scalar = (model_type == "scalar")
        if sensitivity_dtype not in (np.float64, np.float32):
            raise TypeError("sensitivity_dtype must be np.float32 or np.float64")
        self.g_dtype = sensitivity_dtype
        self._g_code = _lib.PF_F64 if sensitivity_dtype is np.float64 else _lib.PF_F32
        self._g_torch = torch.float64 if sensitivity_dtype is np.float64 else torch.float32

        self.host_plan = plan if plan is not None else build_tile_plan(nodes, cell_nodes)
        hp = self.host_plan
        self.nC = hp.n_cells
        self.nblk = 1 if (physics == "gravity" or self.scalar) else 3
        self.nP = self.nC * self.nblk
        self.plan = _lib.PfPlan(n_tiles=0, n_cells=hp.n_cells, tile_cells=TILE_CELLS, tile_nodes=TILE_NODES)
        if hp.has_tiles:
            self._upload_tiles()
        self.use_node_space = bool(use_node_space)
        self.use_marching = bool(use_marching) and hp.grid_shape is not None
        if self.use_marching:
            # tensor mesh: enable the z-marching kernel for G build / forward
            with torch.cuda.device(self.device):
                self._t_axes = [torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
                                for a in hp.grid_axes]
                self._t_colmap = None if hp.grid_colmap is None else \
                    torch.from_numpy(hp.grid_colmap).to(self.device)
                self._t_fp4 = torch.from_numpy(hp.footprints4).to(self.device)
                self._t_fp2 = torch.from_numpy(hp.footprints2).to(self.device)
                fp4h = getattr(hp, "footprints4h", None)
                self._t_fp4h = None if fp4h is None or fp4h.shape[0] == 0 else torch.from_numpy(fp4h).to(self.device)
            self.plan.grid_nx, self.plan.grid_ny, self.plan.grid_nz = hp.grid_shape
            self.plan.grid_xs, self.plan.grid_ys, self.plan.grid_zs = (t.data_ptr() for t in self._t_axes)
            self.plan.grid_colmap = None if self._t_colmap is None else self._t_colmap.data_ptr()
            self.plan.n_footprints4 = hp.footprints4.shape[0]
            self.plan.n_footprints2 = hp.footprints2.shape[0]
            self.plan.footprints4 = self._t_fp4.data_ptr() if hp.footprints4.shape[0] else None
            self.plan.footprints2 = self._t_fp2.data_ptr() if hp.footprints2.shape[0] else None
            self.plan.n_footprints4h = 0 if self._t_fp4h is None else self._t_fp4h.shape[0]
            self.plan.footprints4h = None if self._t_fp4h is None else self._t_fp4h.data_ptr()

        table = _lib.GRAVITY_COMPONENTS if physics == "gravity" else _lib.MAGNETIC_COMPONENTS
        code = _lib.PF_GRAVITY if physics == "gravity" else _lib.PF_MAGNETICS
        b0 = np.zeros(3) if b0 is None else np.asarray(b0, dtype=np.float64).reshape(3)
        self.jobs = []
        self._rx = []
        offset = 0
        for locations, components in receiver_list:
            components = [components] if isinstance(components, str) else list(components)
            unknown = [c for c in components if c not in table]
            if unknown:
                raise NotImplementedError(
                    f"Other components besides {set(table)} aren't implemented yet "
                    f"(got {unknown}).")
            if len(components) > _lib.PF_MAX_COMPONENTS:
                raise ValueError("too many components in one receiver object")
            loc = np.ascontiguousarray(locations, dtype=np.float64).reshape(-1, 3)
            rx = torch.from_numpy(loc).to(self.device)
            self._rx.append(rx)
            job = _lib.PfJob()
            job.physics = code
            job.n_components = len(components)
            for i, c in enumerate(components):
                job.components[i] = table[c]
            job.scalar_model = int(self.scalar)
            job.b0[:] = list(b0)
            job.receivers = rx.data_ptr() if loc.shape[0] else None
            job.n_receivers = loc.shape[0]
            job.row_offset = offset
            self.jobs.append(job)
            offset += loc.shape[0] * len(components)
        self.n_rows = offset
        self.G = None
        self.ld = 0
        self._ws = None

    def _upload_tiles(self):
        """Brick / RCB tile arrays -> device (built lazily for tensor meshes: the z-marching
        kernel does not need them)."""
        if self.plan.n_tiles:
            return
        torch = self.torch
        hp = self.host_plan.ensure_tiles()
        if hp.n_tiles == 0:
            return
        with torch.cuda.device(self.device):
            self._t_xyz = torch.from_numpy(hp.tile_xyz).to(self.device)
            # torch has no uint16 arithmetic; the bytes are only ever read by the kernel
            self._t_cn = torch.from_numpy(hp.tile_cell_nodes.view(np.int16)).to(self.device)
            self._t_col = torch.from_numpy(hp.tile_cell_col).to(self.device)
            self._t_nn = torch.from_numpy(hp.tile_n_nodes).to(self.device)
        self.plan.n_tiles = hp.n_tiles
        self.plan.tile_xyz = self._t_xyz.data_ptr()
        self.plan.tile_cell_nodes = self._t_cn.data_ptr()
        self.plan.tile_cell_col = self._t_col.data_ptr()
        self.plan.tile_n_nodes = self._t_nn.data_ptr()

    def _need_tiles_unless_marching(self, forward):
        if self.plan.n_tiles:
            return
        for job in self.jobs:
            if not self.lib.pf_job_uses_marching(*self._jp(job), int(forward)):
                self._upload_tiles()
                return

    # ------------------------------------------------------------------ utils
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def copy_stream(self):
        """Side stream for D2H transfers that overlap with kernels of the compute stream."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = self.torch.cuda.Stream(device=self.device)
        return self._copy_stream

    def _workspace(self, nbytes):
        nbytes = max(int(nbytes), 8)
        if self._ws is None or self._ws.numel() * 8 < nbytes:
            self._ws = None
            self._ws = self.torch.empty((nbytes + 7) // 8, dtype=self.torch.float64, device=self.device)
        return self._ws

    def _pinned(self, n):
        """Rotating pair of pinned host staging buffers, shared by all engines of the
        process (page-locking memory is slow, so they are grown, never freed)."""
        torch = self.torch
        st = _PINNED
        st["turn"] = 1 - st["turn"]
        i = st["turn"]
        if st["events"][i] is not None:
            st["events"][i].synchronize()   # the previous async copy using it is done
            st["events"][i] = None
        if st["bufs"][i] is None or st["bufs"][i].numel() < n:
            st["bufs"][i] = torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True)
        return i, st["bufs"][i][:n]

    def to_device(self, array, n=None):
        """host vector -> device, through pinned memory (async on the current stream)."""
        a = np.ascontiguousarray(array, dtype=np.float64).reshape(-1)
        if n is not None and a.shape[0] != n:
            raise ValueError(f"expected a vector of length {n}, got {a.shape[0]}")
        torch = self.torch
        n = a.shape[0]
        src = torch.from_numpy(a) if a.flags.writeable else None
        if src is not None and n and src.is_pinned():
            # already page-locked (e.g. a vector this engine returned): no staging copy; the
            # source is kept alive until the next transfer has been ordered after this one
            with torch.cuda.device(self.device):
                dev = torch.empty(n, dtype=torch.float64, device=self.device)
                dev.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            prev = _PINNED.get("direct")
            if prev is not None:
                prev[0].synchronize()
            _PINNED["direct"] = (ev, src)
            return dev
        i, pin = self._pinned(n)
        with torch.cuda.device(self.device):
            dev = torch.empty(n, dtype=torch.float64, device=self.device)
            # long vectors in a few pieces: the async H2D of one piece overlaps the staging copy
            # of the next
            for lo, hi in _pieces(n):
                self.lib.pf_host_copy(pin.data_ptr() + 8 * lo, a.ctypes.data + 8 * lo, 8 * (hi - lo))
                dev[lo:hi].copy_(pin[lo:hi], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        _PINNED["events"][i] = ev
        return dev

    def to_host(self, tensor):
        """device vector -> new numpy array.  The array is backed by page-locked memory from
        torch's caching host allocator (it goes back to the cache when the array is garbage
        collected), so the D2H transfer lands directly in what the caller receives -- no staging
        copy from a pinned buffer into pageable memory (a 24 MB memcpy per Jtvec at BASELINE
        config 4, on every rank at once)."""
        torch = self.torch
        t = tensor.detach().reshape(-1)
        if t.dtype != torch.float64 or t.device.type != "cuda":
            return t.cpu().numpy()
        with torch.cuda.device(self.device):
            out = torch.empty(t.numel(), dtype=torch.float64, pin_memory=True)
            out.copy_(t, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return out.numpy()

    def _jp(self, job):
        return ctypes.byref(self.plan), ctypes.byref(job)

    def describe(self, sink="store"):
        """Kernel instantiations the jobs of this engine launch for a sink ("store", "forward",
        "gtv", "diag") -- asked from the library's own dispatch (``pf_describe``)."""
        code = {"store": 0, "forward": 1, "gtv": 2, "diag": 3}[sink]
        names = []
        for job in self.jobs:
            buf = ctypes.create_string_buffer(1024)
            _lib.check(self.lib.pf_describe(*self._jp(job), code, self._g_code, buf, 1024), "pf_describe")
            name = buf.value.decode()
            if name and name not in names:
                names.append(name)
        return " ; ".join(names)

    # ----------------------------------------------------------- build / sinks
    def build_G(self):
        """Fill the local block of G on the device (rows_local x nP, row-major)."""
        torch = self.torch
        self._need_tiles_unless_marching(False)
        self.ld = -(-max(self.nP, 1) // _G_ALIGN) * _G_ALIGN
        with torch.cuda.device(self.device):
            if self.G is None:
                need = self.n_rows * self.ld * (8 if self.g_dtype is np.float64 else 4)
                free, total = torch.cuda.mem_get_info(self.device)
                reusable = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
                if need > free + reusable:
                    raise RuntimeError(
                        f"the local block of G needs {need / 1e9:.1f} GB but only "
                        f"{(free + reusable) / 1e9:.1f} GB of the {total / 1e9:.0f} GB of this GPU are free: "
                        "shard the receivers over more GPUs (torchrun), use sensitivity_dtype=np.float32 "
                        "or store_sensitivities='forward_only'")
                self.G = torch.empty((self.n_rows, self.ld), dtype=self._g_torch, device=self.device)
            for job in self.jobs:
                _lib.check(self.lib.pf_build_G(*self._jp(job), self.G.data_ptr(), self.ld,
                                               self._g_code, self._stream()), "pf_build_G")
        return self.G

    def forward_dev(self, model_t, out=None):
        """fields = G_local @ model without storing G (device tensors)."""
        torch = self.torch
        self._need_tiles_unless_marching(True)
        with torch.cuda.device(self.device):
            fields = torch.zeros(self.n_rows, dtype=torch.float64, device=self.device) if out is None \
                else out.zero_()
            for job in self.jobs:
                nbytes = self.lib.pf_forward_workspace_bytes(*self._jp(job))
                ws = self._workspace(nbytes)
                _lib.check(self.lib.pf_forward(*self._jp(job), model_t.data_ptr(), fields.data_ptr(),
                                               ws.data_ptr(), ws.numel() * 8, self._stream()),
                           "pf_forward")
        return fields

    def _column_sink(self, fn, wsfn, vec_t, name, out=None):
        torch = self.torch
        with torch.cuda.device(self.device):
            result = torch.zeros(self.nP, dtype=torch.float64, device=self.device) if out is None \
                else out.zero_()
            for job in self.jobs:
                nbytes = wsfn(*self._jp(job))
                ws = self._workspace(nbytes)
                _lib.check(fn(*self._jp(job), vec_t.data_ptr(), result.data_ptr(), ws.data_ptr(),
                              ws.numel() * 8, self._stream()), name)
        return result

    def _enable_node_space(self):
        """Upload the global node table (once) so that pf_gt_dot_v works in node space."""
        if self.plan.n_nodes or not self.use_node_space or self.nC == 0:
            return
        torch = self.torch
        node_xyz, cn32 = self.host_plan.node_table()
        with torch.cuda.device(self.device):
            self._t_node_xyz = torch.from_numpy(node_xyz).to(self.device)
            self._t_cn32 = torch.from_numpy(cn32).to(self.device)
        self.plan.n_nodes = node_xyz.shape[1]
        self.plan.node_xyz = self._t_node_xyz.data_ptr()
        self.plan.cell_nodes32 = self._t_cn32.data_ptr()

    def gt_dot_v_dev(self, v_t, out=None):
        """G_local.T @ v_local without storing G."""
        self._enable_node_space()
        if not self.plan.n_nodes:
            self._upload_tiles()
        return self._column_sink(self.lib.pf_gt_dot_v, self.lib.pf_gt_dot_v_workspace_bytes, v_t,
                                 "pf_gt_dot_v", out)

    def gtg_diag_dev(self, w_t, out=None):
        """sum_rows w[row] G_local[row, :]**2 without storing G."""
        self._need_tiles_unless_marching(False)     # tensor meshes: blocked marching build + column norms
        return self._column_sink(self.lib.pf_gtg_diag, self.lib.pf_gtg_diag_workspace_bytes, w_t,
                                 "pf_gtg_diag", out)

    # ------------------------------------------------------- stored-G products
    def _need_G(self):
        if self.G is None:
            self.build_G()
        return self.G

    def gemv_dev(self, x_t, out=None):
        torch = self.torch
        G = self._need_G()
        with torch.cuda.device(self.device):
            y = torch.empty(self.n_rows, dtype=torch.float64, device=self.device) if out is None else out
            ws = self._workspace(self.lib.pf_gemv_workspace_bytes(self.n_rows, self.nP))
            _lib.check(self.lib.pf_gemv(G.data_ptr(), self.n_rows, self.nP, self.ld, self._g_code,
                                        x_t.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                        self._stream()), "pf_gemv")
        return y

    def gemv_t_dev(self, y_t, out=None, cols=None):
        """z = G_local.T @ y on the device.  ``cols=(c0, c1)``: only that column slab (written to
        ``out[c0:c1]``; c0 a multiple of 16 keeps the 128-bit loads aligned) -- the unit the sharded
        Jtvec overlaps with the all-reduce of the previous slab."""
        torch = self.torch
        G = self._need_G()
        c0, c1 = (0, self.nP) if cols is None else cols
        item = 8 if self.g_dtype is np.float64 else 4
        with torch.cuda.device(self.device):
            z = torch.empty(self.nP, dtype=torch.float64, device=self.device) if out is None else out
            ws = self._workspace(self.lib.pf_gemv_t_workspace_bytes(self.n_rows, c1 - c0))
            _lib.check(self.lib.pf_gemv_t(G.data_ptr() + c0 * item, self.n_rows, c1 - c0, self.ld, self._g_code,
                                          y_t.data_ptr(), z.data_ptr() + 8 * c0, ws.data_ptr(), ws.numel() * 8,
                                          self._stream()), "pf_gemv_t")
        return z

    def colsumsq_dev(self, w_t, out=None):
        torch = self.torch
        G = self._need_G()
        with torch.cuda.device(self.device):
            d = torch.empty(self.nP, dtype=torch.float64, device=self.device) if out is None else out
            ws = self._workspace(self.lib.pf_gemv_t_workspace_bytes(self.n_rows, self.nP))
            _lib.check(self.lib.pf_colsumsq(G.data_ptr(), self.n_rows, self.nP, self.ld, self._g_code,
                                            w_t.data_ptr(), d.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                            self._stream()), "pf_colsumsq")
        return d

    # --------------------------------------------------------------- amplitude
    def amp_normalize_dev(self, b_t):
        """(amplitude (n_loc,), ampderiv (3, n_loc)) of interleaved (bx,by,bz) rows."""
        torch = self.torch
        n_loc = b_t.numel() // 3
        with torch.cuda.device(self.device):
            amp = torch.empty(n_loc, dtype=torch.float64, device=self.device)
            ad = torch.empty((3, n_loc), dtype=torch.float64, device=self.device)
            _lib.check(self.lib.pf_amp_normalize(b_t.data_ptr(), n_loc, amp.data_ptr(), ad.data_ptr(),
                                                 self._stream()), "pf_amp_normalize")
        return amp, ad

    def amp_combine_dev(self, ad_t, jv_t):
        torch = self.torch
        n_loc = ad_t.shape[1]
        with torch.cuda.device(self.device):
            out = torch.empty(n_loc, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.pf_amp_combine(ad_t.data_ptr(), jv_t.data_ptr(), n_loc, out.data_ptr(),
                                               self._stream()), "pf_amp_combine")
        return out

    def amp_spread_dev(self, ad_t, v_t):
        torch = self.torch
        n_loc = ad_t.shape[1]
        with torch.cuda.device(self.device):
            out = torch.empty(3 * n_loc, dtype=torch.float64, device=self.device)
            _lib.check(self.lib.pf_amp_spread(ad_t.data_ptr(), v_t.data_ptr(), n_loc, out.data_ptr(),
                                              self._stream()), "pf_amp_spread")
        return out

    def colsumsq_amp_dev(self, w_t, ad_t):
        torch = self.torch
        G = self._need_G()
        n_loc = self.n_rows // 3
        with torch.cuda.device(self.device):
            d = torch.empty(self.nP, dtype=torch.float64, device=self.device)
            ws = self._workspace(self.lib.pf_gemv_t_workspace_bytes(n_loc, self.nP))
            _lib.check(self.lib.pf_colsumsq_amp(G.data_ptr(), n_loc, self.nP, self.ld, self._g_code,
                                                w_t.data_ptr(), ad_t.data_ptr(), d.data_ptr(),
                                                ws.data_ptr(), ws.numel() * 8, self._stream()),
                       "pf_colsumsq_amp")
        return d

    def _alloc_G(self):
        torch = self.torch
        self.ld = -(-max(self.nP, 1) // _G_ALIGN) * _G_ALIGN
        if self.G is None:
            with torch.cuda.device(self.device):
                self.G = torch.empty((self.n_rows, self.ld), dtype=self._g_torch, device=self.device)
        return self.G

    def G_row_chunks(self, max_bytes=256 << 20):
        """Yield ``(start, stop, array)`` host copies of consecutive local rows of G."""
        G = self._need_G()
        item = 8 if self.g_dtype is np.float64 else 4
        step = max(1, int(max_bytes // max(self.nP * item, 1)))
        for a in range(0, self.n_rows, step):
            b = min(a + step, self.n_rows)
            yield a, b, G[a:b, : self.nP].cpu().numpy()

    def load_G_rows(self, source, rows, max_bytes=256 << 20):
        """Fill the local block of G from ``source[rows]`` (a host array / memmap of the FULL
        matrix, ``rows`` = global index of every local row) instead of evaluating it."""
        torch = self.torch
        G = self._alloc_G()
        item = 8 if self.g_dtype is np.float64 else 4
        step = max(1, int(max_bytes // max(self.nP * item, 1)))
        for a in range(0, self.n_rows, step):
            b = min(a + step, self.n_rows)
            block = np.ascontiguousarray(source[rows[a:b]], dtype=self.g_dtype)
            G[a:b, : self.nP].copy_(torch.from_numpy(block))
        return G

    def G_to_numpy(self):
        """Copy the local block of G to the host (small problems / tests only)."""
        G = self._need_G()
        return G[:, : self.nP].cpu().numpy()


def launch_count() -> int:
    """CUDA kernels launched by libpf_cuda.so in this process so far."""
    return int(_lib.load_library().pf_launch_count())


def probe_fp64_peak(device=None):
    """Measured FP64 FMA-pipe peak in FLOP/s (roofline denominator of the build)."""
    torch = _require_cuda()
    lib = _lib.load_library()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = ctypes.c_double(0.0)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.pf_probe_fp64_peak(ctypes.byref(out), stream), "pf_probe_fp64_peak")
    return float(out.value)
