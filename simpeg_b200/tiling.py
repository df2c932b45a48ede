"""
Tile planner: host-side preparation of the mesh for the CUDA kernels.

The reference evaluates the prism kernels once per *unique active node* and
gathers 8 values per cell through ``cell_nodes``
(``BasePFSimulation._get_active_nodes``, base.py:371-399, consumed by
gravity/_numba/_3d_mesh.py:147-163).  On the GPU the same node sharing is done
at tile granularity: active cells are partitioned into compact tiles of at
most ``TILE_CELLS`` cells whose vertices are at most ``TILE_NODES`` distinct
nodes, so that one CTA can keep the per-receiver node kernel values of a tile
in shared memory.

Two partitioners:

* tensor-like meshes (every active cell has a distinct integer (ix, iy, iz)
  position on the grid of unique cell-centre coordinates; inactive cells
  allowed): axis-aligned bricks whose dimensions are chosen to minimise the
  number of tiles;
* anything else (TreeMesh with hanging nodes, arbitrary cell_nodes tables):
  recursive coordinate bisection with cuts placed between distinct centre
  coordinates, until both limits hold.

The cell order inside G is untouched: every tile slot records the original
(active) cell index as its G column.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

TILE_CELLS = 512
TILE_NODES = 768
_ROUND = 256  # threads per CTA: cost is paid per round of 256 nodes / cells
_MARCH_HALF = 128  # threads per CTA of the half-size variant of the z-marching kernel (footprints4h)


@dataclass
class TilePlan:
    """Host arrays of the tile plan (see include/pf_cuda.h ``pf_plan``)."""

    n_tiles: int
    n_cells: int
    tile_xyz: np.ndarray         # (n_tiles, 3, TILE_NODES) float64
    tile_cell_nodes: np.ndarray  # (n_tiles, TILE_CELLS, 8) uint16
    tile_cell_col: np.ndarray    # (n_tiles, TILE_CELLS) int32, -1 = padding
    tile_n_nodes: np.ndarray     # (n_tiles,) int32
    method: str = ""
    # tensor meshes only (z-marching kernel, see include/pf_cuda.h)
    grid_shape: tuple | None = None
    grid_axes: tuple | None = None        # (xs, ys, zs) float64
    grid_colmap: np.ndarray | None = None  # (nx*ny*nz,) int32, None = all active
    footprints4: np.ndarray | None = None  # (n, 6) int32
    footprints2: np.ndarray | None = None
    footprints4h: np.ndarray | None = None  # <= 512 nodes, for 128-thread CTAs (three-plane programs)
    active_index: np.ndarray | None = None  # tensor meshes: mesh indices of the active cells
    # generic plans keep the (nodes, cell_nodes) they were built from
    nodes: np.ndarray | None = None
    cell_nodes: np.ndarray | None = None

    def ensure_tiles(self):
        """Build the brick tile arrays of a tensor plan created with ``with_tiles=False``."""
        if self.tile_xyz is None:
            xs, ys, zs = self.grid_axes
            (self.n_tiles, self.tile_xyz, self.tile_cell_nodes, self.tile_cell_col,
             self.tile_n_nodes) = _tensor_tiles(xs, ys, zs, self.active_index)
        return self

    @property
    def has_tiles(self):
        return self.tile_xyz is not None

    def node_table(self):
        """
        (node_xyz (3, nN) float64, cell_nodes (nC, 8) int32): the unique active nodes and the
        cell -> node table (what ``_get_active_nodes`` returns, base.py:371-399), for the
        node-space G^T v kernel.
        """
        if self.nodes is not None:
            nodes, cn = self.nodes, self.cell_nodes
        else:
            xs, ys, zs = self.grid_axes
            nx, ny, nz = self.grid_shape
            cells = np.arange(nx * ny * nz, dtype=np.int64) if self.active_index is None \
                else np.asarray(self.active_index, dtype=np.int64)
            ix, iy, iz = cells % nx, (cells // nx) % ny, cells // (nx * ny)
            n0 = ix + (nx + 1) * (iy + (ny + 1) * iz)
            sx, sy = 1, nx + 1
            sz = (nx + 1) * (ny + 1)
            cn = np.stack([n0, n0 + sx, n0 + sy, n0 + sx + sy,
                           n0 + sz, n0 + sx + sz, n0 + sy + sz, n0 + sx + sy + sz], axis=1)
            X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
            nodes = np.stack([X.reshape(-1, order="F"), Y.reshape(-1, order="F"), Z.reshape(-1, order="F")], axis=1)
        # keep only the nodes that are used (inactive cells)
        used = np.zeros(nodes.shape[0], dtype=bool)
        used[np.asarray(cn).ravel()] = True
        if not used.all():
            remap = np.cumsum(used) - 1
            nodes = nodes[used]
            cn = remap[cn]
        return (np.ascontiguousarray(np.asarray(nodes, dtype=np.float64).T),
                np.ascontiguousarray(np.asarray(cn, dtype=np.int32)))

    @property
    def node_evaluations(self) -> int:
        """(receiver-independent) number of node kernel evaluations per receiver"""
        self.ensure_tiles()
        return int(self.tile_n_nodes.sum())

    @property
    def nodes_per_cell(self) -> float:
        return self.node_evaluations / max(self.n_cells, 1)


def _cell_centers(nodes, cell_nodes):
    centers = np.empty((cell_nodes.shape[0], 3))
    for ax in range(3):
        centers[:, ax] = nodes[:, ax][cell_nodes].mean(axis=1)
    return centers


def _tile_cost(n_cells, n_nodes):
    # rounds of 256 threads; a node round costs ~4x a cell round (atan/log vs adds)
    return 4 * -(-n_nodes // _ROUND) + -(-n_cells // _ROUND)


def _brick_tile_ids(centers):
    """Brick decomposition for tensor-like meshes, or None."""
    n = centers.shape[0]
    idx = []
    dims = []
    for ax in range(3):
        u, inv = np.unique(centers[:, ax], return_inverse=True)
        idx.append(inv.astype(np.int64))
        dims.append(len(u))
    if dims[0] * dims[1] * dims[2] < n or dims[0] * dims[1] * dims[2] > 8 * n + 4096:
        return None
    lin = idx[0] + dims[0] * (idx[1] + dims[1] * idx[2])
    if len(np.unique(lin)) != n:
        return None
    bx, by, bz = _best_brick(dims)
    nbx, nby = -(-dims[0] // bx), -(-dims[1] // by)
    tile = (idx[0] // bx) + nbx * ((idx[1] // by) + nby * (idx[2] // bz))
    return tile


def _rcb_tile_ids(centers, cell_nodes):
    """Recursive coordinate bisection with cuts between distinct coordinates."""
    n = centers.shape[0]
    tile = np.zeros(n, dtype=np.int64)
    next_id = 1
    stack = [np.arange(n)]
    done = []
    while stack:
        ids = stack.pop()
        if len(ids) <= TILE_CELLS and len(np.unique(cell_nodes[ids])) <= TILE_NODES:
            done.append(ids)
            continue
        c = centers[ids]
        ext = c.max(axis=0) - c.min(axis=0)
        split = None
        for ax in np.argsort(-ext):
            vals = c[:, ax]
            u = np.unique(vals)
            if len(u) < 2:
                continue
            # cut between distinct coordinates, as close to the median as possible
            med = np.median(vals)
            cuts = 0.5 * (u[:-1] + u[1:])
            cut = cuts[np.argmin(np.abs(cuts - med))]
            left = vals < cut
            split = (ids[left], ids[~left])
            break
        if split is None:
            # all centres coincide (degenerate); split by count
            h = len(ids) // 2
            split = (ids[:h], ids[h:])
        stack.extend(split)
    for k, ids in enumerate(done):
        tile[ids] = k
    del next_id
    return tile


def _enforce_limits(tile, centers, cell_nodes):
    """Split any tile that exceeds the cell or node limits (RCB on that tile)."""
    counts = np.bincount(tile)
    order = np.argsort(tile, kind="stable")
    starts = np.r_[0, np.cumsum(counts)]
    next_id = len(counts)
    for t in range(len(counts)):
        ids = order[starts[t]:starts[t + 1]]
        if len(ids) == 0:
            continue
        if len(ids) <= TILE_CELLS and len(np.unique(cell_nodes[ids])) <= TILE_NODES:
            continue
        sub = _rcb_tile_ids(centers[ids], cell_nodes[ids])
        sub_new = np.where(sub == 0, t, next_id + sub - 1)
        tile[ids] = sub_new
        next_id += int(sub.max())
    # compact ids
    _, tile = np.unique(tile, return_inverse=True)
    return tile.astype(np.int64)


_BRICK_CACHE = {}


def _best_brick(dims):
    """Brick dimensions minimising the total tile cost for a (nx, ny, nz) grid."""
    dims = tuple(int(d) for d in dims)
    if dims in _BRICK_CACHE:
        return _BRICK_CACHE[dims]
    best = None
    for bx in range(1, 17):
        for by in range(1, 17):
            for bz in range(1, 17):
                if bx * by * bz > TILE_CELLS or (bx + 1) * (by + 1) * (bz + 1) > TILE_NODES:
                    continue
                parts = []
                for d, b in zip(dims, (bx, by, bz)):
                    full, rem = divmod(d, b)
                    parts.append([(b, full)] + ([(rem, 1)] if rem else []))
                cost = 0
                for sx, cx in parts[0]:
                    for sy, cy in parts[1]:
                        for sz, cz in parts[2]:
                            cost += cx * cy * cz * _tile_cost(sx * sy * sz, (sx + 1) * (sy + 1) * (sz + 1))
                key = (cost, -bx)  # prefer long x runs (contiguous G columns)
                if best is None or key < best[0]:
                    best = (key, (bx, by, bz))
    _BRICK_CACHE[dims] = best[1]
    return best[1]


def _tensor_tiles(xs, ys, zs, act):
    """Brick tile arrays of a tensor mesh (native planner, csrc/pf_plan.cu)."""
    import ctypes

    from . import _lib

    lib = _lib.load_library()
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    bx, by, bz = _best_brick((nx, ny, nz))
    n_bricks = (-(-nx // bx)) * (-(-ny // by)) * (-(-nz // bz))
    used = np.zeros(n_bricks, dtype=np.uint8)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    n_act = 0 if act is None else act.size
    n_tiles = lib.pf_plan_tensor_count(nx, ny, nz, bx, by, bz, ptr(act), n_act, ptr(used))
    _lib.check(int(min(n_tiles, 0)), "pf_plan_tensor_count")
    tile_xyz = np.empty((n_tiles, 3, TILE_NODES), dtype=np.float64)
    tile_cell_nodes = np.zeros((n_tiles, TILE_CELLS, 8), dtype=np.uint16)
    tile_cell_col = np.empty((n_tiles, TILE_CELLS), dtype=np.int32)
    tile_n_nodes = np.empty(n_tiles, dtype=np.int32)
    _lib.check(lib.pf_plan_tensor_fill(nx, ny, nz, ptr(xs), ptr(ys), ptr(zs), bx, by, bz, ptr(act),
                                       n_act, ptr(used), n_tiles, ptr(tile_xyz), ptr(tile_cell_nodes),
                                       ptr(tile_cell_col), ptr(tile_n_nodes)), "pf_plan_tensor_fill")
    return int(n_tiles), tile_xyz, tile_cell_nodes, tile_cell_col, tile_n_nodes


def build_tile_plan_tensor(axis_nodes, active_index=None, with_tiles=True) -> TilePlan:
    """
    O(n) planner for tensor meshes: pure index arithmetic, done natively by
    ``pf_plan_tensor_count`` / ``pf_plan_tensor_fill`` (csrc/pf_plan.cu).

    Parameters
    ----------
    axis_nodes : (nodes_x, nodes_y, nodes_z) 1-D node coordinates of the mesh
    active_index : int array of the ACTIVE cells' mesh indices (F-order), ascending,
        or None when every cell is active.  Column k of G is active cell k.
    with_tiles : build the brick tile arrays now (True) or only when a kernel that needs
        them is used (``TilePlan.ensure_tiles``); the z-marching kernel only needs the
        footprints.
    """
    xs, ys, zs = (np.ascontiguousarray(a, dtype=np.float64) for a in axis_nodes)
    nx, ny, nz = len(xs) - 1, len(ys) - 1, len(zs) - 1
    act = None if active_index is None else np.ascontiguousarray(active_index, dtype=np.int64)
    n_cells = nx * ny * nz if act is None else act.size
    if n_cells == 0:
        return TilePlan(0, 0, np.zeros((0, 3, TILE_NODES)), np.zeros((0, TILE_CELLS, 8), np.uint16),
                        np.zeros((0, TILE_CELLS), np.int32), np.zeros(0, np.int32), "empty")
    if n_cells >= 2**31:
        raise ValueError("more than 2^31-1 active cells are not supported")
    colmap = None
    if act is not None:
        colmap = np.full(nx * ny * nz, -1, dtype=np.int32)
        colmap[act] = np.arange(act.size, dtype=np.int32)
    plan = TilePlan(0, int(n_cells), None, None, None, None, "tensor", grid_shape=(nx, ny, nz),
                    grid_axes=(xs, ys, zs), grid_colmap=colmap, active_index=act)
    # node slots per receiver of the brick decomposition (without building it)
    bx, by, bz = _best_brick((nx, ny, nz))
    brick_slots = 0.0
    for sx, cx in _axis_parts(nx, bx):
        for sy, cy in _axis_parts(ny, by):
            for sz, cz in _axis_parts(nz, bz):
                brick_slots += cx * cy * cz * (-(-((sx + 1) * (sy + 1) * (sz + 1)) // _ROUND) * _ROUND)
    # keep a footprint set only if it evaluates fewer node slots per cell than the bricks
    for name, ncap, threads in (("footprints4", 1024, _ROUND), ("footprints2", 512, _ROUND),
                                ("footprints4h", 4 * _MARCH_HALF, _MARCH_HALF)):
        fps = march_footprints(nx, ny, nz, ncap, act, threads)
        rounds = march_node_rounds(fps[:, 2].astype(np.int64), fps[:, 3].astype(np.int64), threads)
        slots = float(((rounds * threads) * (fps[:, 5] - fps[:, 4] + 1)).sum())
        setattr(plan, name, fps if slots < 0.95 * brick_slots else np.zeros((0, 6), np.int32))
    # the half-size set only pays when it evaluates no more node slots than the full-size one
    if plan.footprints4.shape[0] and plan.footprints4h.shape[0]:
        def total(fps, threads):
            r = march_node_rounds(fps[:, 2].astype(np.int64), fps[:, 3].astype(np.int64), threads)
            return float(((r * threads) * (fps[:, 5] - fps[:, 4] + 1)).sum())
        if total(plan.footprints4h, _MARCH_HALF) > 1.02 * total(plan.footprints4, _ROUND):
            plan.footprints4h = np.zeros((0, 6), np.int32)
    if with_tiles:
        plan.ensure_tiles()
    return plan


def _axis_parts(d, b):
    full, rem = divmod(d, b)
    return [(b, full)] + ([(rem, 1)] if rem else [])


def march_node_rounds(w, h, threads=_ROUND):
    """Node rounds per layer of a w x h-cell footprint (whole node rows per round of `threads`)."""
    rows = np.maximum(threads // (np.asarray(w) + 1), 1)
    return -(-(np.asarray(h) + 1) // rows)


def march_footprints(nx, ny, nz, ncap, active_index=None, threads=_ROUND):
    """
    (x, y) footprints for the z-marching kernel: an a x b grid of rectangles,
    each with (w + 1)(h + 1) <= ncap nodes per layer and w + 1 <= threads, chosen to
    minimise the number of node rounds per layer (threads = 256, or 128 for the
    half-size CTAs of the three-plane programs); rows are
    (ix0, iy0, w, h, iz_begin, iz_end).

    The kernel takes a layer's nodes floor(threads / (w + 1)) whole rows per round
    (a thread keeps one node column, csrc/pf_build.cu), so a footprint costs
    ceil((h + 1) / floor(threads / (w + 1))) node rounds.
    """
    # every (a, b) split (this runs inside every simulation's first G access): one numpy pass over
    # the whole (a, b) table for ordinary meshes, row blocks of it for huge ones (bounded memory)
    b_all = np.arange(1, ny + 1, dtype=np.int64)[None, :]
    h_all = -(-ny // b_all)
    rows_per_block = max(1, (1 << 20) // ny)
    best = None
    for a0 in range(1, nx + 1, rows_per_block):
        a_blk = np.arange(a0, min(a0 + rows_per_block, nx + 1), dtype=np.int64)[:, None]
        w_blk = -(-nx // a_blk)
        nodes_all = (w_blk + 1) * (h_all + 1)
        rounds_all = march_node_rounds(w_blk, h_all, threads)
        cost = (a_blk * b_all * (4 * rounds_all + -(-(w_blk * h_all) // threads))).astype(np.float64)
        cost[(nodes_all > ncap) | ((w_blk + 1) * 2 > ncap) | (w_blk + 1 > threads)] = np.inf
        flat = int(np.argmin(cost))            # row-major: smallest a, then smallest b, among equals
        c = float(cost.flat[flat])
        if np.isfinite(c) and (best is None or c < best[0]):
            best = (c, a0 + flat // ny, flat % ny + 1)   # strict <: ties keep the smaller a (longer G runs)
    _, a, b = best
    xb = [nx * i // a for i in range(a + 1)]
    yb = [ny * j // b for j in range(b + 1)]
    fps = []
    if active_index is not None:
        act = np.asarray(active_index, dtype=np.int64)
        aix, aiy, aiz = act % nx, (act // nx) % ny, act // (nx * ny)
    for j in range(b):
        for i in range(a):
            x0, x1, y0, y1 = xb[i], xb[i + 1], yb[j], yb[j + 1]
            z0, z1 = 0, nz
            if active_index is not None:
                inside = (aix >= x0) & (aix < x1) & (aiy >= y0) & (aiy < y1)
                if not inside.any():
                    continue
                z0, z1 = int(aiz[inside].min()), int(aiz[inside].max()) + 1
            fps.append((x0, y0, x1 - x0, y1 - y0, z0, z1))
    return np.asarray(fps, dtype=np.int32).reshape(-1, 6)


def tensor_axes_of(mesh):
    """(nodes_x, nodes_y, nodes_z) if ``mesh`` is a tensor mesh whose ``nodes`` /
    ``cell_nodes`` follow the F-order convention of base.py:120-141, else None."""
    if hasattr(mesh, "total_nodes") or not all(
            hasattr(mesh, a) for a in ("nodes_x", "nodes_y", "nodes_z", "shape_cells")):
        return None
    axes = tuple(np.asarray(getattr(mesh, a), dtype=np.float64) for a in ("nodes_x", "nodes_y", "nodes_z"))
    if tuple(len(a) - 1 for a in axes) != tuple(mesh.shape_cells):
        return None
    return axes


def build_tile_plan(nodes, cell_nodes, method="auto") -> TilePlan:
    """
    Parameters
    ----------
    nodes : (nN, 3) float array -- active node locations
    cell_nodes : (nC, 8) int array -- node indices of every active cell, F-order
    method : "auto" | "brick" | "rcb"
    """
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int64)
    if nodes.ndim != 2 or nodes.shape[1] != 3:
        raise ValueError(f"nodes must be (nN, 3), got {nodes.shape}")
    if cell_nodes.ndim != 2 or cell_nodes.shape[1] != 8:
        raise ValueError(f"cell_nodes must be (nC, 8), got {cell_nodes.shape}")
    n_cells = cell_nodes.shape[0]
    n_nodes = nodes.shape[0]
    if n_cells and (cell_nodes.min() < 0 or cell_nodes.max() >= n_nodes):
        raise ValueError("cell_nodes holds indices outside the nodes array")
    if n_cells >= 2**31:
        raise ValueError("more than 2^31-1 active cells are not supported")
    if n_cells == 0:
        return TilePlan(0, 0, np.zeros((0, 3, TILE_NODES)), np.zeros((0, TILE_CELLS, 8), np.uint16),
                        np.zeros((0, TILE_CELLS), np.int32), np.zeros(0, np.int32), "empty")

    from . import _lib

    lib = _lib.load_library()
    ptr = lambda a: a.ctypes.data  # noqa: E731
    tile = None
    used = method
    if method in ("auto", "brick"):
        centers = _cell_centers(nodes, cell_nodes)
        tile = _brick_tile_ids(centers)
        used = "brick"
        if tile is None and method == "brick":
            raise ValueError("mesh is not tensor-like; use method='rcb'")
        if tile is not None:
            _, tile = np.unique(tile, return_inverse=True)  # compact (bricks may be empty)
            tile = _enforce_limits(tile.astype(np.int64), centers, cell_nodes)
    if tile is None and method == "rcb_numpy":
        # the numpy version of the bisection (kept as the readable statement of the rule and as a
        # cross-check of the native planner in the tests)
        centers = _cell_centers(nodes, cell_nodes)
        tile = _rcb_tile_ids(centers, cell_nodes)
        _, tile = np.unique(tile, return_inverse=True)
        tile = _enforce_limits(tile.astype(np.int64), centers, cell_nodes)
        used = "rcb"
    if tile is None:
        # native recursive coordinate bisection (csrc/pf_plan.cu), O(n) per level
        tile32 = np.empty(n_cells, dtype=np.int32)
        n_tiles = int(lib.pf_plan_generic_assign(ptr(nodes), n_nodes, ptr(cell_nodes), n_cells, ptr(tile32)))
        _lib.check(min(n_tiles, 0), "pf_plan_generic_assign")
        used = "rcb"
    else:
        tile32 = np.ascontiguousarray(tile, dtype=np.int32)
        n_tiles = int(tile32.max()) + 1

    tile_xyz = np.empty((n_tiles, 3, TILE_NODES), dtype=np.float64)
    tile_cell_nodes = np.empty((n_tiles * TILE_CELLS, 8), dtype=np.uint16)
    tile_cell_col = np.empty(n_tiles * TILE_CELLS, dtype=np.int32)
    ncounts = np.empty(n_tiles, dtype=np.int32)
    _lib.check(lib.pf_plan_generic_fill(ptr(nodes), n_nodes, ptr(cell_nodes), n_cells, ptr(tile32), n_tiles,
                                        ptr(tile_xyz), ptr(tile_cell_nodes), ptr(tile_cell_col),
                                        ptr(ncounts)), "pf_plan_generic_fill")

    return TilePlan(n_tiles, n_cells, tile_xyz,
                    tile_cell_nodes.reshape(n_tiles, TILE_CELLS, 8),
                    tile_cell_col.reshape(n_tiles, TILE_CELLS),
                    ncounts.astype(np.int32), used, nodes=nodes, cell_nodes=cell_nodes)
