"""
Receiver-row sharding of G across processes (one process per GPU).

Every row of G depends on one receiver only -- the reference parallelises the
same axis with ``prange`` over receivers (gravity/_numba/_3d_mesh.py:145) -- so
each rank owns a contiguous block of the locations of every receiver object
and therefore the rows

    row = object_offset + i_location * n_components + i_component

of those locations.  The mesh and the model vectors are replicated.  Only two
exchanges exist on the path:

* ``dpred`` / ``Jvec`` / ``G @ m``: every rank holds the entries of its rows ->
  all-gather into the full (nD,) vector;
* ``Jtvec`` / ``G.T @ v`` / ``getJtJdiag``: every rank holds a partial (nP,)
  vector -> all-reduce (sum).

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.
Building G needs no communication at all.
"""
from __future__ import annotations

import numpy as np


def block_bounds(n: int, world_size: int, rank: int):
    """Contiguous block [lo, hi) of ``n`` items for ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class ReceiverSharding:
    """
    Parameters
    ----------
    receiver_list : list of (locations (R, 3), components list[str]) -- the whole survey
    rank, world_size : position of this process
    """

    def __init__(self, receiver_list, rank=0, world_size=1):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of size {world_size}")
        self.rank = int(rank)
        self.world_size = int(world_size)
        self._objects = []
        offset = 0
        for locations, components in receiver_list:
            components = [components] if isinstance(components, str) else list(components)
            locations = np.asarray(locations, dtype=np.float64).reshape(-1, 3)
            self._objects.append((locations, components, offset))
            offset += locations.shape[0] * len(components)
        self.nD = offset
        self.local_receiver_list = []
        for locations, components, _ in self._objects:
            lo, hi = block_bounds(locations.shape[0], world_size, rank)
            self.local_receiver_list.append((np.ascontiguousarray(locations[lo:hi]), components))
        self.local_rows = self.rows_of_rank(self.rank)

    def rows_of_rank(self, rank):
        """Global row index of every local row of ``rank``, in local row order."""
        rows = []
        for locations, components, offset in self._objects:
            lo, hi = block_bounds(locations.shape[0], self.world_size, rank)
            c = len(components)
            rows.append(offset + np.arange(lo * c, hi * c, dtype=np.int64))
        return np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)

    @property
    def n_local_rows(self):
        return int(self.local_rows.size)

    def max_local_rows(self):
        return max(self.rows_of_rank(r).size for r in range(self.world_size))

    def gather_source_index(self):
        """
        For the padded all-gather buffer (world_size blocks of ``max_local_rows()`` entries):
        position of every global row, so that ``full = gathered[index]`` (cached).
        """
        if getattr(self, "_gather_index", None) is None:
            m = self.max_local_rows()
            index = np.empty(self.nD, dtype=np.int64)
            for r in range(self.world_size):
                rows = self.rows_of_rank(r)
                index[rows] = r * m + np.arange(rows.size)
            if np.array_equal(index, np.arange(self.nD)):
                index = None       # equal contiguous blocks: the gathered buffer IS the full vector
            self._gather_index = (m, index)
        return self._gather_index


    def gather_source_index_on(self, device):
        """The gather index as a tensor on ``device`` -- cached on THIS object (one entry per
        device), so that it lives and dies with the row layout it describes."""
        import torch

        cache = self.__dict__.setdefault("_gather_index_t", {})
        key = str(device)
        if key not in cache:
            _, index = self.gather_source_index()
            cache[key] = None if index is None else torch.from_numpy(index).to(device)
        return cache[key]


class _Done:
    """Completed-work handle of a single-process "collective"."""

    def wait(self):
        return True


class Collectives:
    """Thin wrapper over ``torch.distributed`` (a no-op for a single process)."""

    def __init__(self, group=None, world_size=None, rank=None):
        self.group = group
        self._dist = None
        if world_size is None:
            try:
                import torch.distributed as dist

                if dist.is_available() and dist.is_initialized():
                    self._dist = dist
                    world_size = dist.get_world_size(group)
                    rank = dist.get_rank(group)
            except ImportError:  # pragma: no cover
                pass
        elif world_size > 1:
            import torch.distributed as dist

            self._dist = dist
        self.world_size = 1 if world_size is None else int(world_size)
        self.rank = 0 if rank is None else int(rank)
        if self.world_size == 1:
            self._dist = None

    def barrier(self):
        if self._dist is not None:
            self._dist.barrier(group=self.group)

    def all_reduce_sum(self, tensor):
        """In-place sum over ranks of a partial (nP,) vector (Jtvec, JtJ diagonal)."""
        if self._dist is not None:
            self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group)
        return tensor

    def all_reduce_sum_async(self, tensor):
        """Start the in-place sum over ranks and return a handle whose ``wait()`` orders the
        current stream after it (NCCL runs the reduction on its own stream, so kernels launched
        in between overlap with it)."""
        if self._dist is None:
            return _Done()
        return self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_gather_slices(self, make_slice, n):
        """Full (n,) device vector from equal slices: rank r provides ``make_slice(lo, hi)`` for its
        block of ``ceil(n / world)`` entries (the last block may be short and is padded)."""
        import torch

        m = -(-int(n) // self.world_size)
        lo = min(self.rank * m, n)
        hi = min(lo + m, n)
        piece = make_slice(lo, hi)
        if hi - lo < m:
            padded = torch.zeros(m, dtype=piece.dtype, device=piece.device)
            padded[: hi - lo] = piece
            piece = padded
        full = torch.empty(m * self.world_size, dtype=piece.dtype, device=piece.device)
        self._dist.all_gather_into_tensor(full, piece.contiguous(), group=self.group)
        return full[:n]

    def all_gather_rows(self, local, sharding: ReceiverSharding):
        """Assemble the full (nD,) vector from every rank's local rows (dpred, Jvec)."""
        import torch

        if self._dist is None:
            return local            # one process owns every row, already in global order
        m, index = sharding.gather_source_index()
        index_t = sharding.gather_source_index_on(local.device) if index is not None else None
        if local.numel() == m:
            padded = local.contiguous()
        else:
            padded = torch.zeros(m, dtype=local.dtype, device=local.device)
            padded[: local.numel()] = local
        gathered = torch.empty(m * self.world_size, dtype=local.dtype, device=local.device)
        self._dist.all_gather_into_tensor(gathered, padded, group=self.group)
        return gathered[: sharding.nD] if index is None else gathered[index_t]
