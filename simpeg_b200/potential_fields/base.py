"""
``BasePFSimulation`` with ``engine="cuda"`` -- the host-side mirror of
simpeg/potential_fields/base.py:24-408 for the integral-equation hot path.

Kept from the reference: constructor keywords, ``store_sensitivities``
('ram' / 'forward_only' / 'disk': built on the GPU, kept in HBM and also written to
``sensitivity_path`` in the reference's memmap layout),
``sensitivity_dtype`` (forced to float64 when ``forward_only``,
base.py:196-215), ``engine`` validation with the reference's error text
(base.py:241-256), ``_get_active_nodes`` (base.py:371-399, cached here) and
``_get_components_and_receivers`` (base.py:401-408).

New: everything below the four private methods runs on the GPU through
``simpeg_b200.engine.CudaEngine`` (C ABI ``include/pf_cuda.h``).  Receivers are
sharded over the ranks of ``torch.distributed`` when it is initialised
(``simpeg_b200/sharding.py``).  There is no CPU fallback.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator, aslinearoperator

from .. import zarr_store
from ..utils import validate_active_indices, validate_string

ENGINES = ("geoana", "choclo", "cuda")
SLICED_UPLOAD_MIN = 1 << 18   # replicated host vectors at least this long are uploaded in rank slices


JTVEC_SLABS = int(os.environ.get("PF_JTVEC_SLABS", "2"))               # column slabs of a sharded Jtvec (kernel of slab k+1 overlaps all-reduce of k)
JTVEC_SLAB_ALIGN = 4096       # slab starts are multiples of this many columns (aligned 128-bit loads)
JTVEC_SLAB_MIN = 1 << 18      # narrower matrices are not split


def column_slabs(n, n_slabs, align=None, min_width=None):
    """[0, n) cut into at most ``n_slabs`` ranges whose starts are multiples of ``align`` columns
    and which are at least ``min_width`` wide (narrow matrices are not worth splitting)."""
    align = JTVEC_SLAB_ALIGN if align is None else align
    min_width = JTVEC_SLAB_MIN if min_width is None else min_width
    k = max(1, min(int(n_slabs), n // min_width))
    bounds = [0]
    for j in range(1, k):
        b = (n * j // k) // align * align
        if b > bounds[-1]:
            bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]


class _Zero:
    """Stand-in for ``simpeg.utils.Zero`` (derivative of a fixed property)."""

    def __matmul__(self, other):
        return 0

    __rmatmul__ = __matmul__

    @property
    def T(self):
        return self


class DeviceLinearOperator(LinearOperator):
    """
    ``scipy.sparse.linalg.LinearOperator`` whose matvec / rmatvec run on the
    GPU.  Returned for ``G`` by ``engine="cuda"``: with ``store_sensitivities=
    "ram"`` the products stream the device-resident ``G`` (which cannot come
    back to the host at production sizes), with ``"forward_only"`` they
    evaluate the kernels on the fly -- the same object the reference builds at
    gravity/simulation.py:673-688.  ``to_numpy()`` copies a stored ``G`` to the
    host for small problems.
    """

    def __init__(self, shape, matvec, rmatvec, to_numpy=None):
        self._mv, self._rmv, self._to_numpy = matvec, rmatvec, to_numpy
        super().__init__(dtype=np.float64, shape=shape)

    def _matvec(self, x):
        return self._mv(np.asarray(x, dtype=np.float64).reshape(-1))

    def _rmatvec(self, x):
        return self._rmv(np.asarray(x, dtype=np.float64).reshape(-1))

    # real operator: transpose == adjoint, without scipy's np.conj round trips
    def _adjoint(self):
        return DeviceLinearOperator(self.shape[::-1], self._rmv, self._mv)

    _transpose = _adjoint

    def to_numpy(self):
        if self._to_numpy is None:
            raise NotImplementedError(
                "G is never stored with store_sensitivities='forward_only'")
        return self._to_numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a if dtype is None else a.astype(dtype)


class ShardedDeviceOps:
    """
    Host-vector <-> device operations of one simulation: the CUDA engine of the
    local receiver shard plus the two exchanges of the path (all-gather of
    rows, all-reduce of column partials).  Holds no reference to the
    simulation object.
    """

    def __init__(self, engine, coll, sharding):
        self.engine, self.coll, self.sharding = engine, coll, sharding
        self._local_rows_t = engine.torch.from_numpy(sharding.local_rows).to(engine.device)

    # local <-> global vectors
    def rows_to_host(self, local_t):
        """local rows (device) -> full (nD,) numpy vector (all-gather over ranks)."""
        if self.coll.world_size == 1:
            return self.engine.to_host(local_t)
        return self.engine.to_host(self.coll.all_gather_rows(local_t, self.sharding))

    def cols_to_host(self, partial_t):
        """partial (nP,) device vector -> numpy (all-reduce over ranks)."""
        return self.engine.to_host(self.coll.all_reduce_sum(partial_t))

    def local_rows_of(self, vector):
        """full (nD,) host vector -> local rows on the device."""
        v = self.engine.to_device(vector, n=self.sharding.nD)
        if self.coll.world_size == 1:
            return v
        return v[self._local_rows_t].contiguous()

    def cols_to_device(self, vector):
        """Replicated (nP,) host vector -> device.  Every rank is handed the same vector (the
        model / Jvec argument of an SPMD run), so with several ranks each one uploads only its
        1/N slice over PCIe and the slices are all-gathered over NVLink: N times less H2D traffic
        per rank, which is what the host-vector Jvec of an 8-GPU run was waiting for."""
        eng, coll = self.engine, self.coll
        n = eng.nP
        vector = np.ascontiguousarray(vector, dtype=np.float64).reshape(-1)
        if vector.shape[0] != n:
            raise ValueError(f"expected a vector of length {n}, got {vector.shape[0]}")
        if coll.world_size == 1 or n < SLICED_UPLOAD_MIN:
            return eng.to_device(vector, n=n)
        return coll.all_gather_slices(lambda lo, hi: eng.to_device(vector[lo:hi]), n)

    # matrix-free
    def forward(self, model):
        eng = self.engine
        return self.rows_to_host(eng.forward_dev(self.cols_to_device(model)))

    def gt_dot_v(self, vector):
        return self.cols_to_host(self.engine.gt_dot_v_dev(self.local_rows_of(vector)))

    def gtg_diag(self, weights):
        return self.cols_to_host(self.engine.gtg_diag_dev(self.local_rows_of(weights)))

    # stored G
    def gemv(self, x):
        return self.rows_to_host(self.engine.gemv_dev(self.cols_to_device(x)))

    def gemv_t(self, y):
        """Host-vector Jtvec.  The product runs in column slabs; the all-reduce of a finished slab
        (several ranks) and its D2H transfer into the page-locked result run on NCCL's / a copy
        stream while the next slab still streams G, so only the last slab's reduction and transfer
        are exposed (the 24 MB result of BASELINE config 4 costs 2 ms per rank over PCIe when all
        eight ranks fetch it at once)."""
        eng, coll = self.engine, self.coll
        torch = eng.torch
        y_local = self.local_rows_of(y)
        n = eng.nP
        slabs = column_slabs(n, max(JTVEC_SLABS, 2))
        if len(slabs) == 1:
            return eng.to_host(self.gemv_t_allreduce_dev(y_local))
        with torch.cuda.device(eng.device):
            z = torch.empty(n, dtype=torch.float64, device=eng.device)
            out = torch.empty(n, dtype=torch.float64, pin_memory=True)
            copy_stream = eng.copy_stream()
            z.record_stream(copy_stream)
            for c0, c1 in slabs:
                eng.gemv_t_dev(y_local, out=z, cols=(c0, c1))
                if coll.world_size > 1:
                    done = coll.all_reduce_sum_async(z[c0:c1])
                else:
                    done = torch.cuda.Event()
                    done.record()
                with torch.cuda.stream(copy_stream):
                    if coll.world_size > 1:
                        done.wait()                       # the copy stream waits for the reduction
                    else:
                        copy_stream.wait_event(done)
                    out[c0:c1].copy_(z[c0:c1], non_blocking=True)
            copy_stream.synchronize()
        return out.numpy()

    def gemv_t_allreduce_dev(self, y_local_t, out=None):
        """z = sum over ranks of G_local^T y_local, on the device (the Jtvec of a sharded run).

        The column-owner kernel works on disjoint column slabs, so the all-reduce of a finished
        slab runs (NCCL, its own stream) while the next slab is still streaming G: only the last
        slab's reduction is exposed instead of the whole (nP,) vector (24 MB at BASELINE config 4).
        """
        eng, coll = self.engine, self.coll
        if coll.world_size == 1:
            return eng.gemv_t_dev(y_local_t, out=out)
        n = eng.nP
        z = eng.torch.empty(n, dtype=eng.torch.float64, device=eng.device) if out is None else out
        works = []
        for c0, c1 in column_slabs(n, JTVEC_SLABS):
            eng.gemv_t_dev(y_local_t, out=z, cols=(c0, c1))
            works.append(coll.all_reduce_sum_async(z[c0:c1]))
        for w in works:
            w.wait()                      # the compute stream waits for the reductions, not the host
        return z

    def G_to_numpy(self):
        if self.coll.world_size != 1:
            raise NotImplementedError("G.to_numpy() is only available on a single process")
        return self.engine.G_to_numpy()


class BasePFSimulation:
    _physics = None            # "gravity" | "magnetics"
    _prop_name = None          # "rho" | "chi"

    def __init__(self, mesh, active_cells=None, store_sensitivities="ram", n_processes=1,
                 sensitivity_dtype=np.float32, engine="cuda", numba_parallel=True,
                 survey=None, sensitivity_path=None, counter=None, verbose=False,
                 linear_model=None, model_map=None, G=None, comm=None, device=None,
                 **kwargs):
        if kwargs.pop("ind_active", None) is not None:
            raise TypeError("'ind_active' has been removed in SimPEG v0.24.0, "
                            "please use 'active_cells' instead.")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self.mesh = mesh
        self.store_sensitivities = store_sensitivities
        self.sensitivity_dtype = sensitivity_dtype
        self.engine = engine
        self.numba_parallel = bool(numba_parallel)  # accepted for API parity; unused by "cuda"
        self.survey = survey
        # BaseSimulation.__init__ (simpeg/simulation.py:62-80): default path, counter, verbose
        if sensitivity_path is None:
            sensitivity_path = os.path.join(".", "sensitivity")
        self.sensitivity_path = validate_string("sensitivity_path", sensitivity_path,
                                                case_sensitive=True)
        self.counter = counter          # kept for API parity (nothing on this path counts)
        self.verbose = bool(verbose)
        # LinearSimulation.__init__ (simpeg/simulation.py:655-668): linear_model / model_map are
        # the generic names of rho|chi / rhoMap|chiMap; a user-supplied G replaces the kernel
        # evaluation (uploaded to HBM on first use, see the G property)
        self.linear_model, self.model_map = linear_model, model_map
        self._G_source = G
        self.n_processes = n_processes
        n_cells = mesh.n_cells
        if active_cells is None:
            active_cells = np.ones(n_cells, dtype=bool)
        else:
            active_cells = validate_active_indices("active_cells", active_cells, n_cells)
        self._active_cells = active_cells
        self.nC = int(active_cells.sum())
        self._comm = comm
        self._device = device
        self._model = None
        self._engine_obj = None
        self._active_nodes_cache = None

    # ------------------------------------------------------------ properties
    @property
    def mesh(self):
        return self._mesh

    @mesh.setter
    def mesh(self, value):
        if not hasattr(value, "dim"):
            raise TypeError("mesh must be a 3D tensor or tree mesh (missing 'dim')")
        if value.dim != 3:
            raise ValueError(
                f"{type(self).__name__} mesh must be 3D, received a {value.dim}D mesh.")
        for attr in ("n_cells", "cell_nodes"):
            if not hasattr(value, attr):
                raise TypeError(f"mesh must be a 3D tensor or tree mesh (missing {attr!r})")
        self._mesh = value

    @property
    def active_cells(self):
        return self._active_cells

    @property
    def store_sensitivities(self):
        return self._store_sensitivities

    @store_sensitivities.setter
    def store_sensitivities(self, value):
        self._store_sensitivities = validate_string(
            "store_sensitivities", value, ["disk", "ram", "forward_only"])

    @property
    def sensitivity_dtype(self):
        if self.store_sensitivities == "forward_only":
            return np.float64
        return self._sensitivity_dtype

    @sensitivity_dtype.setter
    def sensitivity_dtype(self, value):
        if value is not np.float32 and value is not np.float64:
            raise TypeError("sensitivity_dtype must be either np.float32 or np.float64.")
        self._sensitivity_dtype = value

    @property
    def engine(self):
        return self._engine

    @engine.setter
    def engine(self, value):
        validate_string("engine", value, string_list=ENGINES, case_sensitive=True)
        if value != "cuda":
            raise NotImplementedError(
                f"engine={value!r} is the reference's CPU path; this package only "
                "provides engine='cuda'.")
        self._engine = value

    @property
    def n_processes(self):
        return self._n_processes

    @n_processes.setter
    def n_processes(self, value):
        if value is not None:
            if int(value) != value or value < 1:
                raise ValueError("'n_processes' must be an integer >= 1")
        self._n_processes = value

    # ---------------------------------------------------- model plumbing (props)
    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, value):
        """props.py:457-472: always takes the new values -- the caller may have changed the same
        array object in place, and ``_model`` only aliases its memory for contiguous float64
        input.  Whether the values changed matters only for the caches listed in
        ``_delete_on_model_update`` (amplitude data); the comparison is made against a private
        copy of the values those caches were computed for and skipped when there are none."""
        if value is not None:
            value = np.asarray(value, dtype=np.float64).reshape(-1)
        self._model = value
        deletes = self._delete_on_model_update
        if not deletes:
            return
        seen = getattr(self, "_model_seen", None)
        if value is not None and seen is not None and seen.shape == value.shape \
                and np.array_equal(seen, value):
            return
        self._model_seen = None if value is None else value.copy()
        for name in deletes:
            if hasattr(self, name):
                delattr(self, name)

    @property
    def _delete_on_model_update(self):
        return []

    def _physical_property(self, fixed, mapping):
        if mapping is not None:
            if self._model is None:
                raise AttributeError(f"no model set: cannot evaluate {self._prop_name}Map * model")
            return np.asarray(mapping * self._model, dtype=np.float64)
        return fixed

    def _property_deriv(self, mapping):
        if mapping is None or self._model is None:
            return _Zero()
        d = mapping.deriv(self._model)
        return d if sp.issparse(d) else sp.csr_matrix(d)

    @staticmethod
    def _jtj_diagonal(gtg_diagonal, deriv):
        """``((sdiag(sqrt(gtg)) @ deriv).power(2)).sum(axis=0)`` (gravity/simulation.py:279-283):
        diag(J^T W^T W J) from diag(G^T W^T W G) and the model-map derivative.  An identity
        derivative returns the diagonal itself (the sparse products cost more than the 98 GB
        column sums they follow)."""
        if getattr(deriv, "is_identity", False):
            return np.array(gtg_diagonal, dtype=np.float64)
        from ..utils import mkvc, sdiag

        return mkvc((sdiag(np.sqrt(gtg_diagonal)) @ deriv).power(2).sum(axis=0))

    @staticmethod
    def _deriv_dot(deriv, v):
        """deriv @ v, skipping the sparse product for an identity mapping."""
        if getattr(deriv, "is_identity", False):
            return np.asarray(v, dtype=np.float64)
        return np.asarray(deriv @ v, dtype=np.float64)

    @staticmethod
    def _deriv_t_dot(deriv, v):
        """deriv.T @ v"""
        if getattr(deriv, "is_identity", False):
            return np.asarray(v)
        return np.asarray(deriv.T @ v)

    # ------------------------------------------------------------ mesh -> arrays
    def _get_active_nodes(self):
        """
        base.py:371-399: nodes of the active cells and the (nC, 8) table of
        indices into them.  Cached (the reference recomputes it per call).
        """
        if self._active_nodes_cache is not None:
            return self._active_nodes_cache
        mesh = self.mesh
        if hasattr(mesh, "total_nodes"):
            nodes = mesh.total_nodes          # TreeMesh: hanging nodes included
        elif hasattr(mesh, "nodes"):
            nodes = mesh.nodes
        else:
            raise TypeError(f"Invalid mesh of type {mesh.__class__.__name__}.")
        cell_nodes = np.asarray(mesh.cell_nodes)
        if self.nC != mesh.n_cells:
            cell_nodes = cell_nodes[self.active_cells]
            unique_nodes, active_cell_nodes = np.unique(cell_nodes, return_inverse=True)
            nodes = np.asarray(nodes)[unique_nodes]
            cell_nodes = active_cell_nodes.reshape(cell_nodes.shape)
        self._active_nodes_cache = (np.asarray(nodes, dtype=np.float64), cell_nodes)
        return self._active_nodes_cache

    def _tile_plan(self):
        """Tile plan of the active cells: native O(n) planner for tensor meshes,
        the generic (nodes, cell_nodes) planner otherwise (tree meshes)."""
        from ..tiling import build_tile_plan, build_tile_plan_tensor, tensor_axes_of

        axes = tensor_axes_of(self.mesh)
        if axes is not None:
            active = None if self.nC == self.mesh.n_cells else np.nonzero(self.active_cells)[0]
            return build_tile_plan_tensor(axes, active, with_tiles=False)
        # tree meshes go straight to the bisection (the brick test costs more than the plan)
        method = "rcb" if hasattr(self.mesh, "total_nodes") else "auto"
        return self._cached_generic_plan(lambda: build_tile_plan(*self._get_active_nodes(), method))

    def _cached_generic_plan(self, build):
        """The generic (nodes, cell_nodes) planner costs seconds per million cells and depends only
        on the mesh and the active cells: keep its result on the mesh object, so that further
        simulations on the same mesh (other surveys, components, store modes) reuse it."""
        import hashlib

        key = hashlib.sha1(np.packbits(self.active_cells).tobytes()).hexdigest()
        key += "|" + getattr(self, "_plan_cache_tag", "")
        cache = getattr(self.mesh, "_pf_tile_plans", None)
        if cache is None:
            cache = {}
            try:
                self.mesh._pf_tile_plans = cache
            except (AttributeError, TypeError):       # meshes that refuse new attributes
                pass
        if key not in cache:
            cache.clear()                              # one plan per mesh is enough
            cache[key] = build()
        return cache[key]

    def _get_components_and_receivers(self):
        """base.py:401-408"""
        if not hasattr(self.survey, "source_field"):
            raise AttributeError(f"The survey '{self.survey}' has no 'source_field' attribute.")
        for receiver_object in self.survey.source_field.receiver_list:
            yield receiver_object.components, receiver_object.locations

    # ------------------------------------------------------------ device engine
    def _n_columns(self):
        return self.nC

    def _engine_kwargs(self):
        return {}

    @property
    def _ops(self):
        """Lazily created device operations (engine + sharding) of this simulation."""
        if self._engine_obj is None:
            from ..engine import CudaEngine
            from ..sharding import Collectives, ReceiverSharding

            coll = self._comm if self._comm is not None else Collectives()
            receiver_list = [(np.asarray(loc, dtype=np.float64), list(comps))
                             for comps, loc in self._get_components_and_receivers()]
            sharding = ReceiverSharding(receiver_list, coll.rank, coll.world_size)
            engine = CudaEngine(
                self._physics, None, None, sharding.local_receiver_list,
                sensitivity_dtype=self.sensitivity_dtype, device=self._device,
                plan=self._tile_plan(), **self._engine_kwargs())
            self._engine_obj = ShardedDeviceOps(engine, coll, sharding)
        return self._engine_obj

    @property
    def _dev(self):
        return self._ops.engine

    @property
    def _coll(self):
        return self._ops.coll

    @property
    def _sharding(self):
        return self._ops.sharding

    def _rows_to_host(self, local_t):
        return self._ops.rows_to_host(local_t)

    def _cols_to_host(self, partial_t):
        return self._ops.cols_to_host(partial_t)

    def _local_rows_of(self, vector):
        return self._ops.local_rows_of(vector)

    # the four private methods of the reference, on the GPU -------------------------
    def _forward(self, model):
        """G @ model without storing G (gravity/simulation.py:541-583, mag 829-929)."""
        return self._ops.forward(model)

    def _sensitivity_matrix(self):
        """Build the local block of G on the device (gravity/simulation.py:585-629).  With
        ``store_sensitivities="disk"`` the matrix is also written to ``sensitivity_path`` in the
        reference's format (lines 599-606: a raw row-major ``np.memmap`` of ``sensitivity_dtype``,
        shape ``(nD, nP)``); products keep using the copy in HBM."""
        G = self._dev.build_G()
        if self.store_sensitivities == "disk":
            self._write_sensitivities()
        return G

    def _write_sensitivities(self):
        """Every rank writes its rows into the one file (rank 0 creates it).  A ``sensitivity_path``
        that ends in ``.zarr`` selects an uncompressed zarr (format 2) directory store in row chunks
        -- what the reference's dask path keeps (``simpeg/dask/potential_fields/base.py:46-66``) and
        ``dask.array.from_zarr`` reads back -- instead of the raw memmap of
        gravity/simulation.py:599-606."""
        ops = self._ops
        coll, rows = ops.coll, ops.sharding.local_rows
        shape = (self.survey.nD, self._n_columns())
        if zarr_store.wants_zarr(self.sensitivity_path):
            writer = zarr_store.ZarrMatrixWriter(self.sensitivity_path, shape, self.sensitivity_dtype,
                                                 create=coll.rank == 0)
            coll.barrier()
            for a, b, block in ops.engine.G_row_chunks():
                writer.write_rows(rows[a:b], block)
            coll.barrier()
            return
        if coll.rank == 0:
            np.memmap(self.sensitivity_path, shape=shape, dtype=self.sensitivity_dtype, order="C",
                      mode="w+").flush()
        coll.barrier()
        out = np.memmap(self.sensitivity_path, shape=shape, dtype=self.sensitivity_dtype, order="C",
                        mode="r+")
        for a, b, block in ops.engine.G_row_chunks():
            out[rows[a:b]] = block
        out.flush()
        del out
        coll.barrier()

    def load_sensitivities(self, path=None):
        """
        Use a sensitivity matrix that is already on disk -- written by this engine or by the
        reference with ``store_sensitivities="disk"`` -- instead of evaluating it: every rank uploads
        its rows to HBM.  ``path`` is the raw row-major memmap of gravity/simulation.py:599-606, a
        zarr (format 2) store, or a directory that holds ``sensitivity.zarr`` (the dask path's
        layout, ``simpeg/dask/potential_fields/base.py:46-48``).
        """
        path = self.sensitivity_path if path is None else path
        shape = (self.survey.nD, self._n_columns())
        nested = os.path.join(str(path), "sensitivity.zarr")
        if zarr_store.is_zarr_store(path) or zarr_store.is_zarr_store(nested):
            source = zarr_store.ZarrMatrixReader(path if zarr_store.is_zarr_store(path) else nested)
            if source.shape != shape:
                raise ValueError(f"zarr store holds a {source.shape} array, the simulation needs {shape}")
        else:
            source = np.memmap(path, shape=shape, dtype=self.sensitivity_dtype, order="C", mode="r")
        ops = self._ops
        ops.engine.load_G_rows(source, ops.sharding.local_rows)
        del source
        if hasattr(self, "_G"):
            del self._G
        self._G_loaded = True
        return self.G

    def _sensitivity_matrix_transpose_dot_vec(self, vector):
        """G.T @ v without storing G (gravity/simulation.py:631-671)."""
        return self._ops.gt_dot_v(vector)

    def _gtg_diagonal_without_building_g(self, weights):
        """diag(G.T W.T W G) without storing G (gravity/simulation.py:690-724)."""
        return self._ops.gtg_diag(weights)

    def _sensitivity_matrix_as_operator(self):
        """gravity/simulation.py:673-688"""
        ops = self._ops
        shape = (self.survey.nD, self._n_columns())
        return DeviceLinearOperator(shape, ops.forward, ops.gt_dot_v)

    @property
    def G(self):
        """
        Forward operator.  ``engine="cuda"`` always returns a
        :class:`DeviceLinearOperator`; for 'ram' the matrix lives in HBM and is
        built on first access (gravity/simulation.py:421-443).  The operator
        only references the device engine, not the simulation, so dropping both
        releases the HBM immediately (no reference cycle).
        """
        if not hasattr(self, "_G"):
            if self._G_source is not None:
                # LinearSimulation(G=...) / ``sim.G = array`` (simulation.py:664-665, 684-688):
                # the caller's matrix goes to HBM instead of being evaluated
                source = np.asarray(self._G_source)
                shape = (self.survey.nD, self._n_columns())
                if source.shape != shape:
                    raise ValueError(f"G must have shape {shape}, got {source.shape}")
                ops = self._ops
                ops.engine.load_G_rows(source, ops.sharding.local_rows)
                self._G_loaded = True
            if self.store_sensitivities == "forward_only" and self._G_source is None:
                self._G = self._sensitivity_matrix_as_operator()
            else:
                if not getattr(self, "_G_loaded", False):
                    self._sensitivity_matrix()
                ops = self._ops
                shape = (self.survey.nD, self._n_columns())
                self._G = DeviceLinearOperator(shape, ops.gemv, ops.gemv_t, ops.G_to_numpy)
        return self._G

    @G.setter
    def G(self, value):
        if hasattr(self, "_G"):
            del self._G
        self._G_source = value
        self._G_loaded = False

    def _jacobian(self, deriv):
        """``getJ`` of the reference (gravity/simulation.py:318-355): ``G @ deriv`` as a dense
        ndarray for 'ram' / 'disk' -- as long as it fits ``getJ_dense_max_bytes`` on the host and
        one process owns every row -- and as a LinearOperator otherwise ('forward_only', sharded
        runs, production sizes, where the matrix only exists in HBM)."""
        G = self.G
        nbytes = self.survey.nD * self._n_columns() * np.dtype(self.sensitivity_dtype).itemsize
        dense = (self.store_sensitivities != "forward_only" and self._coll.world_size == 1
                 and nbytes <= self.getJ_dense_max_bytes)
        if not dense:
            return G @ aslinearoperator(deriv)
        Gh = G.to_numpy()
        return Gh if getattr(deriv, "is_identity", False) else np.asarray(Gh @ deriv)

    getJ_dense_max_bytes = 1 << 30

    def _warn_n_processes(self, kwargs, what):
        if "n_processes" in kwargs:
            warnings.warn(
                f"The 'n_processes' will be ignored when selecting 'cuda' as the "
                f"engine in the {what} simulation.", UserWarning, stacklevel=3)
            self.n_processes = None

    # LinearSimulation.dpred (simulation.py:698-704)
    def dpred(self, m=None, f=None):
        if m is None:
            m = self.model
        if f is not None:
            return f
        return self.fields(m)


def layer_nodes_and_cells(cell_bounds, cell_z_top, cell_z_bottom):
    """
    The prisms of an equivalent-source layer as an unstructured hexahedral "mesh" for the
    node-sharing kernels: ``nodes (nN, 3)`` and ``cell_nodes (nC, 8)`` with the vertex order of
    base.py:431-441 (x fastest, then y, then bottom before top).  Vertices that coincide -- all of
    them in the horizontal directions when top and bottom are constant -- are stored once, so a
    flat layer costs ~2 node evaluations per (receiver, cell) pair instead of the 8 of the
    reference's per-cell loops (``_numba_utils.py:48-255``).
    """
    b = np.ascontiguousarray(cell_bounds, dtype=np.float64).reshape(-1, 4)
    n = b.shape[0]
    top = np.ascontiguousarray(cell_z_top, dtype=np.float64).reshape(n)
    bot = np.ascontiguousarray(cell_z_bottom, dtype=np.float64).reshape(n)
    x = np.stack([b[:, 0], b[:, 1], b[:, 0], b[:, 1]] * 2, axis=1)      # (n, 8)
    y = np.stack([b[:, 2], b[:, 2], b[:, 3], b[:, 3]] * 2, axis=1)
    z = np.stack([bot] * 4 + [top] * 4, axis=1)
    ux, ix = np.unique(x, return_inverse=True)
    uy, iy = np.unique(y, return_inverse=True)
    uz, iz = np.unique(z, return_inverse=True)
    key = (iz.reshape(-1).astype(np.int64) * len(uy) + iy.reshape(-1)) * len(ux) + ix.reshape(-1)
    ukey, inv = np.unique(key, return_inverse=True)
    kx = ukey % len(ux)
    ky = (ukey // len(ux)) % len(uy)
    kz = ukey // (len(ux) * len(uy))
    nodes = np.ascontiguousarray(np.stack([ux[kx], uy[ky], uz[kz]], axis=1))
    return nodes, np.ascontiguousarray(inv.reshape(n, 8).astype(np.int64))


class BaseEquivalentSourceLayerSimulation(BasePFSimulation):
    """
    Equivalent-source layer: a 2D tensor / tree mesh plus the elevation of the top and bottom
    face of every active cell (simpeg/potential_fields/base.py:411-474).  With ``engine="cuda"``
    the layer's prisms go through the same device kernels as the 3D integral simulations.

    Parameters
    ----------
    mesh : 2D mesh exposing ``dim``, ``n_cells`` and ``cell_bounds`` (n_cells, 4)
    cell_z_top, cell_z_bottom : float or (n_active_cells,) array
    """

    def __init__(self, mesh, cell_z_top, cell_z_bottom, **kwargs):
        super().__init__(mesh=mesh, **kwargs)
        if isinstance(cell_z_top, (int, float)):
            cell_z_top = float(cell_z_top) * np.ones(self.nC)
        if isinstance(cell_z_bottom, (int, float)):
            cell_z_bottom = float(cell_z_bottom) * np.ones(self.nC)
        if (self.nC != len(cell_z_top)) | (self.nC != len(cell_z_bottom)):
            raise AttributeError(
                "'cell_z_top' and 'cell_z_bottom' must have length equal to number of",
                "cells, and match the number of active cells.",
            )
        self._cell_z_top = np.asarray(cell_z_top, dtype=np.float64)
        self._cell_z_bottom = np.asarray(cell_z_bottom, dtype=np.float64)

    @property
    def cell_z_top(self):
        """Elevations for the top face of all cells in the layer."""
        return self._cell_z_top

    @property
    def cell_z_bottom(self):
        """Elevations for the bottom face of all cells in the layer."""
        return self._cell_z_bottom

    def _set_layer_mesh(self, value):
        for attr in ("dim", "n_cells", "cell_bounds"):
            if not hasattr(value, attr):
                raise TypeError(f"mesh must be a 2D tensor or tree mesh (missing {attr!r})")
        if value.dim != 2:
            raise ValueError(
                f"{type(self).__name__} mesh must be 2D, received a {value.dim}D mesh.")
        self._mesh = value

    mesh = property(BasePFSimulation.mesh.fget, _set_layer_mesh)

    def _get_active_nodes(self):
        if self._active_nodes_cache is None:
            bounds = np.asarray(self.mesh.cell_bounds, dtype=np.float64)[self.active_cells]
            self._active_nodes_cache = layer_nodes_and_cells(bounds, self.cell_z_top,
                                                             self.cell_z_bottom)
        return self._active_nodes_cache

    def _tile_plan(self):
        """A flat layer on a tensor mesh IS a one-cell-thick 3D tensor mesh (z-marching kernel);
        anything else goes through the generic (nodes, cell_nodes) planner."""
        from ..tiling import build_tile_plan, build_tile_plan_tensor

        mesh, top, bot = self.mesh, self.cell_z_top, self.cell_z_bottom
        flat = self.nC > 0 and np.all(top == top[0]) and np.all(bot == bot[0]) and bot[0] < top[0]
        if flat and hasattr(mesh, "nodes_x") and hasattr(mesh, "nodes_y") \
                and not hasattr(mesh, "total_nodes"):
            xs = np.asarray(mesh.nodes_x, dtype=np.float64)
            ys = np.asarray(mesh.nodes_y, dtype=np.float64)
            if (len(xs) - 1) * (len(ys) - 1) == mesh.n_cells:
                axes = (xs, ys, np.array([bot[0], top[0]]))
                active = None if self.nC == mesh.n_cells else np.nonzero(self.active_cells)[0]
                return build_tile_plan_tensor(axes, active, with_tiles=False)
        import hashlib

        self._plan_cache_tag = "layer:" + hashlib.sha1(
            np.ascontiguousarray(top).tobytes() + np.ascontiguousarray(bot).tobytes()).hexdigest()
        method = "auto" if flat else "rcb"      # only a flat layer can be brick-like
        return self._cached_generic_plan(lambda: build_tile_plan(*self._get_active_nodes(), method))
