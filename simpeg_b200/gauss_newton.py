"""
Device-resident inner loop of the Gauss-Newton step for the stored-G simulations
(SURVEY.md 8(f) row f2): what the reference does on the host in

* ``ProjectedGNCG.findSearchDirection`` (simpeg/optimization.py:1625-1711): projected,
  preconditioned CG on ``H step = -g`` restricted to the free variables, plus the scaled
  gradient on the active-but-not-binding set;
* ``L2DataMisfit.deriv2`` (simpeg/data_misfit.py:341-343):
  ``2 * Jtvec(m, W * (W * Jvec(m, v)))`` -- one ``Jvec`` and one ``Jtvec`` per CG iteration, each
  moving numpy vectors through the simulation;
* ``UpdatePreconditioner`` (simpeg/directives/_directives.py:2205-2233):
  ``approxHinv = 1 / (JtJdiag + beta * diag(reg.deriv2))``.

Here ``p, r, q, h`` and the step live in HBM for the whole loop.  Per iteration: ``pf_gemv``
(local rows of ``G (d * p)``), row weights, ``pf_gemv_t``, ONE all-reduce of the (nP,) partial
(``W`` is diagonal, so the weighted rows never have to be gathered), then three fused vector
kernels (``csrc/pf_cg.cu``).  The host reads one scalar (``r.r``) per iteration for the stopping
test.  Arithmetic is fp64 throughout; dot products are reduced in a fixed order, so every rank of
a sharded run walks through identical iterates.

There is no CPU fallback: the class needs a simulation with ``engine="cuda"`` and
``store_sensitivities="ram"``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import scipy.sparse as sp

from . import _lib


class DeviceGaussNewtonCG:
    """
    Parameters
    ----------
    simulation : gravity / magnetics ``Simulation3DIntegral`` (or equivalent-source layer) with
        ``engine="cuda"``, ``store_sensitivities="ram"``; the derivative of its model map may be
        the identity, diagonal (exp, ...) or any sparse matrix (``TileMap``, ``Wires``, ``LinearMap``).
    data_weights : (nD,) array, optional
        Diagonal of ``W`` of the data misfit (1 / standard deviation); default ones.
    reg_hessian : None | (nP,) array | scipy.sparse matrix (nP, nP)
        Regularisation Hessian ``reg.deriv2(m)`` (without ``beta``): absent, diagonal or sparse.
    beta : float
        Trade-off parameter multiplying ``reg_hessian``.
    lower, upper : float or (nP,) array
        Bounds of ``ProjectedGNCG`` (optimization.py:1524-1525).
    cg_maxiter, cg_rtol, cg_atol, step_active_set, active_set_grad_scale
        As ``ProjectedGNCG`` (optimization.py:1526-1530; defaults are the reference's future
        defaults ``cg_rtol=1e-3, cg_atol=0``).
    """

    def __init__(self, simulation, data_weights=None, reg_hessian=None, beta=0.0,
                 lower=-np.inf, upper=np.inf, cg_maxiter=5, cg_rtol=1e-3, cg_atol=0.0,
                 step_active_set=True, active_set_grad_scale=1e-2):
        if getattr(simulation, "engine", None) != "cuda" or simulation.store_sensitivities \
                not in ("ram", "disk"):
            raise ValueError("DeviceGaussNewtonCG needs engine='cuda' and store_sensitivities='ram'")
        if getattr(simulation, "is_amplitude_data", False):
            raise NotImplementedError("amplitude data are not supported by the device CG loop")
        self.simulation = simulation
        self.ops = simulation._ops
        self.engine = self.ops.engine
        self.lib = self.engine.lib
        self.torch = self.engine.torch
        self.nP = simulation._n_columns()
        self.nD = simulation.survey.nD
        self.n_model = None
        self.beta = float(beta)
        self.lower, self.upper = lower, upper
        self.cg_maxiter, self.cg_rtol, self.cg_atol = int(cg_maxiter), float(cg_rtol), float(cg_atol)
        self.step_active_set = bool(step_active_set)
        self.active_set_grad_scale = float(active_set_grad_scale)
        self.cg_count, self.cg_abs_resid, self.cg_rel_resid = 0, np.inf, np.inf
        self._approx_hinv = self._minv_t = None
        torch, dev = self.torch, self.engine.device

        w = np.ones(self.nD) if data_weights is None else np.asarray(data_weights, dtype=np.float64)
        if w.shape != (self.nD,):
            raise ValueError(f"data_weights must have shape ({self.nD},)")
        self._w_host = w
        self._w2_local = self.ops.local_rows_of(w * w)          # W^T W on the rows of this rank

        self._reg = reg_hessian
        self._rdiag = self._indptr = self._indices = self._values = None
        self._deriv_t = None
        self._scal = torch.zeros(8, dtype=torch.float64, device=dev)
        self._ws = torch.empty(self.lib.pf_cg_workspace_bytes(0) // 8, dtype=torch.float64, device=dev)

    @property
    def approxHinv(self):
        """Diagonal of the preconditioner (``opt.approxHinv`` of the reference), or None."""
        return self._approx_hinv

    @approxHinv.setter
    def approxHinv(self, value):
        self._approx_hinv = None if value is None else np.asarray(value, dtype=np.float64).reshape(-1)
        self._minv_t = None                   # uploaded once, at the next solve

    # ------------------------------------------------------------------ set-up
    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def _upload_reg(self, n):
        reg = self._reg
        torch, dev = self.torch, self.engine.device
        if reg is None:
            return
        if sp.issparse(reg):
            R = sp.csr_matrix(reg)
            if R.shape != (n, n):
                raise ValueError(f"reg_hessian must be ({n}, {n}), got {R.shape}")
            if R.nnz >= 2**31:
                raise ValueError("reg_hessian has too many non-zeros for int32 indices")
            R.sort_indices()
            self._indptr = torch.from_numpy(R.indptr.astype(np.int32)).to(dev)
            self._indices = torch.from_numpy(R.indices.astype(np.int32)).to(dev)
            self._values = torch.from_numpy(R.data.astype(np.float64)).to(dev)
            self._reg_diag_host = np.asarray(R.diagonal(), dtype=np.float64)
        else:
            d = np.asarray(reg, dtype=np.float64).reshape(-1)
            if d.shape != (n,):
                raise ValueError(f"diagonal reg_hessian must have {n} entries")
            self._rdiag = torch.from_numpy(np.ascontiguousarray(d)).to(dev)
            self._reg_diag_host = d

    def set_model(self, m):
        """Linearisation point: fixes the derivative of the model map -- the identity, a diagonal
        (``ExpMap`` ...) or a general sparse matrix (``maps.TileMap``, ``maps.Wires``,
        ``LinearMap``: applied as device CSR products around the two products over G)."""
        sim = self.simulation
        sim.model = m
        deriv = sim.rhoDeriv if sim._physics == "gravity" else sim.chiDeriv
        n = int(np.asarray(m).size)
        torch, dev = self.torch, self.engine.device
        self._deriv_t = self._deriv_host = self._csr = self._csr_t = None
        if not getattr(deriv, "is_identity", False):
            D = sp.csr_matrix(deriv)
            if D.shape != (self.nP, n):
                raise ValueError(f"the model-map derivative must be ({self.nP}, {n}), got {D.shape}")
            if self.nP == n and (D - sp.diags(D.diagonal())).nnz == 0:
                self._deriv_host = np.asarray(D.diagonal(), dtype=np.float64)
                self._deriv_t = torch.from_numpy(self._deriv_host).to(dev)
            else:
                if D.nnz >= 2**31:
                    raise ValueError("model-map derivative has too many non-zeros for int32 indices")

                def upload(A):
                    A = sp.csr_matrix(A)
                    A.sort_indices()
                    return (torch.from_numpy(A.indptr.astype(np.int32)).to(dev),
                            torch.from_numpy(A.indices.astype(np.int32)).to(dev),
                            torch.from_numpy(A.data.astype(np.float64)).to(dev))

                self._csr, self._csr_t = upload(D), upload(D.T)
        if self.n_model != n:
            self.n_model = n
            self._upload_reg(n)
        self._m = np.asarray(m, dtype=np.float64)
        sim.G     # make sure the matrix is resident
        return self

    def jacobi_preconditioner(self):
        """``UpdatePreconditioner.initialize`` (directives/_directives.py:2205-2228):
        ``approxHinv = 1 / (getJtJdiag(m, W) + beta * diag(reg.deriv2))`` (zeros kept)."""
        diag = self.simulation.getJtJdiag(self._m, W=sp.diags(self._w_host))
        if self._reg is not None:
            diag = diag + self.beta * self._reg_diag_host
        out = np.zeros_like(diag)
        nz = diag != 0
        out[nz] = 1.0 / diag[nz]
        self.approxHinv = out
        return out

    # ----------------------------------------------------------------- products
    def _spmv(self, csr, n_rows, x_t):
        y = self.torch.empty(n_rows, dtype=self.torch.float64, device=self.engine.device)
        _lib.check(self.lib.pf_csr_spmv(csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr(), n_rows,
                                        x_t.data_ptr(), y.data_ptr(), self.engine._stream()), "pf_csr_spmv")
        return y

    def _normal_product(self, p_t):
        """z = D^T G^T W^2 G D p summed over ranks (device tensors; D = model-map derivative:
        identity, diagonal -- its transpose is applied inside pf_cg_hvp_combine -- or CSR)."""
        eng, lib = self.engine, self.lib
        stream = eng._stream()
        torch = self.torch
        x = p_t
        if self._deriv_t is not None:
            x = torch.empty_like(p_t)
            _lib.check(lib.pf_vec_mul(self._deriv_t.data_ptr(), p_t.data_ptr(), self.nP, x.data_ptr(),
                                      stream), "pf_vec_mul")
        elif self._csr is not None:
            x = self._spmv(self._csr, self.nP, p_t)
        y = eng.gemv_dev(x)
        _lib.check(lib.pf_vec_mul(y.data_ptr(), self._w2_local.data_ptr(), y.numel(), y.data_ptr(),
                                  stream), "pf_vec_mul")
        z = self.ops.gemv_t_allreduce_dev(y)        # column slabs overlapped with their all-reduce
        if self._csr_t is not None:
            z = self._spmv(self._csr_t, self.n_model, z)
        return z

    def _hvp_combine(self, z, p, q, mask):
        eng = self.engine
        _lib.check(self.lib.pf_cg_hvp_combine(
            z.data_ptr(), self._ptr(self._deriv_t), p.data_ptr(), self._ptr(self._indptr),
            self._ptr(self._indices), self._ptr(self._values), self._ptr(self._rdiag),
            ctypes.c_double(2.0), ctypes.c_double(self.beta), self._ptr(mask), self.n_model, q.data_ptr(),
            self._scal.data_ptr(), self._ws.data_ptr(), self._ws.numel() * 8, eng._stream()),
            "pf_cg_hvp_combine")

    def _model_to_device(self, v):
        """(n_model,) host vector -> device (rank-sliced upload when it has nP entries)."""
        if self.n_model == self.nP:
            return self.ops.cols_to_device(v)
        return self.engine.to_device(v, n=self.n_model)

    def hessian_vector(self, v):
        """``H v = 2 J^T W^T W J v + beta R v`` for a host vector (device round trip; tests)."""
        eng = self.engine
        p = self._model_to_device(v)
        q = self.torch.empty_like(p)
        with self.torch.cuda.device(eng.device):
            self._hvp_combine(self._normal_product(p), p, q, None)
        return eng.to_host(q)

    # ------------------------------------------------------------------- solve
    def _bounds_sets(self, x, g):
        """activeSet / bindingSet of optimization.py:981-1018."""
        active = np.zeros(x.shape, dtype=bool)
        bound = np.zeros(x.shape, dtype=bool)
        if self.lower is not None:
            active |= x <= self.lower
            bound |= (x <= self.lower) & (g >= 0)
        if self.upper is not None:
            active |= x >= self.upper
            bound |= (x >= self.upper) & (g <= 0)
        return active, bound

    def find_search_direction(self, g, xc=None):
        """
        ``ProjectedGNCG.findSearchDirection`` (optimization.py:1625-1711) with ``self.g = g``,
        ``self.xc = xc`` (default: the model given to :meth:`set_model`) and
        ``self.H = data-misfit + beta * regularisation`` Hessians.  Returns the step (numpy).
        """
        if self.n_model is None:
            raise RuntimeError("call set_model(m) first")
        eng, lib, torch = self.engine, self.lib, self.torch
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1)
        xc = self._m if xc is None else np.asarray(xc, dtype=np.float64).reshape(-1)
        n = self.n_model
        if g.shape != (n,) or xc.shape != (n,):
            raise ValueError(f"gradient and model must have {n} entries")
        # no finite bound: every variable is free (activeSet / bindingSet of optimization.py:981-1018
        # are empty) -- skip the six passes over (nP,) host vectors they cost
        unbounded = (self.lower is None or np.all(np.isneginf(self.lower))) and \
            (self.upper is None or np.all(np.isposinf(self.upper)))
        if unbounded:
            active = bound = inactive = None
        else:
            active, bound = self._bounds_sets(xc, g)
            inactive = ~active
        ws, wsb = self._ws.data_ptr(), self._ws.numel() * 8
        with torch.cuda.device(eng.device):
            stream = eng._stream()
            g_t = self._model_to_device(g)
            mask = None if (unbounded or inactive.all()) else \
                torch.from_numpy(inactive.astype(np.uint8)).to(eng.device)
            if self._approx_hinv is not None and self._minv_t is None:
                self._minv_t = self._model_to_device(self._approx_hinv)
            minv = self._minv_t
            step, r, p, q, h = (torch.empty(n, dtype=torch.float64, device=eng.device) for _ in range(5))
            scal = self._scal
            _lib.check(lib.pf_cg_init(g_t.data_ptr(), self._ptr(mask), self._ptr(minv), n, step.data_ptr(),
                                      r.data_ptr(), p.data_ptr(), scal.data_ptr(), ws, wsb, stream),
                       "pf_cg_init")
            r_norm0 = float(np.sqrt(scal[2].item()))
            atol = max(self.cg_rtol * r_norm0, self.cg_atol)
            r_norm, count = r_norm0, 0
            while r_norm > atol and count < self.cg_maxiter:
                count += 1
                z = self._normal_product(p)
                self._hvp_combine(z, p, q, mask)
                _lib.check(lib.pf_cg_update(p.data_ptr(), q.data_ptr(), self._ptr(minv), n,
                                            step.data_ptr(), r.data_ptr(), h.data_ptr(), scal.data_ptr(),
                                            ws, wsb, stream), "pf_cg_update")
                _lib.check(lib.pf_cg_direction(h.data_ptr(), n, p.data_ptr(), scal.data_ptr(), stream),
                           "pf_cg_direction")
                r_norm = float(np.sqrt(scal[2].item()))
            step_host = eng.to_host(step)
        self.cg_count, self.cg_abs_resid = count, r_norm
        self.cg_rel_resid = r_norm / r_norm0 if r_norm0 else 0.0

        # optimization.py:1688-1709: gradient on the active, non-binding set; zero the binding set
        if unbounded:
            return step_host
        active_not_bound = active & (~bound)
        if self.step_active_set and np.any(active_not_bound):
            rhs_a = active_not_bound * -g
            dm_i = max(abs(step_host))
            dm_a = max(abs(rhs_a))
            step_host += self.active_set_grad_scale * (rhs_a * dm_i / dm_a)
        step_host[bound] = 0
        return step_host
