// pf_cg.cu -- device-resident inner loop of the Gauss-Newton step.
//
// The reference solves H step = -g with preconditioned, projected CG on the host
// (simpeg/optimization.py:1625-1711 ProjectedGNCG.findSearchDirection), where one application
// of H is  2 J^T W^T W J p  (simpeg/data_misfit.py:341-343 L2DataMisfit.deriv2) plus the
// regularisation Hessian times beta, and every Jvec / Jtvec call moves (nP,) / (nD,) numpy
// vectors through the simulation.  Here p, r, q, h and the step stay in HBM for the whole
// loop; the two products over G are pf_gemv / pf_gemv_t (pf_products.cu) and the kernels
// below do everything between them: the regularisation term (CSR or diagonal), the model-map
// chain rule for diagonal derivatives, the active-set mask, the vector updates and all dot
// products (block partials combined in a fixed order by one CTA: deterministic, and identical
// on every rank of a sharded run).  CG scalars live in a small device array:
//
//     scal[0] = sold = r.h (previous)   scal[1] = p.q   scal[2] = r.r   scal[3] = snew = r.h
//
// so that alpha = sold / p.q and snew / sold are formed on the device; the host only reads
// r.r once per iteration for the stopping test.
#include "pf_common.cuh"

namespace {

constexpr int CG_THREADS = 256;
constexpr int CG_MAX_BLOCKS = 2048;   // workspace bound; launches use 8 CTAs per SM of the device
enum { S_SOLD = 0, S_PQ = 1, S_RR = 2, S_SNEW = 3 };

int cg_blocks(int64_t n)
{
    int64_t b = pf_ceil_div(n > 0 ? n : 1, CG_THREADS);
    int64_t cap = (int64_t)pf_sm_count() * 8;
    if (cap > CG_MAX_BLOCKS) cap = CG_MAX_BLOCKS;
    return (int)(b < cap ? b : cap);
}

// block sum of two quantities; thread 0 writes partial[blockIdx.x] and partial[gridDim.x + blockIdx.x]
__device__ __forceinline__ void block_sum2(double a, double b, double *partial)
{
    __shared__ double sa[CG_THREADS / 32], sb[CG_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int w = 0; w < CG_THREADS / 32; ++w) { ta += sa[w]; tb += sb[w]; }
        partial[blockIdx.x] = ta;
        partial[gridDim.x + blockIdx.x] = tb;
    }
}

// scal[ia] = sum partial[0..nb), scal[ib] = sum partial[nb..2nb)   (ib < 0: one quantity)
__global__ void __launch_bounds__(CG_THREADS) cg_finish_kernel(const double *partial, int nb, double *scal,
                                                               int ia, int ib)
{
    __shared__ double sa[CG_THREADS], sb[CG_THREADS];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nb; i += CG_THREADS) {
        a += partial[i];
        if (ib >= 0) b += partial[nb + i];
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = CG_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sa[threadIdx.x] += sa[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        scal[ia] = sa[0];
        if (ib >= 0) scal[ib] = sb[0];
    }
}

// optimization.py:1638-1652: step = 0; r = inactive * (-g); p = approxHinv * r; sold = r.p
__global__ void __launch_bounds__(CG_THREADS) cg_init_kernel(const double *__restrict__ g,
                                                             const uint8_t *__restrict__ inactive,
                                                             const double *__restrict__ minv, int64_t n,
                                                             double *__restrict__ step, double *__restrict__ r,
                                                             double *__restrict__ p, double *partial)
{
    double rr = 0.0, rp = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
        const double ri = (inactive == nullptr || inactive[i]) ? -g[i] : 0.0;
        const double pi = (minv ? minv[i] : 1.0) * ri;
        step[i] = 0.0;
        r[i] = ri;
        p[i] = pi;
        rr = fma(ri, ri, rr);
        rp = fma(ri, pi, rp);
    }
    block_sum2(rr, rp, partial);
}

// q = inactive * ( scale * d * z + beta * (R p) ),  partial p.q
//   z = G^T W^2 G (d * p) summed over ranks (data_misfit.py:341-343, scale = 2),
//   d = diagonal of the model-map derivative (nullptr: identity map),
//   R = regularisation Hessian as CSR (indptr/indices/values) or as a diagonal (rdiag) or absent.
__global__ void __launch_bounds__(CG_THREADS) cg_hvp_kernel(
    const double *__restrict__ z, const double *__restrict__ d, const double *__restrict__ p,
    const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
    const double *__restrict__ values, const double *__restrict__ rdiag, double scale, double beta,
    const uint8_t *__restrict__ inactive, int64_t n, double *__restrict__ q, double *partial)
{
    double pq = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
        double rp = 0.0;
        if (indptr) {
            const int32_t k1 = indptr[i + 1];
            for (int32_t k = indptr[i]; k < k1; ++k) rp = fma(values[k], __ldg(p + indices[k]), rp);
        } else if (rdiag) {
            rp = rdiag[i] * p[i];
        }
        double qi = fma(beta, rp, scale * (d ? d[i] : 1.0) * z[i]);
        if (inactive && !inactive[i]) qi = 0.0;
        q[i] = qi;
        pq = fma(p[i], qi, pq);
    }
    block_sum2(pq, 0.0, partial);
}

// optimization.py:1666-1676: alpha = sold / p.q; step += alpha p; r -= alpha q; h = approxHinv * r
__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(const double *__restrict__ p,
                                                               const double *__restrict__ q,
                                                               const double *__restrict__ minv, int64_t n,
                                                               const double *__restrict__ scal,
                                                               double *__restrict__ step, double *__restrict__ r,
                                                               double *__restrict__ h, double *partial)
{
    const double alpha = scal[S_SOLD] / scal[S_PQ];
    double rr = 0.0, rh = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
        step[i] = fma(alpha, p[i], step[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        const double hi = (minv ? minv[i] : 1.0) * ri;
        r[i] = ri;
        h[i] = hi;
        rr = fma(ri, ri, rr);
        rh = fma(ri, hi, rh);
    }
    block_sum2(rr, rh, partial);
}

// optimization.py:1678-1682: p = h + (snew / sold) p
__global__ void __launch_bounds__(CG_THREADS) cg_direction_kernel(const double *__restrict__ h, int64_t n,
                                                                  const double *__restrict__ scal,
                                                                  double *__restrict__ p)
{
    const double b = scal[S_SNEW] / scal[S_SOLD];
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS)
        p[i] = fma(b, p[i], h[i]);
}

__global__ void cg_commit_kernel(double *scal) { scal[S_SOLD] = scal[S_SNEW]; }

__global__ void __launch_bounds__(CG_THREADS) vec_mul_kernel(const double *__restrict__ a,
                                                             const double *__restrict__ b, int64_t n,
                                                             double *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS)
        out[i] = a[i] * b[i];
}

// y = A x for a CSR matrix (one thread per row): the model-map derivative of a TileMap / Wires /
// LinearMap mapping and its transpose around the two products over G
__global__ void __launch_bounds__(CG_THREADS) csr_spmv_kernel(const int32_t *__restrict__ indptr,
                                                              const int32_t *__restrict__ indices,
                                                              const double *__restrict__ values, int64_t n_rows,
                                                              const double *__restrict__ x, double *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * CG_THREADS) {
        double t = 0.0;
        const int32_t k1 = indptr[i + 1];
        for (int32_t k = indptr[i]; k < k1; ++k) t = fma(values[k], __ldg(x + indices[k]), t);
        y[i] = t;
    }
}

int check_ws(const void *ws, int64_t bytes, int64_t n)
{
    if (!ws || bytes < pf_cg_workspace_bytes(n)) {
        pf_set_error("pf_cg: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    return PF_OK;
}

}  // namespace

extern "C" int64_t pf_cg_workspace_bytes(int64_t n)
{
    (void)n;
    return 2 * (int64_t)CG_MAX_BLOCKS * (int64_t)sizeof(double);
}

extern "C" int pf_cg_init(const double *g, const uint8_t *inactive, const double *minv, int64_t n,
                          double *step, double *r, double *p, double *scal, void *workspace,
                          int64_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n >= 0, "negative length");
    PF_CHECK_ARG(scal != nullptr, "null scalars");
    PF_CHECK_ARG(n == 0 || (g && step && r && p), "null vector");
    int st = check_ws(workspace, workspace_bytes, n);
    if (st != PF_OK) return st;
    const int nb = cg_blocks(n);
    cg_init_kernel<<<nb, CG_THREADS, 0, stream>>>(g, inactive, minv, n, step, r, p, (double *)workspace);
    PF_COUNT_LAUNCH();
    cg_finish_kernel<<<1, CG_THREADS, 0, stream>>>((const double *)workspace, nb, scal, S_RR, S_SOLD);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_cg_hvp_combine(const double *z, const double *d, const double *p, const int32_t *indptr,
                                 const int32_t *indices, const double *values, const double *rdiag,
                                 double scale, double beta, const uint8_t *inactive, int64_t n, double *q,
                                 double *scal, void *workspace, int64_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n >= 0, "negative length");
    PF_CHECK_ARG(scal != nullptr, "null scalars");
    PF_CHECK_ARG(n == 0 || (z && p && q), "null vector");
    PF_CHECK_ARG(indptr == nullptr || (indices && values), "CSR arrays incomplete");
    int st = check_ws(workspace, workspace_bytes, n);
    if (st != PF_OK) return st;
    const int nb = cg_blocks(n);
    cg_hvp_kernel<<<nb, CG_THREADS, 0, stream>>>(z, d, p, indptr, indices, values, rdiag, scale, beta,
                                                 inactive, n, q, (double *)workspace);
    PF_COUNT_LAUNCH();
    cg_finish_kernel<<<1, CG_THREADS, 0, stream>>>((const double *)workspace, nb, scal, S_PQ, -1);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_cg_update(const double *p, const double *q, const double *minv, int64_t n, double *step,
                            double *r, double *h, double *scal, void *workspace, int64_t workspace_bytes,
                            void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n >= 0, "negative length");
    PF_CHECK_ARG(scal != nullptr, "null scalars");
    PF_CHECK_ARG(n == 0 || (p && q && step && r && h), "null vector");
    int st = check_ws(workspace, workspace_bytes, n);
    if (st != PF_OK) return st;
    const int nb = cg_blocks(n);
    cg_update_kernel<<<nb, CG_THREADS, 0, stream>>>(p, q, minv, n, scal, step, r, h, (double *)workspace);
    PF_COUNT_LAUNCH();
    cg_finish_kernel<<<1, CG_THREADS, 0, stream>>>((const double *)workspace, nb, scal, S_RR, S_SNEW);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_cg_direction(const double *h, int64_t n, double *p, double *scal, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n >= 0, "negative length");
    PF_CHECK_ARG(scal != nullptr, "null scalars");
    PF_CHECK_ARG(n == 0 || (h && p), "null vector");
    cg_direction_kernel<<<cg_blocks(n), CG_THREADS, 0, stream>>>(h, n, scal, p);
    PF_COUNT_LAUNCH();
    cg_commit_kernel<<<1, 1, 0, stream>>>(scal);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_vec_mul(const double *a, const double *b, int64_t n, double *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n >= 0, "negative length");
    if (n == 0) return PF_OK;
    PF_CHECK_ARG(a && b && out, "null vector");
    vec_mul_kernel<<<cg_blocks(n), CG_THREADS, 0, stream>>>(a, b, n, out);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

extern "C" int pf_csr_spmv(const int32_t *indptr, const int32_t *indices, const double *values, int64_t n_rows,
                           const double *x, double *y, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PF_CHECK_ARG(n_rows >= 0, "negative row count");
    if (n_rows == 0) return PF_OK;
    PF_CHECK_ARG(indptr && y, "null pointer");
    csr_spmv_kernel<<<cg_blocks(n_rows), CG_THREADS, 0, stream>>>(indptr, indices, values, n_rows, x, y);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}
