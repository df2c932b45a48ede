// pf_build.cu -- the FP64-bound kernels of the hot path: per-(receiver, prism)
// closed-form kernel evaluation with four sinks (store G / forward / G^T v /
// diag(G^T W G)).  Compiled with -fmad=false so that the kernel arguments, the
// radius and the 8-vertex sums round exactly like the reference's numba code
// (which never contracts to FMA).
//
// Execution model (one CTA = one tile of <=512 cells x one chunk of receivers):
//   - the tile's node coordinates, the cells' 8 local node ids and their G
//     columns are loaded ONCE into registers;
//   - for every receiver of the chunk:
//       phase 1  each thread evaluates the closed-form kernels on <=3 tile
//                nodes (atan/log/sqrt in fp64) and writes the needed node
//                quantities as planes in shared memory  -- this is the
//                reference's per-receiver `kernels` array
//                (gravity/_numba/_3d_mesh.py:147-157), at tile granularity;
//       phase 2  each thread combines the 8 vertices of its <=2 cells with the
//                reference's sign pattern and order (_numba_utils.py:18-45),
//                applies the per-component algebra and feeds the sink.
//   - nothing is communicated between CTAs; column/row partial sums go to a
//     workspace and are reduced in a fixed order (deterministic, no atomics).
#include "pf_common.cuh"

namespace {

// ---------------------------------------------------------------------------
// node quantities
// ---------------------------------------------------------------------------
// family 12: first + second order kernels (3 atan + 3 log primitives)
enum { Q_EE = 0, Q_NN, Q_UU, Q_EN, Q_EU, Q_NU, Q_E, Q_N, Q_U, Q_UV };
// family 3: third order kernels (algebraic)
enum { T_EEE = 0, T_NNN, T_UUU, T_EEN, T_EEU, T_ENN, T_NNU, T_EUU, T_NUU, T_ENU };
// primitive mask bits (family 12)
enum { NEED_AEE = 1, NEED_ANN = 2, NEED_AUU = 4, NEED_LEN = 8, NEED_LEU = 16, NEED_LNU = 32 };

enum { CT_SINGLE = 0, CT_MAG3 = 1, CT_TMI6 = 2 };
enum { SINK_STORE = 0, SINK_FORWARD = 1, SINK_GTV = 2, SINK_DIAG = 3 };

constexpr int MAX_SLOTS = 10;
constexpr int MAX_DESC = PF_MAX_COMPONENTS;
constexpr int WARPS = PF_THREADS / 32;

struct CompDesc {
    int type;     // CT_*
    int ci;       // position in the receiver object's component list
    int slot[6];  // shared-memory planes holding the node quantities it needs
    double factor;  // gravity: G_newton * unit factor; magnetics: (1/4pi) * |B0|
};

struct LaunchParams {
    pf_plan plan;
    const double *receivers;
    int64_t n_receivers;
    int rc;            // receivers per chunk (blockIdx.y)
    int n_comp_total;  // components of the receiver object (row stride)
    int64_t row_offset;
    int n_desc;
    int n_slots;
    int need;          // family-12 primitive mask
    int nblk;          // 1 (gravity, scalar) or 3 (vector model)
    int slot_q[MAX_SLOTS];
    CompDesc desc[MAX_DESC];
    double fx, fy, fz;  // unit inducing field
    // sinks
    void *G;
    int64_t ldG;
    const double *model;  // forward
    const double *vec;    // v (G^T v) or w (diag), full nD vectors
    double *partial;      // workspace
};

// ---------------------------------------------------------------------------
// choclo restatement (device).  Same formulas and operation order as
// oracle/pf_oracle.c; see DESIGN.md for the provenance of each formula.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double safe_atan2(double y, double x)
{
    // choclo._utils._safe_atan2: arctan(y/x); +-pi/2 (0) when x == 0
    const double half_pi = 1.5707963267948966;
    double res = atan(y / x);
    if (x == 0.0) res = (y > 0.0) ? half_pi : ((y < 0.0) ? -half_pi : 0.0);
    return res;
}

__device__ __forceinline__ double safe_log(double x, double y, double z, double r)
{
    // choclo._utils._safe_log: log(x + r), evaluated as
    // log((y^2+z^2)/(r-x)) for x<0 and -log(-2x) on the axis (r == -x).
    // One log call on a selected argument keeps the warp convergent.
    const bool neg = x < 0.0;
    const bool axis = neg && (r == -x);
    double arg = x + r;
    if (neg) arg = (y * y + z * z) / (r - x);
    if (axis) arg = -2 * x;
    double res = log(arg);
    if (axis) res = -res;
    if (r == 0.0) res = 0.0;
    return res;
}

__device__ __forceinline__ double kernel_iii(double xi, double xj, double xk, double r)
{
    if ((xi == 0.0 && xj == 0.0) || (xi == 0.0 && xk == 0.0)) return 0.0;
    double xi2 = xi * xi, xj2 = xj * xj, xk2 = xk * xk;
    double numerator = -xj * xk * (2 * xi2 + xj2 + xk2);
    double denominator = (xi2 + xj2) * (xi2 + xk2) * r;
    return numerator / denominator;
}

__device__ __forceinline__ double kernel_iij(double xi, double xj, double xk, double r)
{
    if (xi == 0.0 && xj == 0.0 && xk < 0.0) return 0.0;
    return -xi / ((xk + r) * r);
}

// Evaluate the requested node quantities of one (receiver, node) pair and
// store them in their planes.
template <int FAM>
__device__ __forceinline__ void eval_node(const LaunchParams &p, double e, double n, double u,
                                          double *planes, int j)
{
    // gravity/_numba/_3d_mesh.py:494-498: distance = sqrt(dx**2 + dy**2 + dz**2)
    const double r = sqrt(e * e + n * n + u * u);
    if (FAM == 12) {
        const int need = p.need;
        double Aee = 0, Ann = 0, Auu = 0, Len = 0, Leu = 0, Lnu = 0;
        if (need & NEED_AEE) Aee = safe_atan2(n * u, e * r);
        if (need & NEED_ANN) Ann = safe_atan2(e * u, n * r);
        if (need & NEED_AUU) Auu = safe_atan2(e * n, u * r);
        if (need & NEED_LEN) Len = safe_log(u, e, n, r);
        if (need & NEED_LEU) Leu = safe_log(n, e, u, r);
        if (need & NEED_LNU) Lnu = safe_log(e, n, u, r);
        for (int s = 0; s < p.n_slots; ++s) {
            double val;
            switch (p.slot_q[s]) {
            case Q_EE: val = -Aee; break;
            case Q_NN: val = -Ann; break;
            case Q_UU: val = -Auu; break;
            case Q_EN: val = Len; break;
            case Q_EU: val = Leu; break;
            case Q_NU: val = Lnu; break;
            // choclo.prism.kernel_e / kernel_n / kernel_u
            case Q_E: val = -(n * Len + u * Leu - e * Aee); break;
            case Q_N: val = -(u * Lnu + e * Len - n * Ann); break;
            case Q_U: val = -(e * Leu + n * Lnu - u * Auu); break;
            // gravity/simulation.py:23-30 kernel_uv = 0.5 (kernel_nn - kernel_ee)
            default: val = 0.5 * ((-Ann) - (-Aee)); break;
            }
            planes[s * PF_TILE_NODES + j] = val;
        }
    } else {
        for (int s = 0; s < p.n_slots; ++s) {
            double val;
            switch (p.slot_q[s]) {
            case T_EEE: val = kernel_iii(e, n, u, r); break;
            case T_NNN: val = kernel_iii(n, u, e, r); break;
            case T_UUU: val = kernel_iii(u, e, n, r); break;
            case T_EEN: val = kernel_iij(e, n, u, r); break;
            case T_EEU: val = kernel_iij(e, u, n, r); break;
            case T_ENN: val = kernel_iij(n, e, u, r); break;
            case T_NNU: val = kernel_iij(n, u, e, r); break;
            case T_EUU: val = kernel_iij(u, e, n, r); break;
            case T_NUU: val = kernel_iij(u, n, e, r); break;
            default: val = (r == 0.0) ? 0.0 : -1.0 / r; break;  // kernel_enu
            }
            planes[s * PF_TILE_NODES + j] = val;
        }
    }
}

// _numba_utils.py:18-45 kernels_in_nodes_to_cell: -k0+k1+k2-k3+k4-k5-k6+k7,
// evaluated left to right.
__device__ __forceinline__ double vertex_sum(const double *pl, const uint4 ci)
{
    const int i0 = ci.x & 0xffff, i1 = ci.x >> 16, i2 = ci.y & 0xffff, i3 = ci.y >> 16;
    const int i4 = ci.z & 0xffff, i5 = ci.z >> 16, i6 = ci.w & 0xffff, i7 = ci.w >> 16;
    return -pl[i0] + pl[i1] + pl[i2] - pl[i3] + pl[i4] - pl[i5] - pl[i6] + pl[i7];
}

// Per-cell, per-component algebra.  Returns g[0..nblk).
__device__ __forceinline__ void cell_component(const LaunchParams &p, const CompDesc &d,
                                               const double *planes, const uint4 ci,
                                               double &g0, double &g1, double &g2)
{
    g1 = 0.0;
    g2 = 0.0;
    if (d.type == CT_SINGLE) {
        // gravity/_numba/_3d_mesh.py:159-163
        g0 = d.factor * vertex_sum(planes + d.slot[0] * PF_TILE_NODES, ci);
        return;
    }
    double bx, by, bz;
    if (d.type == CT_MAG3) {
        // magnetics/_numba/_3d_mesh.py:182-187
        bx = vertex_sum(planes + d.slot[0] * PF_TILE_NODES, ci);
        by = vertex_sum(planes + d.slot[1] * PF_TILE_NODES, ci);
        bz = vertex_sum(planes + d.slot[2] * PF_TILE_NODES, ci);
    } else {
        // magnetics/_numba/_3d_mesh.py:344-354
        const double uxx = vertex_sum(planes + d.slot[0] * PF_TILE_NODES, ci);
        const double uyy = vertex_sum(planes + d.slot[1] * PF_TILE_NODES, ci);
        const double uzz = vertex_sum(planes + d.slot[2] * PF_TILE_NODES, ci);
        const double uxy = vertex_sum(planes + d.slot[3] * PF_TILE_NODES, ci);
        const double uxz = vertex_sum(planes + d.slot[4] * PF_TILE_NODES, ci);
        const double uyz = vertex_sum(planes + d.slot[5] * PF_TILE_NODES, ci);
        bx = uxx * p.fx + uxy * p.fy + uxz * p.fz;
        by = uxy * p.fx + uyy * p.fy + uyz * p.fz;
        bz = uxz * p.fx + uyz * p.fy + uzz * p.fz;
    }
    if (p.nblk == 1) {
        g0 = d.factor * (bx * p.fx + by * p.fy + bz * p.fz);
    } else {
        g0 = d.factor * bx;
        g1 = d.factor * by;
        g2 = d.factor * bz;
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int FAM, int SINK, typename GT>
__global__ void __launch_bounds__(PF_THREADS, 2) pf_tile_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ double smem[];
    double *planes = smem;                                      // [n_slots][TN]
    double *wpart = smem + (size_t)p.n_slots * PF_TILE_NODES;   // [WARPS][n_desc] (forward)

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t r_begin = (int64_t)blockIdx.y * p.rc;
    const int64_t r_end = min(r_begin + (int64_t)p.rc, p.n_receivers);
    const int64_t nC = p.plan.n_cells;

    // ---- tile-stationary state in registers --------------------------------
    const int nn = p.plan.tile_n_nodes[tile];
    double nx[PF_NPT], ny[PF_NPT], nz[PF_NPT];
    const double *xyz = p.plan.tile_xyz + tile * (3 * PF_TILE_NODES);
#pragma unroll
    for (int k = 0; k < PF_NPT; ++k) {
        const int j = tid + k * PF_THREADS;
        nx[k] = xyz[j];
        ny[k] = xyz[PF_TILE_NODES + j];
        nz[k] = xyz[2 * PF_TILE_NODES + j];
    }
    int col[PF_CPT];
    uint4 cidx[PF_CPT];
    double mdl[PF_CPT][3];
    double acc[PF_CPT][3];
#pragma unroll
    for (int s = 0; s < PF_CPT; ++s) {
        const int64_t slot = tile * PF_TILE_CELLS + tid + s * PF_THREADS;
        col[s] = p.plan.tile_cell_col[slot];
        cidx[s] = reinterpret_cast<const uint4 *>(p.plan.tile_cell_nodes)[slot];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            acc[s][b] = 0.0;
            mdl[s][b] = 0.0;
            if (SINK == SINK_FORWARD && col[s] >= 0 && b < p.nblk)
                mdl[s][b] = p.model[col[s] + b * nC];
        }
    }

    for (int64_t r = r_begin; r < r_end; ++r) {
        const double rx = __ldg(p.receivers + 3 * r);
        const double ry = __ldg(p.receivers + 3 * r + 1);
        const double rz = __ldg(p.receivers + 3 * r + 2);

        // ---- phase 1: node quantities -> shared memory planes ---------------
#pragma unroll
        for (int k = 0; k < PF_NPT; ++k) {
            const int j = tid + k * PF_THREADS;
            if (j < nn) eval_node<FAM>(p, nx[k] - rx, ny[k] - ry, nz[k] - rz, planes, j);
        }
        __syncthreads();

        // ---- phase 2: cells ---------------------------------------------------
        for (int d = 0; d < p.n_desc; ++d) {
            const CompDesc &cd = p.desc[d];
            const int64_t row = p.row_offset + r * p.n_comp_total + cd.ci;
            double rowval = 0.0;  // v[row] or w[row]
            if (SINK == SINK_GTV || SINK == SINK_DIAG) rowval = __ldg(p.vec + row);
            double fsum = 0.0;
#pragma unroll
            for (int s = 0; s < PF_CPT; ++s) {
                if (col[s] < 0) continue;
                double g0, g1, g2;
                cell_component(p, cd, planes, cidx[s], g0, g1, g2);
                if (SINK == SINK_STORE) {
                    GT *grow = reinterpret_cast<GT *>(p.G) + row * p.ldG + col[s];
                    grow[0] = (GT)g0;
                    if (p.nblk == 3) {
                        grow[nC] = (GT)g1;
                        grow[2 * nC] = (GT)g2;
                    }
                } else if (SINK == SINK_FORWARD) {
                    fsum += g0 * mdl[s][0] + g1 * mdl[s][1] + g2 * mdl[s][2];
                } else if (SINK == SINK_GTV) {
                    acc[s][0] += rowval * g0;
                    acc[s][1] += rowval * g1;
                    acc[s][2] += rowval * g2;
                } else {
                    acc[s][0] += rowval * (g0 * g0);
                    acc[s][1] += rowval * (g1 * g1);
                    acc[s][2] += rowval * (g2 * g2);
                }
            }
            if (SINK == SINK_FORWARD) {
                fsum = warp_sum(fsum);
                if ((tid & 31) == 0) wpart[(tid >> 5) * p.n_desc + d] = fsum;
            }
        }
        __syncthreads();
        if (SINK == SINK_FORWARD && tid < p.n_desc) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += wpart[w * p.n_desc + tid];
            // partial[tile][local row]; local row = r * n_desc_total-of-this-launch...
            const int64_t n_local = p.n_receivers * p.n_desc;
            p.partial[tile * n_local + r * p.n_desc + tid] = t;
        }
    }

    if (SINK == SINK_GTV || SINK == SINK_DIAG) {
        const int64_t nP = nC * p.nblk;
        double *out = p.partial + (int64_t)blockIdx.y * nP;
#pragma unroll
        for (int s = 0; s < PF_CPT; ++s) {
            if (col[s] < 0) continue;
            out[col[s]] = acc[s][0];
            if (p.nblk == 3) {
                out[col[s] + nC] = acc[s][1];
                out[col[s] + 2 * nC] = acc[s][2];
            }
        }
    }
}

// fields[row_offset + r*n_comp_total + ci[d]] += sum_tiles partial[tile][r*n_desc + d]
struct ReduceFwdParams {
    const double *partial;
    int64_t n_tiles;
    int64_t n_receivers;
    int n_desc;
    int n_comp_total;
    int64_t row_offset;
    int ci[MAX_DESC];
    double *fields;
};

__global__ void __launch_bounds__(256) pf_reduce_forward_kernel(const __grid_constant__ ReduceFwdParams q)
{
    const int64_t n_local = q.n_receivers * q.n_desc;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    double t = 0.0;
    for (int64_t tile = 0; tile < q.n_tiles; ++tile) t += q.partial[tile * n_local + i];
    const int64_t r = i / q.n_desc;
    const int d = (int)(i - r * q.n_desc);
    q.fields[q.row_offset + r * q.n_comp_total + q.ci[d]] += t;
}

// out[k] += sum_chunks partial[chunk][k]
__global__ void __launch_bounds__(256) pf_reduce_columns_kernel(const double *partial, int64_t n_chunks,
                                                                int64_t nP, double *out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nP) return;
    double t = 0.0;
    for (int64_t c = 0; c < n_chunks; ++c) t += partial[c * nP + k];
    out[k] += t;
}

// ---------------------------------------------------------------------------
// host side: job -> launches
// ---------------------------------------------------------------------------
struct CompInfo {
    int family;  // 12 or 3
    int type;
    int q[6];
    int nq;
};

bool component_info(int physics, int comp, CompInfo &ci)
{
    auto set = [&](int fam, int type, std::initializer_list<int> qs) {
        ci.family = fam;
        ci.type = type;
        ci.nq = 0;
        for (int q : qs) ci.q[ci.nq++] = q;
        return true;
    };
    if (physics == PF_GRAVITY) {
        switch (comp) {
        case PF_GX: return set(12, CT_SINGLE, {Q_E});
        case PF_GY: return set(12, CT_SINGLE, {Q_N});
        case PF_GZ: return set(12, CT_SINGLE, {Q_U});
        case PF_GXX: return set(12, CT_SINGLE, {Q_EE});
        case PF_GYY: return set(12, CT_SINGLE, {Q_NN});
        case PF_GZZ: return set(12, CT_SINGLE, {Q_UU});
        case PF_GXY: return set(12, CT_SINGLE, {Q_EN});
        case PF_GXZ: return set(12, CT_SINGLE, {Q_EU});
        case PF_GYZ: return set(12, CT_SINGLE, {Q_NU});
        case PF_GUV: return set(12, CT_SINGLE, {Q_UV});
        default: return false;
        }
    }
    switch (comp) {
    case PF_TMI: return set(12, CT_TMI6, {Q_EE, Q_NN, Q_UU, Q_EN, Q_EU, Q_NU});
    case PF_BX: return set(12, CT_MAG3, {Q_EE, Q_EN, Q_EU});
    case PF_BY: return set(12, CT_MAG3, {Q_EN, Q_NN, Q_NU});
    case PF_BZ: return set(12, CT_MAG3, {Q_EU, Q_NU, Q_UU});
    case PF_BXX: return set(3, CT_MAG3, {T_EEE, T_EEN, T_EEU});
    case PF_BYY: return set(3, CT_MAG3, {T_ENN, T_NNN, T_NNU});
    case PF_BZZ: return set(3, CT_MAG3, {T_EUU, T_NUU, T_UUU});
    case PF_BXY: return set(3, CT_MAG3, {T_EEN, T_ENN, T_ENU});
    case PF_BXZ: return set(3, CT_MAG3, {T_EEU, T_ENU, T_EUU});
    case PF_BYZ: return set(3, CT_MAG3, {T_ENU, T_NNU, T_NUU});
    case PF_TMI_X: return set(3, CT_TMI6, {T_EEE, T_ENN, T_EUU, T_EEN, T_EEU, T_ENU});
    case PF_TMI_Y: return set(3, CT_TMI6, {T_EEN, T_NNN, T_NUU, T_ENN, T_ENU, T_NNU});
    case PF_TMI_Z: return set(3, CT_TMI6, {T_EEU, T_NNU, T_UUU, T_ENU, T_EUU, T_NUU});
    default: return false;
    }
}

int need_of_quantity(int q)
{
    switch (q) {
    case Q_EE: return NEED_AEE;
    case Q_NN: return NEED_ANN;
    case Q_UU: return NEED_AUU;
    case Q_EN: return NEED_LEN;
    case Q_EU: return NEED_LEU;
    case Q_NU: return NEED_LNU;
    case Q_E: return NEED_LEN | NEED_LEU | NEED_AEE;
    case Q_N: return NEED_LNU | NEED_LEN | NEED_ANN;
    case Q_U: return NEED_LEU | NEED_LNU | NEED_AUU;
    default: return NEED_ANN | NEED_AEE;  // Q_UV
    }
}

int validate(const pf_plan *plan, const pf_job *job)
{
    PF_CHECK_ARG(plan && job, "null plan or job");
    PF_CHECK_ARG(plan->tile_cells == PF_TILE_CELLS && plan->tile_nodes == PF_TILE_NODES,
                 "tile geometry does not match the library (512 cells / 768 nodes)");
    PF_CHECK_ARG(plan->n_tiles >= 0 && plan->n_cells >= 0, "negative sizes");
    PF_CHECK_ARG(plan->n_tiles == 0 ||
                     (plan->tile_xyz && plan->tile_cell_nodes && plan->tile_cell_col && plan->tile_n_nodes),
                 "null plan array");
    PF_CHECK_ARG(job->physics == PF_GRAVITY || job->physics == PF_MAGNETICS, "bad physics");
    PF_CHECK_ARG(job->n_components >= 1 && job->n_components <= PF_MAX_COMPONENTS,
                 "n_components out of range");
    PF_CHECK_ARG(job->n_receivers >= 0, "negative n_receivers");
    PF_CHECK_ARG(job->n_receivers == 0 || job->receivers, "null receivers");
    for (int c = 0; c < job->n_components; ++c) {
        CompInfo ci;
        if (!component_info(job->physics, job->components[c], ci)) {
            pf_set_error("component code %d is not supported for physics %d",
                         job->components[c], job->physics);
            return PF_ERR_UNSUPPORTED;
        }
    }
    if (job->physics == PF_MAGNETICS) {
        const double a = job->b0[0] * job->b0[0] + job->b0[1] * job->b0[1] + job->b0[2] * job->b0[2];
        PF_CHECK_ARG(a > 0.0, "magnetics needs a non-zero inducing field b0");
    }
    return PF_OK;
}

int n_blocks_of(const pf_job *job)
{
    return (job->physics == PF_MAGNETICS && !job->scalar_model) ? 3 : 1;
}

// Fill the launch parameters of one family; returns false if the job has no
// component of that family.
bool fill_family(const pf_plan *plan, const pf_job *job, int family, LaunchParams &p)
{
    std::memset(&p, 0, sizeof(p));
    p.plan = *plan;
    p.receivers = job->receivers;
    p.n_receivers = job->n_receivers;
    p.n_comp_total = job->n_components;
    p.row_offset = job->row_offset;
    p.nblk = n_blocks_of(job);
    int slot_of_q[16];
    for (int i = 0; i < 16; ++i) slot_of_q[i] = -1;
    double amp = 1.0;
    if (job->physics == PF_MAGNETICS) {
        // magnetics/_numba/_3d_mesh.py:165-169
        const double fx = job->b0[0], fy = job->b0[1], fz = job->b0[2];
        amp = sqrt(fx * fx + fy * fy + fz * fz);
        p.fx = fx / amp;
        p.fy = fy / amp;
        p.fz = fz / amp;
    }
    for (int c = 0; c < job->n_components; ++c) {
        CompInfo ci;
        component_info(job->physics, job->components[c], ci);
        if (ci.family != family) continue;
        CompDesc &d = p.desc[p.n_desc++];
        d.type = ci.type;
        d.ci = c;
        for (int k = 0; k < ci.nq; ++k) {
            const int q = ci.q[k];
            if (slot_of_q[q] < 0) {
                slot_of_q[q] = p.n_slots;
                p.slot_q[p.n_slots++] = q;
                if (family == 12) p.need |= need_of_quantity(q);
            }
            d.slot[k] = slot_of_q[q];
        }
        if (job->physics == PF_GRAVITY) {
            // gravity/simulation.py:101-111 and 580: constants.G * conversion_factor
            const double conv = (job->components[c] <= PF_GZ) ? 1e8 : 1e12;
            d.factor = PF_NEWTON_G * conv;
        } else {
            // magnetics/simulation.py:958 constant_factor = 1/4/pi, times |B0|
            const double cf = 1.0 / 4.0 / 3.141592653589793;
            d.factor = cf * amp;
        }
    }
    return p.n_desc > 0;
}

// receivers per chunk: enough CTAs to fill the machine a few times over
int choose_rc(int64_t n_tiles, int64_t n_receivers, int64_t max_chunks)
{
    const int64_t target_ctas = 148 * 2 * 8;
    int64_t chunks = pf_ceil_div(target_ctas, n_tiles > 0 ? n_tiles : 1);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t rc = pf_ceil_div(n_receivers, chunks);
    if (rc < 1) rc = 1;
    if (rc > 65535) rc = 65535;
    return (int)rc;
}

template <int FAM, int SINK, typename GT>
int launch_tile_kernel(const LaunchParams &p, cudaStream_t stream)
{
    const size_t smem = ((size_t)p.n_slots * PF_TILE_NODES + (SINK == SINK_FORWARD ? WARPS * p.n_desc : 0))
                        * sizeof(double);
    auto kernel = pf_tile_kernel<FAM, SINK, GT>;
    PF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t chunks = pf_ceil_div(p.n_receivers, p.rc);
    PF_CHECK_ARG(chunks <= 65535, "too many receiver chunks");
    dim3 grid((unsigned)p.plan.n_tiles, (unsigned)chunks);
    kernel<<<grid, PF_THREADS, smem, stream>>>(p);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

template <int SINK, typename GT>
int launch_family(int family, const LaunchParams &p, cudaStream_t stream)
{
    if (family == 12) return launch_tile_kernel<12, SINK, GT>(p, stream);
    return launch_tile_kernel<3, SINK, GT>(p, stream);
}

const int FAMILIES[2] = {12, 3};

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int64_t pf_job_columns(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    return plan->n_cells * n_blocks_of(job);
}

extern "C" int pf_build_G(const pf_plan *plan, const pf_job *job, void *G, int64_t ldG,
                          int g_dtype, void *stream)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    PF_CHECK_ARG(g_dtype == PF_F64 || g_dtype == PF_F32, "bad g_dtype");
    if (plan->n_tiles == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(G != nullptr, "null G");
    PF_CHECK_ARG(ldG >= pf_job_columns(plan, job), "ldG smaller than the number of columns");
    for (int family : FAMILIES) {
        LaunchParams p;
        if (!fill_family(plan, job, family, p)) continue;
        p.G = G;
        p.ldG = ldG;
        p.rc = choose_rc(plan->n_tiles, job->n_receivers, 65535);
        st = (g_dtype == PF_F64) ? launch_family<SINK_STORE, double>(family, p, (cudaStream_t)stream)
                                 : launch_family<SINK_STORE, float>(family, p, (cudaStream_t)stream);
        if (st != PF_OK) return st;
    }
    return PF_OK;
}

extern "C" int64_t pf_forward_workspace_bytes(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    return plan->n_tiles * job->n_receivers * (int64_t)job->n_components * (int64_t)sizeof(double);
}

extern "C" int pf_forward(const pf_plan *plan, const pf_job *job, const double *model,
                          double *fields, void *workspace, int64_t workspace_bytes, void *stream)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    if (plan->n_tiles == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(model && fields, "null model or fields");
    if (workspace_bytes < pf_forward_workspace_bytes(plan, job) || !workspace) {
        pf_set_error("pf_forward: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    for (int family : FAMILIES) {
        LaunchParams p;
        if (!fill_family(plan, job, family, p)) continue;
        p.model = model;
        p.partial = (double *)workspace;
        p.rc = choose_rc(plan->n_tiles, job->n_receivers, 65535);
        st = launch_family<SINK_FORWARD, double>(family, p, (cudaStream_t)stream);
        if (st != PF_OK) return st;
        ReduceFwdParams q;
        q.partial = p.partial;
        q.n_tiles = plan->n_tiles;
        q.n_receivers = job->n_receivers;
        q.n_desc = p.n_desc;
        q.n_comp_total = job->n_components;
        q.row_offset = job->row_offset;
        for (int d = 0; d < p.n_desc; ++d) q.ci[d] = p.desc[d].ci;
        q.fields = fields;
        const int64_t n_local = job->n_receivers * p.n_desc;
        pf_reduce_forward_kernel<<<(unsigned)pf_ceil_div(n_local, 256), 256, 0, (cudaStream_t)stream>>>(q);
    PF_COUNT_LAUNCH();
        PF_CUDA_OK(cudaGetLastError());
    }
    return PF_OK;
}

namespace {

constexpr int64_t MAX_COLUMN_CHUNKS = 64;

int64_t column_chunks(const pf_plan *plan, const pf_job *job)
{
    const int rc = choose_rc(plan->n_tiles, job->n_receivers, MAX_COLUMN_CHUNKS);
    return pf_ceil_div(job->n_receivers > 0 ? job->n_receivers : 1, rc);
}

template <int SINK>
int column_sink(const pf_plan *plan, const pf_job *job, const double *vec, double *out,
                void *workspace, int64_t workspace_bytes, void *stream)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    if (plan->n_tiles == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(vec && out, "null vector or output");
    const int64_t nP = pf_job_columns(plan, job);
    const int64_t chunks = column_chunks(plan, job);
    if (!workspace || workspace_bytes < chunks * nP * (int64_t)sizeof(double)) {
        pf_set_error("column sink: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    for (int family : FAMILIES) {
        LaunchParams p;
        if (!fill_family(plan, job, family, p)) continue;
        p.vec = vec;
        p.partial = (double *)workspace;
        p.rc = choose_rc(plan->n_tiles, job->n_receivers, MAX_COLUMN_CHUNKS);
        st = launch_family<SINK, double>(family, p, (cudaStream_t)stream);
        if (st != PF_OK) return st;
        pf_reduce_columns_kernel<<<(unsigned)pf_ceil_div(nP, 256), 256, 0, (cudaStream_t)stream>>>(
            p.partial, chunks, nP, out);
    PF_COUNT_LAUNCH();
        PF_CUDA_OK(cudaGetLastError());
    }
    return PF_OK;
}

}  // namespace

extern "C" int64_t pf_gt_dot_v_workspace_bytes(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    return column_chunks(plan, job) * pf_job_columns(plan, job) * (int64_t)sizeof(double);
}

extern "C" int pf_gt_dot_v(const pf_plan *plan, const pf_job *job, const double *v, double *result,
                           void *workspace, int64_t workspace_bytes, void *stream)
{
    return column_sink<SINK_GTV>(plan, job, v, result, workspace, workspace_bytes, stream);
}

extern "C" int64_t pf_gtg_diag_workspace_bytes(const pf_plan *plan, const pf_job *job)
{
    return pf_gt_dot_v_workspace_bytes(plan, job);
}

extern "C" int pf_gtg_diag(const pf_plan *plan, const pf_job *job, const double *w, double *diag,
                           void *workspace, int64_t workspace_bytes, void *stream)
{
    return column_sink<SINK_DIAG>(plan, job, w, diag, workspace, workspace_bytes, stream);
}
