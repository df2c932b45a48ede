// pf_build.cu -- the FP64-bound kernels of the hot path: per-(receiver, prism)
// closed-form kernel evaluation with four sinks (store G / forward / G^T v /
// diag(G^T W G)).  Compiled with -fmad=false so that the kernel arguments, the
// radius and the 8-vertex sums round like the reference's numba code (which
// never contracts to FMA); the explicit fma() calls of
// pf_math.cuh are unaffected.
//
// Execution model (one CTA = one tile of <=512 cells x one chunk of receivers):
//   - the tile's node coordinates, the cells' 8 local node ids and their G
//     columns are loaded ONCE into registers;
//   - for every receiver of the chunk:
//       phase 1  each thread evaluates the closed-form kernels on <=3 tile
//                nodes (pf_math.cuh: 1 sqrt, 3 atan, 3 log in fp64) and writes
//                one value per OUTPUT PLANE to shared memory.  A plane is a
//                linear combination of node kernels (a raw kernel, k . f, or
//                f^T K f): because the 8-vertex sum is linear, the per-cell
//                algebra of the reference (magnetics/_numba/_3d_mesh.py:
//                344-372) is hoisted to the nodes, so that phase 2 gathers
//                8 values per output instead of 8 per kernel;
//       phase 2  each thread combines the 8 vertices of its <=2 cells with the
//                reference's sign pattern and order (_numba_utils.py:18-45),
//                scales and feeds the sink.
//   - nothing is communicated between CTAs; column/row partial sums go to a
//     workspace and are reduced in a fixed order (deterministic, no atomics).
#include "pf_common.cuh"
#include "pf_math.cuh"

namespace {

// ---------------------------------------------------------------------------
// node kernels ("raw" quantities) and output planes
// ---------------------------------------------------------------------------
// family 12: first + second order kernels (3 atan + 3 log primitives)
enum { Q_EE = 0, Q_NN, Q_UU, Q_EN, Q_EU, Q_NU, Q_E, Q_N, Q_U, Q_UV };
// family 3: third order kernels (algebraic)
enum { T_EEE = 0, T_NNN, T_UUU, T_EEN, T_EEU, T_ENN, T_NNU, T_EUU, T_NUU, T_ENU };
// primitive mask bits (family 12)
enum { NEED_AEE = 1, NEED_ANN = 2, NEED_AUU = 4, NEED_LEN = 8, NEED_LEU = 16, NEED_LNU = 32 };

enum { CT_SINGLE = 0, CT_MAG3 = 1, CT_TMI6 = 2 };
enum { PK_RAW = 0, PK_DOT3 = 1, PK_QUAD = 2 };
enum { SINK_STORE = 0, SINK_FORWARD = 1, SINK_GTV = 2, SINK_DIAG = 3 };
// PROG_TMI_*: one component "tmi" (the value is the number of column blocks)
// PROG_ALL6: family 12 with all six primitives needed (descriptor-driven planes, but the node
// evaluation is branch-free so the compiler interleaves the five transcendentals)
// PROG_G1: one first-order gravity component (gx, gy or gz) -- one atan and two logs advanced in
// lock-step (the generic program evaluates them one after the other behind run-time masks)
enum { PROG_GENERIC = 0, PROG_TMI_SCALAR = 1, PROG_G1 = 2, PROG_TMI_VECTOR = 3, PROG_ALL6 = 6 };
#define PF_IS_TMI(P) ((P) == PROG_TMI_SCALAR || (P) == PROG_TMI_VECTOR)
// programs whose outputs are known at compile time: output o = plane o = column block o of the one
// component, one common factor (PF_NOUT of them)
#define PF_FIXED(P) (PF_IS_TMI(P) || (P) == PROG_G1)
#define PF_NOUT(P) ((P) == PROG_TMI_VECTOR ? 3 : 1)

constexpr int MAX_PLANES = 10;
constexpr int MAX_OUT = 32;
constexpr int WARPS = PF_THREADS / 32;

// value stored per (receiver, node):  RAW: k[q0];  DOT3: k[q0] fx + k[q1] fy + k[q2] fz;
// QUAD: f^T K f with K = (xx, yy, zz, xy, xz, yz) = k[q0..q5]
struct PlaneDesc {
    int kind;
    int q[6];
};

// G[row(ci), col + blk * nC] = factor * vertex_sum(plane)
struct OutDesc {
    int plane;
    int ci;    // position in the receiver object's component list (row)
    int blk;   // column block (0 for scalar models, 0..2 for vector models)
    int fi;    // index of the component inside this launch (forward partial sums)
    int last;  // last output of its component
    int role;  // ROLE_*: plain output, kept as operand A / B, or derived from A and B (no plane)
    double factor;  // gravity: G_newton * unit factor; magnetics: (1/4pi) * |B0|
};
// Derived outputs (marching kernel only).  gravity gzz needs no plane of its own: per node
// kernel_ee + kernel_nn + kernel_uu = -(Aee + Ann + Auu) = -S with S = +-pi/2 or 0 (the atan
// identity of eval_raw), so per cell  gzz = -(gxx + gyy) - factor * sum_vertices s_v S_v,  and
// the last term vanishes unless the receiver lies inside or on the closed prism (cell_S below).
enum { ROLE_PLAIN = 0, ROLE_KEEP_A = 1, ROLE_KEEP_B = 2, ROLE_DERIVED = 3 };

struct LaunchParams {
    pf_plan plan;
    const double *receivers;
    int64_t n_receivers;
    int rc;            // receivers per chunk (blockIdx.y)
    int n_comp_total;  // components of the receiver object (row stride)
    int64_t row_offset;
    int n_planes;
    int n_out;
    int n_fwd;         // components handled by this launch
    int need;          // family-12 primitive mask / family-3 kernel mask
    int nblk;          // 1 (gravity, scalar) or 3 (vector model)
    int direct;        // outs[o].plane == o for every o (phase 2 can unroll with constant offsets)
    int raw_mask;      // raw kernels stored as planes: bit q set -> slot = rank of q among the set bits
    int n_combo;       // planes [n_planes - n_combo, n_planes) are DOT3 / QUAD combinations
    int n_buf;         // marching kernel: 2 = node planes double-buffered (one barrier per layer), 1 = single
    int n_derived;     // outputs [n_out - n_derived, n_out) have no plane (ROLE_DERIVED)
    int g1_axis;       // PROG_G1: 0 = gx (kernel_e), 1 = gy (kernel_n), 2 = gz (kernel_u)
    // combination planes (DOT3 / QUAD) as linear combinations of the ten raw node kernels:
    // value = sum_q combo_coef[s][q] k[q] over the bits of combo_mask[s] (filled by make_passes)
    int combo_mask[MAX_PLANES];
    double combo_coef[MAX_PLANES][10];
    PlaneDesc planes[MAX_PLANES];
    OutDesc outs[MAX_OUT];
    double fx, fy, fz;  // unit inducing field
    double gx, gy, gz;  // TMI programs: factor x unit inducing field (the outputs' factors are 1)
    double f2x, f2y, f2z, g2x, g2y, g2z;  // 2 f and 2 g: coefficients of the halved logarithms
    // sinks
    void *G;
    int64_t ldG;
    const double *model;  // forward
    const double *vec;    // v (G^T v) or w (diag), full nD vectors
    double *partial;      // workspace
};

struct Raw10 {
    double v0, v1, v2, v3, v4, v5, v6, v7, v8, v9;
};

// choclo.prism third-order kernels (restated, see oracle/pf_oracle.c)
__device__ __forceinline__ double kernel_iii(double xi, double xj, double xk, double r)
{
    double xi2 = xi * xi, xj2 = xj * xj, xk2 = xk * xk;
    double numerator = -xj * xk * (2 * xi2 + xj2 + xk2);
    double denominator = (xi2 + xj2) * (xi2 + xk2) * r;
    double res = pf_div(numerator, denominator);
    if ((xi == 0.0 && xj == 0.0) || (xi == 0.0 && xk == 0.0)) res = 0.0;
    return res;
}

__device__ __forceinline__ double kernel_iij(double xi, double xj, double xk, double r)
{
    double res = pf_div(-xi, (xk + r) * r);
    if (xi == 0.0 && xj == 0.0 && xk < 0.0) res = 0.0;
    return res;
}

// Node kernels of one (receiver, node) pair.
// HALF_LOGS (TMI programs, all six primitives): k.v3 .. k.v5 hold log / 2 (pf_math.cuh)
template <int FAM, bool HALF_LOGS = false>
__device__ __forceinline__ void eval_raw(int need, double e, double n, double u, Raw10 &k)
{
    // gravity/_numba/_3d_mesh.py:494-498: distance = sqrt(dx**2 + dy**2 + dz**2)
    const double r2 = e * e + n * n + u * u;
    const bool r_is_zero = r2 == 0.0;
    const double r = pf_sqrt(r2, r_is_zero);
    if (FAM == 12) {
        double Aee = 0, Ann = 0, Auu = 0, Len = 0, Leu = 0, Lnu = 0;
        constexpr int NEED_ALL = NEED_AEE | NEED_ANN | NEED_AUU | NEED_LEN | NEED_LEU | NEED_LNU;
#if !defined(PF_NO_ATAN_IDENTITY) && !defined(PF_NO_LOCKSTEP)
        const bool lockstep = (need == NEED_ALL);
#else
        const bool lockstep = false;
#endif
        if (lockstep) {
            // all six primitives: two atans + three logs in lock-step (pf_math.cuh), third atan
            // from the identity below
            pf_node_atan2_log3<HALF_LOGS>(e, n, u, r, r_is_zero, Aee, Ann, Len, Leu, Lnu);
            need = NEED_AUU;
        }
        // one atan + the two logs of the same axis (bx / by / bz, gx / gy / gz): lock-step as well
        if (need == (NEED_AEE | NEED_LEN | NEED_LEU)) {
            pf_node_atan1_log2(e, n, u, r, r_is_zero, Aee, Len, Leu);
            need = 0;
        } else if (need == (NEED_ANN | NEED_LEN | NEED_LNU)) {
            pf_node_atan1_log2(n, u, e, r, r_is_zero, Ann, Lnu, Len);
            need = 0;
        } else if (need == (NEED_AUU | NEED_LEU | NEED_LNU)) {
            pf_node_atan1_log2(u, e, n, r, r_is_zero, Auu, Leu, Lnu);
            need = 0;
        }
        if (need & NEED_AEE) Aee = pf_safe_atan2(n * u, e * r);
        if (need & NEED_ANN) Ann = pf_safe_atan2(e * u, n * r);
#ifndef PF_NO_ATAN_IDENTITY
        if (lockstep || (need & (NEED_AEE | NEED_ANN | NEED_AUU)) == (NEED_AEE | NEED_ANN | NEED_AUU)) {
            // The three arguments a = nu/(er), b = eu/(nr), c = en/(ur) share the sign of
            // e n u and satisfy ab + bc + ca = 1, hence atan a + atan b + atan c = +-pi/2;
            // with exactly one zero coordinate the sum is +-pi/2 by the sign of the other
            // two, with two or more it is 0 (safe_atan2 conventions).  Saves one atan.
            const int he = __double2hiint(e), hn = __double2hiint(n), hu = __double2hiint(u);
            const int zeros = (int)(((he << 1) | __double2loint(e)) == 0) + (int)(((hn << 1) | __double2loint(n)) == 0)
                              + (int)(((hu << 1) | __double2loint(u)) == 0);
            const int sgn = (he ^ hn ^ hu) & 0x80000000;
            double S = __hiloint2double(0x3ff921fb ^ sgn, 0x54442d18);   // +-pi/2
            if (zeros >= 2) S = 0.0;
            Auu = (S - Aee) - Ann;
        } else
#endif
        if (need & NEED_AUU) Auu = pf_safe_atan2(e * n, u * r);
        if (need & NEED_LEN) Len = pf_safe_log(u, e, n, r);
        if (need & NEED_LEU) Leu = pf_safe_log(n, e, u, r);
        if (need & NEED_LNU) Lnu = pf_safe_log(e, n, u, r);
        if (HALF_LOGS && !lockstep) { Len *= 0.5; Leu *= 0.5; Lnu *= 0.5; }   // (ablation builds only)
        k.v0 = -Aee;  // kernel_ee
        k.v1 = -Ann;  // kernel_nn
        k.v2 = -Auu;  // kernel_uu
        k.v3 = Len;   // kernel_en
        k.v4 = Leu;   // kernel_eu
        k.v5 = Lnu;   // kernel_nu
        // choclo.prism.kernel_e / kernel_n / kernel_u
        k.v6 = -(n * Len + u * Leu - e * Aee);
        k.v7 = -(u * Lnu + e * Len - n * Ann);
        k.v8 = -(e * Leu + n * Lnu - u * Auu);
        // gravity/simulation.py:23-30 kernel_uv = 0.5 (kernel_nn - kernel_ee)
        k.v9 = 0.5 * ((-Ann) - (-Aee));
    } else {
        k.v0 = k.v1 = k.v2 = k.v3 = k.v4 = k.v5 = k.v6 = k.v7 = k.v8 = k.v9 = 0.0;
        if (need == 0x3ff) {
            // all ten third-order kernels (tmi_x/y/z, the gradient tensor): straight-line, so that
            // the seven reciprocal refinements and ten quotients interleave.  The six kernel_iij
            // share three denominators (x_k + r) r; per kernel the operations are those of
            // kernel_iii / kernel_iij with pf_div (same bits).
            const double e2 = e * e, n2 = n * n, u2 = u * u;
            const double s_en = e2 + n2, s_eu = e2 + u2, s_nu = n2 + u2;
            const double d_e = s_en * s_eu * r, d_n = s_nu * s_en * r, d_u = s_eu * s_nu * r;
            const double m_e = -n * u * (2 * e2 + n2 + u2);
            const double m_n = -u * e * (2 * n2 + u2 + e2);
            const double m_u = -e * n * (2 * u2 + e2 + n2);
            const double g_u = (u + r) * r, g_n = (n + r) * r, g_e = (e + r) * r;
            const double den[7] = {d_e, d_n, d_u, g_u, g_n, g_e, r};
            double rc[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) rc[i] = pf_rcp_approx(den[i]);
#pragma unroll
            for (int i = 0; i < 7; ++i) { const double t = fma(-den[i], rc[i], 1.0); rc[i] = fma(rc[i], t, rc[i]); }
            const double num[10] = {m_e, m_n, m_u, -e, -e, -n, -n, -u, -u, 1.0};
            const int which[10] = {0, 1, 2, 3, 4, 3, 5, 4, 5, 6};
            double q[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) q[i] = num[i] * rc[which[i]];
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const double rem = fma(-den[which[i]], q[i], num[i]);
                q[i] = fma(rem, rc[which[i]], q[i]);
            }
            const bool ze = e == 0.0, zn = n == 0.0, zu = u == 0.0;
            k.v0 = ((ze && zn) || (ze && zu)) ? 0.0 : q[0];   // eee = iii(e, n, u)
            k.v1 = ((zn && zu) || (zn && ze)) ? 0.0 : q[1];   // nnn = iii(n, u, e)
            k.v2 = ((zu && ze) || (zu && zn)) ? 0.0 : q[2];   // uuu = iii(u, e, n)
            k.v3 = (ze && zn && u < 0.0) ? 0.0 : q[3];        // een = iij(e, n, u)
            k.v4 = (ze && zu && n < 0.0) ? 0.0 : q[4];        // eeu = iij(e, u, n)
            k.v5 = (zn && ze && u < 0.0) ? 0.0 : q[5];        // enn = iij(n, e, u)
            k.v6 = (zn && zu && e < 0.0) ? 0.0 : q[6];        // nnu = iij(n, u, e)
            k.v7 = (zu && ze && n < 0.0) ? 0.0 : q[7];        // euu = iij(u, e, n)
            k.v8 = (zu && zn && e < 0.0) ? 0.0 : q[8];        // nuu = iij(u, n, e)
            k.v9 = (r == 0.0) ? 0.0 : -q[9];                  // enu = -1 / r
            return;
        }
        if (need & (1 << T_EEE)) k.v0 = kernel_iii(e, n, u, r);
        if (need & (1 << T_NNN)) k.v1 = kernel_iii(n, u, e, r);
        if (need & (1 << T_UUU)) k.v2 = kernel_iii(u, e, n, r);
        if (need & (1 << T_EEN)) k.v3 = kernel_iij(e, n, u, r);
        if (need & (1 << T_EEU)) k.v4 = kernel_iij(e, u, n, r);
        if (need & (1 << T_ENN)) k.v5 = kernel_iij(n, e, u, r);
        if (need & (1 << T_NNU)) k.v6 = kernel_iij(n, u, e, r);
        if (need & (1 << T_EUU)) k.v7 = kernel_iij(u, e, n, r);
        if (need & (1 << T_NUU)) k.v8 = kernel_iij(u, n, e, r);
        if (need & (1 << T_ENU)) {
            k.v9 = -pf_div(1.0, r);  // kernel_enu
            if (r == 0.0) k.v9 = 0.0;
        }
    }
}

// choclo.prism.kernel_e / kernel_n / kernel_u by cyclic symmetry: for the axis a with the other two
// coordinates (b, c) in the reference's order
//     k_a = -(b log(c + r) + c log(b + r) - a atan(b c / (a r)))
// (kernel_u: a = u, b = e, c = n; kernel_e: a = e, b = n, c = u; kernel_n: a = n, b = u, c = e)
template <int AXIS>
__device__ __forceinline__ double g1_value(double e, double n, double u)
{
    const double a = (AXIS == 0) ? e : ((AXIS == 1) ? n : u);
    const double b = (AXIS == 0) ? n : ((AXIS == 1) ? u : e);
    const double c = (AXIS == 0) ? u : ((AXIS == 1) ? e : n);
    const double r2 = e * e + n * n + u * u;
    const bool r_is_zero = r2 == 0.0;
    const double r = pf_sqrt(r2, r_is_zero);
    double A, Lc, Lb;
    pf_node_atan1_log2(a, b, c, r, r_is_zero, A, Lc, Lb);
    return -(b * Lc + c * Lb - a * A);
}

// Evaluate the output planes of one (receiver, node) pair.
// G1_AXIS: PROG_G1 evaluates one straight-line copy per axis; -1 = pick the copy here behind a
// uniform branch (marching and node-space kernels), >= 0 = the caller's loop is already specialised
// (the tile kernel, which keeps 16 vertex offsets live across its node loop and spills when the three
// copies sit inside one loop, runs one of three loops).
template <int FAM, int PROG, int STRIDE, int G1_AXIS = -1>
__device__ __forceinline__ void eval_node(const LaunchParams &p, double e, double n, double u,
                                          double *planes, int j)
{
    Raw10 k;
    const double fx = p.fx, fy = p.fy, fz = p.fz;
    if (PF_IS_TMI(PROG)) {
        // one component "tmi" (magnetics/_numba/_3d_mesh.py:352-372), hoisted to the node
        // the three logarithms arrive halved (HALF_LOGS) and meet doubled field coefficients (exact)
        eval_raw<12, true>(NEED_AEE | NEED_ANN | NEED_AUU | NEED_LEN | NEED_LEU | NEED_LNU, e, n, u, k);
        // (explicit FMAs: this combination is hoisted to the node anyway, see the header)
        // the unit factor cf |B0| of the component rides on the (last) contraction with the field
        // direction -- g = factor f, formed on the host -- instead of one multiplication per cell
        // and output
        if (PROG == PROG_TMI_SCALAR) {
            const double bx = fma(k.v4, p.f2z, fma(k.v3, p.f2y, k.v0 * fx));
            const double by = fma(k.v5, p.f2z, fma(k.v1, fy, k.v3 * p.f2x));
            const double bz = fma(k.v2, fz, fma(k.v5, p.f2y, k.v4 * p.f2x));
            planes[j] = fma(bz, p.gz, fma(by, p.gy, bx * p.gx));
        } else {
            planes[j] = fma(k.v4, p.g2z, fma(k.v3, p.g2y, k.v0 * p.gx));
            planes[STRIDE + j] = fma(k.v5, p.g2z, fma(k.v1, p.gy, k.v3 * p.g2x));
            planes[2 * STRIDE + j] = fma(k.v2, p.gz, fma(k.v5, p.g2y, k.v4 * p.g2x));
        }
        return;
    }
    if (PROG == PROG_G1) {
        // one straight-line copy per axis behind a uniform branch: with a run-time axis the (a, b, c)
        // selection cost 18 register moves per node and hid the squares shared with r^2
        double v;
        if (G1_AXIS >= 0) {
            v = g1_value<(G1_AXIS >= 0 ? G1_AXIS : 0)>(e, n, u);     // the caller hoisted the branch
        } else {
            const int axis = p.g1_axis;
            if (axis == 2) v = g1_value<2>(e, n, u);
            else if (axis == 0) v = g1_value<0>(e, n, u);
            else v = g1_value<1>(e, n, u);
        }
        planes[j] = v;
        return;
    }
    eval_raw<FAM>((PROG == PROG_ALL6) ? (NEED_AEE | NEED_ANN | NEED_AUU | NEED_LEN | NEED_LEU | NEED_LNU) : p.need,
                  e, n, u, k);
    // raw planes: straight-line predicated stores (no descriptor loop)
    {
        const int rm = p.raw_mask;   // raw planes occupy slots 0.. in increasing q (host guarantees it)
        double *dst = planes + j;
#define PF_RAW(q, v) if (rm & (1 << q)) { *dst = (v); dst += STRIDE; }
        PF_RAW(0, k.v0) PF_RAW(1, k.v1) PF_RAW(2, k.v2) PF_RAW(3, k.v3) PF_RAW(4, k.v4)
        PF_RAW(5, k.v5) PF_RAW(6, k.v6) PF_RAW(7, k.v7) PF_RAW(8, k.v8) PF_RAW(9, k.v9)
#undef PF_RAW
    }
    // combination planes: k . f (DOT3) and f^T K f (QUAD) are linear in the raw kernels, so each is
    // a short chain of DFMAs with uniform coefficients (predicated on a uniform mask) -- no run-time
    // selection of kernel values (the former pick() switches cost more than the kernels themselves)
    for (int s = p.n_planes - p.n_combo; s < p.n_planes; ++s) {   // raw planes come first
        const double *c = p.combo_coef[s];
        const int m = p.combo_mask[s];
        double val = 0.0;
#define PF_TERM(q, v) if (m & (1 << q)) val = fma(c[q], (v), val);
        PF_TERM(0, k.v0) PF_TERM(1, k.v1) PF_TERM(2, k.v2) PF_TERM(3, k.v3) PF_TERM(4, k.v4)
        PF_TERM(5, k.v5) PF_TERM(6, k.v6) PF_TERM(7, k.v7) PF_TERM(8, k.v8) PF_TERM(9, k.v9)
#undef PF_TERM
        planes[s * STRIDE + j] = val;
    }
}

// The 8 vertices of a cell as BYTE offsets into a plane (kept in registers for
// the whole receiver loop).
struct CellVerts {
    int o0, o1, o2, o3, o4, o5, o6, o7;
};

__device__ __forceinline__ CellVerts unpack_verts(const uint4 ci)
{
    CellVerts v;
    v.o0 = (ci.x & 0xffff) * 8; v.o1 = (ci.x >> 16) * 8;
    v.o2 = (ci.y & 0xffff) * 8; v.o3 = (ci.y >> 16) * 8;
    v.o4 = (ci.z & 0xffff) * 8; v.o5 = (ci.z >> 16) * 8;
    v.o6 = (ci.w & 0xffff) * 8; v.o7 = (ci.w >> 16) * 8;
    return v;
}

// _numba_utils.py:18-45 kernels_in_nodes_to_cell: -k0+k1+k2-k3+k4-k5-k6+k7,
// evaluated left to right.
__device__ __forceinline__ double vertex_sum(const char *plane, const CellVerts &v)
{
#define PF_AT(o) (*reinterpret_cast<const double *>(plane + (o)))
    return -PF_AT(v.o0) + PF_AT(v.o1) + PF_AT(v.o2) - PF_AT(v.o3) + PF_AT(v.o4) - PF_AT(v.o5)
           - PF_AT(v.o6) + PF_AT(v.o7);
#undef PF_AT
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

constexpr int PLANE_BYTES = PF_TILE_NODES * 8;

#ifndef PF_MIN_BLOCKS
#define PF_MIN_BLOCKS 3
#endif

template <int FAM, int PROG, int SINK, typename GT>
__global__ void __launch_bounds__(PF_THREADS, PF_MIN_BLOCKS) pf_tile_kernel(const __grid_constant__ LaunchParams p)
{
    extern __shared__ double smem[];
    double *sxyz = smem;                                         // [3][TN] node coordinates
    double *planes = smem + 3 * PF_TILE_NODES;                   // [n_planes][TN]
    double *wpart = planes + (size_t)p.n_planes * PF_TILE_NODES; // [WARPS][n_fwd] (forward)

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t r_begin = (int64_t)blockIdx.y * p.rc;
    const int64_t r_end = min(r_begin + (int64_t)p.rc, p.n_receivers);
    const int64_t nC = p.plan.n_cells;
    const int nblk = PF_FIXED(PROG) ? PF_NOUT(PROG) : p.nblk;

    // ---- tile-stationary state ------------------------------------------------
    const int nn = p.plan.tile_n_nodes[tile];
    const double *xyz = p.plan.tile_xyz + tile * (3 * PF_TILE_NODES);
    for (int i = tid; i < 3 * PF_TILE_NODES; i += PF_THREADS) sxyz[i] = xyz[i];
    int col[PF_CPT];
    CellVerts verts[PF_CPT];
    double mdl[PF_CPT][3];
    double acc[PF_CPT][3];
#pragma unroll
    for (int s = 0; s < PF_CPT; ++s) {
        const int64_t slot = tile * PF_TILE_CELLS + tid + s * PF_THREADS;
        col[s] = p.plan.tile_cell_col[slot];
        verts[s] = unpack_verts(reinterpret_cast<const uint4 *>(p.plan.tile_cell_nodes)[slot]);
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            acc[s][b] = 0.0;
            mdl[s][b] = 0.0;
            if (SINK == SINK_FORWARD && col[s] >= 0 && b < nblk)
                mdl[s][b] = p.model[col[s] + b * nC];
        }
    }
    __syncthreads();

    for (int64_t r = r_begin; r < r_end; ++r) {
        const double rx = __ldg(p.receivers + 3 * r);
        const double ry = __ldg(p.receivers + 3 * r + 1);
        const double rz = __ldg(p.receivers + 3 * r + 2);

        // ---- phase 1: output planes of the tile nodes -> shared memory -------
        if constexpr (PROG == PROG_G1) {
            const int axis = p.g1_axis;
#define PF_TILE_NODE_LOOP(AX)                                                                          \
            _Pragma("unroll 1") for (int j = tid; j < nn; j += PF_THREADS)                           \
                eval_node<FAM, PROG, PF_TILE_NODES, AX>(p, sxyz[j] - rx, sxyz[PF_TILE_NODES + j] - ry, \
                                                        sxyz[2 * PF_TILE_NODES + j] - rz, planes, j);
            if (axis == 2) { PF_TILE_NODE_LOOP(2) }
            else if (axis == 0) { PF_TILE_NODE_LOOP(0) }
            else { PF_TILE_NODE_LOOP(1) }
#undef PF_TILE_NODE_LOOP
        } else {
#pragma unroll 1
            for (int j = tid; j < nn; j += PF_THREADS)
                eval_node<FAM, PROG, PF_TILE_NODES>(p, sxyz[j] - rx, sxyz[PF_TILE_NODES + j] - ry,
                                                    sxyz[2 * PF_TILE_NODES + j] - rz, planes, j);
        }
        __syncthreads();

        // ---- phase 2: cells ---------------------------------------------------
        const char *pbase = reinterpret_cast<const char *>(planes);
        if (PF_FIXED(PROG)) {
            // single component: row = row_offset + r, factor of the only component
            const int64_t row = p.row_offset + r * p.n_comp_total + p.outs[0].ci;
            const double factor = p.outs[0].factor;
            double rowval = 0.0;
            if (SINK == SINK_GTV || SINK == SINK_DIAG) rowval = __ldg(p.vec + row);
            GT *grow = reinterpret_cast<GT *>(p.G) + row * p.ldG;
            double fsum = 0.0;
#pragma unroll
            for (int s = 0; s < PF_CPT; ++s) {
                if (col[s] < 0) continue;
#pragma unroll
                for (int b = 0; b < PF_NOUT(PROG); ++b) {
                    const double g = factor * vertex_sum(pbase + b * PLANE_BYTES, verts[s]);
                    if (SINK == SINK_STORE) grow[(unsigned)col[s] + (uint64_t)b * nC] = (GT)g;
                    else if (SINK == SINK_FORWARD) fsum += g * mdl[s][b];
                    else if (SINK == SINK_GTV) acc[s][b] += rowval * g;
                    else acc[s][b] += rowval * (g * g);
                }
            }
            if (SINK == SINK_FORWARD) {
                fsum = warp_sum(fsum);
                if ((tid & 31) == 0) wpart[tid >> 5] = fsum;
            }
        } else {
            double fsum = 0.0;
            for (int o = 0; o < p.n_out; ++o) {
                const OutDesc &od = p.outs[o];
                const char *pl = pbase + od.plane * PLANE_BYTES;
                const int64_t row = p.row_offset + r * p.n_comp_total + od.ci;
                double rowval = 0.0;  // v[row] or w[row]
                if (SINK == SINK_GTV || SINK == SINK_DIAG) rowval = __ldg(p.vec + row);
                GT *grow = reinterpret_cast<GT *>(p.G) + row * p.ldG + od.blk * nC;
#pragma unroll
                for (int s = 0; s < PF_CPT; ++s) {
                    if (col[s] < 0) continue;
                    // gravity/_numba/_3d_mesh.py:159-163: constant_factor * vertex sum
                    const double g = od.factor * vertex_sum(pl, verts[s]);
                    if (SINK == SINK_STORE) {
                        grow[(unsigned)col[s]] = (GT)g;
                    } else if (SINK == SINK_FORWARD) {
                        const double m = (od.blk == 0) ? mdl[s][0] : ((od.blk == 1) ? mdl[s][1] : mdl[s][2]);
                        fsum += g * m;
                    } else {
                        const double t = (SINK == SINK_GTV) ? rowval * g : rowval * (g * g);
                        if (od.blk == 0) acc[s][0] += t;
                        else if (od.blk == 1) acc[s][1] += t;
                        else acc[s][2] += t;
                    }
                }
                if (SINK == SINK_FORWARD && od.last) {
                    fsum = warp_sum(fsum);
                    if ((tid & 31) == 0) wpart[(tid >> 5) * p.n_fwd + od.fi] = fsum;
                    fsum = 0.0;
                }
            }
        }
        __syncthreads();
        if (SINK == SINK_FORWARD && tid < p.n_fwd) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += wpart[w * p.n_fwd + tid];
            const int64_t n_local = p.n_receivers * p.n_fwd;
            p.partial[tile * n_local + r * p.n_fwd + tid] = t;
        }
    }

    if (SINK == SINK_GTV || SINK == SINK_DIAG) {
        const int64_t nP = nC * nblk;
        double *out = p.partial + (int64_t)blockIdx.y * nP;
#pragma unroll
        for (int s = 0; s < PF_CPT; ++s) {
            if (col[s] < 0) continue;
            out[col[s]] = acc[s][0];
            if (nblk == 3) {
                out[col[s] + nC] = acc[s][1];
                out[col[s] + 2 * nC] = acc[s][2];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// z-marching kernel for tensor meshes (SINK_STORE, SINK_FORWARD).
//
// One CTA = one (x, y) footprint of the mesh x one chunk of receivers.  For
// every receiver the CTA sweeps the node layers bottom-up.  Per layer L:
//   phase 1  the output planes of the layer's nodes are evaluated into one of
//            two alternating shared-memory buffers;
//   phase 2  every cell column forms the corner difference of the layer,
//                D_L = k(x0,y0) - k(x1,y0) - k(x0,y1) + k(x1,y1),
//            and emits the cell between layers L-1 and L as D_L - D_{L-1}
//            (= -k0+k1+k2-k3+k4-k5-k6+k7 of _numba_utils.py:18-45, associated
//            layer by layer), keeping D_L in a thread-private shared slot for
//            the next slab: 4 instead of 8 gathers per cell and output, and a
//            node layer is evaluated once for the two slabs that touch it.
// One __syncthreads per layer.  Nodes evaluated per cell:
// (fw+1)(fh+1)/(fw fh) x (nz+1)/nz ~ 1.1-1.2 (bricks: 1.45-1.5).
// ---------------------------------------------------------------------------
#ifndef PF_MARCH_MIN_BLOCKS
#define PF_MARCH_MIN_BLOCKS 3
#endif

// shared-memory accesses through 32-bit shared-window addresses with compile-time offsets: the
// cell phase forms one address per slot and everything else is an immediate (left to itself the
// compiler re-derives the CTA's shared window and the buffer bases for every slot)
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v));
}

// sign of a node's atan sum S = Aee + Ann + Auu as an integer: +-1 for +-pi/2, 0 when two or
// more coordinates vanish (the safe_atan2 conventions, see eval_raw)
__device__ __forceinline__ int node_S_sign(double e, double n, double u)
{
    const int zeros = (int)(e == 0.0) + (int)(n == 0.0) + (int)(u == 0.0);
    const int sb = (__double2hiint(e) ^ __double2hiint(n) ^ __double2hiint(u)) & 0x80000000;
    return zeros >= 2 ? 0 : (sb ? -1 : 1);
}

template <int FAM, int PROG, int SINK, typename GT, int NR, int MT = PF_THREADS>
__global__ void __launch_bounds__(MT, PF_MARCH_MIN_BLOCKS * (PF_THREADS / MT))
    pf_march_kernel(const __grid_constant__ LaunchParams p)
{
    constexpr int NCAP = NR * MT;   // node slots per layer
    constexpr int LAYER_BYTES = NCAP * 8;
    extern __shared__ double smem[];
    const int n_planes = p.n_planes;
    const int n_out = PF_FIXED(PROG) ? PF_NOUT(PROG) : p.n_out;
    // outputs [0, n_plain) read a plane and own a dprev slot; [n_plain, n_out) are derived
    const int n_plain = PF_FIXED(PROG) ? PF_NOUT(PROG) : p.n_out - p.n_derived;
    const int buf_doubles = n_planes * NCAP;
    const int n_buf = p.n_buf;    // 1 when two buffers would not leave room for 3 CTAs per SM
    double *dprev = smem + n_buf * (size_t)buf_doubles;       // [n_plain][NR][THREADS] corner differences
    double *wacc = dprev + (size_t)n_plain * NR * MT; // [(MT / 32)][n_fwd] (forward)
    // per-output bookkeeping, refreshed once per receiver: destination of column 0 of
    // the output's row/block (store) or of its model block (forward), its factor and role
    __shared__ unsigned long long s_dst[MAX_OUT];
    __shared__ double s_fac[MAX_OUT];
    __shared__ int s_role[MAX_OUT];

    const int tid = threadIdx.x;
    const unsigned smem32 = (unsigned)__cvta_generic_to_shared(smem);          // shared-window addresses
    const unsigned dprev32 = (unsigned)__cvta_generic_to_shared(dprev) + tid * 8;
    const int nx = p.plan.grid_nx, ny = p.plan.grid_ny;
    const int32_t *fp = ((MT != PF_THREADS) ? p.plan.footprints4h
                                            : ((NR == 4) ? p.plan.footprints4 : p.plan.footprints2)) + 6 * blockIdx.x;
    const int ix0 = fp[0], iy0 = fp[1], fw = fp[2], fh = fp[3];
    const int iz_begin = fp[4], iz_end = fp[5];   // cell layers that hold active cells
    const int W1 = fw + 1;
    const int n_layer_nodes = W1 * (fh + 1);
    const int64_t r_begin = (int64_t)blockIdx.y * p.rc;
    const int64_t r_end = min(r_begin + (int64_t)p.rc, p.n_receivers);
    const int64_t nC = p.plan.n_cells;
    const int64_t layer_stride = (int64_t)nx * ny;
    const int sW = W1 * 8;  // byte stride of one node row inside a layer plane
    const bool direct = PF_FIXED(PROG) || p.direct;
    const bool has_derived = !PF_FIXED(PROG) && n_out > n_plain;

    // ---- footprint-stationary state -------------------------------------------
    // Node phase: the layer's nodes are taken R = THREADS / W1 whole rows at a time, so thread tid
    // owns node column lx = tid % W1 in rows tid / W1 + R k -- plane index tid + k R W1, the same
    // x for every node of the thread (e = x - rx is formed once per receiver and its square hoisted
    // out of the layer loop; the y of a round is one L1 load).  Footprints wider than THREADS - 1
    // cells are a plan error (include/pf_cuda.h; tiling.march_footprints never makes them).
    if (W1 > MT) __trap();
    const int RW = (MT / W1) * W1;            // node slots per round
    const int rows_per_round = MT / W1;
    const bool node_thread = tid < RW;
    const double xnode = __ldg(p.plan.grid_xs + ix0 + (node_thread ? tid % W1 : 0));
    const double *yrow = p.plan.grid_ys + iy0 + (node_thread ? tid / W1 : 0);
    const int n_layer_cells = fw * fh;
    int o00[NR];       // byte offset of the cell's (x0, y0) node inside a layer plane
    int cbase[NR];     // mesh index of the cell in layer 0, -1 = no cell in this slot
#pragma unroll
    for (int s = 0; s < NR; ++s) {
        const int c = tid + s * MT;
        const int cy = c / fw, cx = c - cy * fw;
        const bool ok = c < n_layer_cells;
        o00[s] = (cx + W1 * cy) * 8;
        cbase[s] = ok ? (ix0 + cx) + nx * (iy0 + cy) : -1;
    }
    if (SINK == SINK_FORWARD) {
        for (int i = tid; i < (MT / 32) * p.n_fwd; i += MT) wacc[i] = 0.0;
    }
    // closed (x, y) bounding box of the footprint's nodes (derived outputs only)
    double box_x0 = 0, box_x1 = 0, box_y0 = 0, box_y1 = 0;
    if (has_derived) {
        box_x0 = __ldg(p.plan.grid_xs + ix0);
        box_x1 = __ldg(p.plan.grid_xs + ix0 + fw);
        box_y0 = __ldg(p.plan.grid_ys + iy0);
        box_y1 = __ldg(p.plan.grid_ys + iy0 + fh);
    }

    for (int64_t r = r_begin; r < r_end; ++r) {
        const double rx = __ldg(p.receivers + 3 * r);
        const double ry = __ldg(p.receivers + 3 * r + 1);
        const double rz = __ldg(p.receivers + 3 * r + 2);
        if (tid < n_out) {
            const OutDesc &od = p.outs[tid];
            const int64_t row = p.row_offset + r * p.n_comp_total + od.ci;
            s_dst[tid] = (SINK == SINK_STORE)
                             ? (unsigned long long)(reinterpret_cast<GT *>(p.G) + row * p.ldG + (int64_t)od.blk * nC)
                             : (unsigned long long)(p.model + (int64_t)od.blk * nC);
            s_fac[tid] = od.factor;
            s_role[tid] = od.role;
        }
        // sum_v s_v S_v of a cell can only be non-zero when the receiver lies inside or on the
        // closed prism: a CTA-uniform test per receiver (and per layer below) guards the exact path
        const bool rx_in_box = has_derived && rx >= box_x0 && rx <= box_x1 && ry >= box_y0 && ry <= box_y1;

        // ---- phase 1: output planes of node layer L -> planes ---------------------------
        const double e = xnode - rx;
        auto phase1 = [&](int L, double *planes) {
            const double u = __ldg(p.plan.grid_zs + L) - rz;
            if (node_thread) {
                const double *yp = yrow;
#pragma unroll 1
                for (int j = tid; j < n_layer_nodes; j += RW, yp += rows_per_round)
                    eval_node<FAM, PROG, NCAP>(p, e, __ldg(yp) - ry, u, planes, j);
            }
        };
        // exact sum_v s_v S_v of the cell in slot s between node layers L-1 and L (rare path)
        auto cell_S = [&](int s, int L) -> double {
            const int c = tid + s * MT;
            const int cy = c / fw, cx = c - cy * fw;
            const double e0 = __ldg(p.plan.grid_xs + ix0 + cx) - rx, e1 = __ldg(p.plan.grid_xs + ix0 + cx + 1) - rx;
            const double n0 = __ldg(p.plan.grid_ys + iy0 + cy) - ry, n1 = __ldg(p.plan.grid_ys + iy0 + cy + 1) - ry;
            const double u0 = __ldg(p.plan.grid_zs + L - 1) - rz, u1 = __ldg(p.plan.grid_zs + L) - rz;
            const int k = -node_S_sign(e0, n0, u0) + node_S_sign(e1, n0, u0) + node_S_sign(e0, n1, u0)
                          - node_S_sign(e1, n1, u0) + node_S_sign(e0, n0, u1) - node_S_sign(e1, n0, u1)
                          - node_S_sign(e0, n1, u1) + node_S_sign(e1, n1, u1);
            return (double)k * 1.5707963267948966;
        };
        // ---- phase 2 for passes with derived outputs (gravity gzz from gxx and gyy): output-major,
        //      plain outputs first (they fill keep_a / keep_b), then the derived ones -------------
        auto phase2_derived = [&](int L, const double *planes) {
            const char *pl = reinterpret_cast<const char *>(planes);
            const bool emit = L > iz_begin;
            int col[NR];
            const int layer_off = (int)(layer_stride * (L - 1));
            const int32_t *colmap = p.plan.grid_colmap;
#pragma unroll
            for (int s = 0; s < NR; ++s) {
                col[s] = -1;
                if (emit && cbase[s] >= 0) {
                    const int idx = cbase[s] + layer_off;
                    col[s] = colmap ? __ldg(colmap + idx) : idx;
                }
            }
            // sum_v s_v S_v of a cell can only be non-zero when the receiver lies inside or on the
            // closed prism: CTA-uniform guard (footprint box per receiver, z range per layer)
            bool needS = false;
            if (rx_in_box && emit) needS = rz >= __ldg(p.plan.grid_zs + L - 1) && rz <= __ldg(p.plan.grid_zs + L);
            double keep_a[NR], keep_b[NR];
#pragma unroll
            for (int s = 0; s < NR; ++s) keep_a[s] = keep_b[s] = 0.0;
            double fsum = 0.0;
#define PF_PHASE2_PLAIN(O, POFF)                                                                  \
            {                                                                                     \
                GT *grow = reinterpret_cast<GT *>(s_dst[O]);                                      \
                const double *mrow = reinterpret_cast<const double *>(s_dst[O]);                  \
                const double factor = s_fac[O];                                                   \
                const int role = s_role[O];                                                       \
                double *dp = dprev + (size_t)(O) * (NR * MT) + tid;                       \
                _Pragma("unroll") for (int s = 0; s < NR; ++s) {                                  \
                    if (cbase[s] < 0) continue;                                                   \
                    const char *q = pl + (POFF) + o00[s];                                         \
                    const double d_new = *reinterpret_cast<const double *>(q)                     \
                                         - *reinterpret_cast<const double *>(q + 8)               \
                                         - *reinterpret_cast<const double *>(q + sW)              \
                                         + *reinterpret_cast<const double *>(q + sW + 8);         \
                    const double d_old = dp[s * MT];                                      \
                    dp[s * MT] = d_new;                                                   \
                    const double g = factor * (d_new - d_old);                                    \
                    if (role == ROLE_KEEP_A) keep_a[s] = g;                                       \
                    if (role == ROLE_KEEP_B) keep_b[s] = g;                                       \
                    if (col[s] < 0) continue;                                                     \
                    if (SINK == SINK_STORE) grow[(unsigned)col[s]] = (GT)g;                       \
                    else fsum += g * __ldg(mrow + (unsigned)col[s]);                              \
                }                                                                                 \
                if (SINK == SINK_FORWARD && emit && p.outs[O].last) {                             \
                    fsum = warp_sum(fsum);                                                        \
                    if ((tid & 31) == 0) wacc[(tid >> 5) * p.n_fwd + p.outs[O].fi] += fsum;       \
                    fsum = 0.0;                                                                   \
                }                                                                                 \
            }
            if (direct) {
#pragma unroll
                for (int o = 0; o < MAX_PLANES; ++o) {
                    if (o >= n_plain) break;
                    PF_PHASE2_PLAIN(o, o * LAYER_BYTES)
                }
            } else {
                for (int o = 0; o < n_plain; ++o) {
                    const int poff = p.outs[o].plane * LAYER_BYTES;
                    PF_PHASE2_PLAIN(o, poff)
                }
            }
#undef PF_PHASE2_PLAIN
            for (int o = n_plain; o < n_out; ++o) {
                GT *grow = reinterpret_cast<GT *>(s_dst[o]);
                const double *mrow = reinterpret_cast<const double *>(s_dst[o]);
                const double factor = s_fac[o];
#pragma unroll
                for (int s = 0; s < NR; ++s) {
                    if (col[s] < 0) continue;
                    double g = -(keep_a[s] + keep_b[s]);
                    if (needS) g -= factor * cell_S(s, L);
                    if (SINK == SINK_STORE) grow[(unsigned)col[s]] = (GT)g;
                    else fsum += g * __ldg(mrow + (unsigned)col[s]);
                }
                if (SINK == SINK_FORWARD && emit) {
                    fsum = warp_sum(fsum);
                    if ((tid & 31) == 0) wacc[(tid >> 5) * p.n_fwd + p.outs[o].fi] += fsum;
                    fsum = 0.0;
                }
            }
        };
        // ---- phase 2: corner differences of layer L; emits the cells between L-1 and L ----
        double fsum_rx = 0.0;    // forward sink of the fixed-output programs: sum over the receiver's layers
        auto phase2 = [&](int L, const double *planes) {
            const char *pl = reinterpret_cast<const char *>(planes);
            const bool emit = L > iz_begin;
            int col[NR];
            const int layer_off = (int)(layer_stride * (L - 1));   // mesh cells < 2^31 (validate_grid)
            const int32_t *colmap = p.plan.grid_colmap;
            if (colmap) {
#pragma unroll
                for (int s = 0; s < NR; ++s) {
                    const bool ok = emit && cbase[s] >= 0;
                    col[s] = ok ? __ldg(colmap + (ok ? cbase[s] + layer_off : 0)) : -1;
                }
            } else {
#pragma unroll
                for (int s = 0; s < NR; ++s) col[s] = (emit && cbase[s] >= 0) ? cbase[s] + layer_off : -1;
            }
            double fsum = 0.0;
            if constexpr (PF_FIXED(PROG)) {
                if (SINK == SINK_FORWARD) fsum = fsum_rx;     // one accumulator per thread and receiver
                // one component, PF_NOUT outputs = planes = column blocks, one factor.  Slot-major; the
                // <= 3 row pointers and the factor are read once per layer, so that an output costs
                // 5 shared loads, 1 shared store, 5 FP64 and a 2-instruction global store
                constexpr int NO = PF_NOUT(PROG);
                const double factor = s_fac[0];
                unsigned long long dst[NO];
#pragma unroll
                for (int o = 0; o < NO; ++o) dst[o] = s_dst[o];
                // formed once per layer and made opaque: the compiler otherwise rematerialises both
                // bases (shared window, buffer parity, thread index) in front of every slot
                unsigned pl32 = smem32 + (unsigned)(pl - reinterpret_cast<const char *>(smem));
                unsigned dp32 = dprev32;
                asm volatile("" : "+r"(pl32), "+r"(dp32));
#pragma unroll
                for (int s = 0; s < NR; ++s) {
                    if (cbase[s] < 0) continue;
                    const unsigned q0 = pl32 + o00[s];       // node (x0, y0) of the cell, plane 0
                    const unsigned q1 = q0 + sW;             // node (x0, y1)
                    const bool st = col[s] >= 0;
                    const unsigned cu = (unsigned)col[s];
#define PF_FIXED_OUT(O)                                                                              \
                    if (O < NO) {                                                                     \
                        const double d_new = lds_f64<(O) * LAYER_BYTES>(q0) - lds_f64<(O) * LAYER_BYTES + 8>(q0) \
                                             - lds_f64<(O) * LAYER_BYTES>(q1) + lds_f64<(O) * LAYER_BYTES + 8>(q1); \
                        const double d_old = (s == 0)   ? lds_f64<((O) * NR + 0) * MT * 8>(dp32) \
                                             : (s == 1) ? lds_f64<((O) * NR + 1 % NR) * MT * 8>(dp32) \
                                             : (s == 2) ? lds_f64<((O) * NR + 2 % NR) * MT * 8>(dp32) \
                                                        : lds_f64<((O) * NR + 3 % NR) * MT * 8>(dp32); \
                        if (s == 0) sts_f64<((O) * NR + 0) * MT * 8>(dp32, d_new);            \
                        else if (s == 1) sts_f64<((O) * NR + 1 % NR) * MT * 8>(dp32, d_new);  \
                        else if (s == 2) sts_f64<((O) * NR + 2 % NR) * MT * 8>(dp32, d_new);  \
                        else sts_f64<((O) * NR + 3 % NR) * MT * 8>(dp32, d_new);              \
                        const double g = PF_IS_TMI(PROG) ? (d_new - d_old) : factor * (d_new - d_old); \
                        if (SINK == SINK_STORE) {                                                     \
                            if (st) reinterpret_cast<GT *>(dst[(O) < NO ? (O) : 0])[cu] = (GT)g;      \
                        } else {                                                                      \
                            if (st) fsum += g * __ldg(reinterpret_cast<const double *>(dst[(O) < NO ? (O) : 0]) + cu); \
                        }                                                                             \
                    }
                    PF_FIXED_OUT(0)
                    PF_FIXED_OUT(1)
                    PF_FIXED_OUT(2)
#undef PF_FIXED_OUT
                }
                // (the warp reduction happens once per receiver, after the last layer)
                if (SINK == SINK_FORWARD) fsum_rx = fsum;
                return;
            }
#define PF_PHASE2_BODY(O, POFF)                                                                   \
            {                                                                                     \
                GT *grow = reinterpret_cast<GT *>(s_dst[O]);                                      \
                const double *mrow = reinterpret_cast<const double *>(s_dst[O]);                  \
                const double factor = s_fac[O];                                                   \
                double *dp = dprev + (size_t)(O) * (NR * MT) + tid;                       \
                _Pragma("unroll") for (int s = 0; s < NR; ++s) {                                  \
                    if (cbase[s] < 0) continue;                                                   \
                    const char *q = pl + (POFF) + o00[s];                                         \
                    const double d_new = *reinterpret_cast<const double *>(q)                     \
                                         - *reinterpret_cast<const double *>(q + 8)               \
                                         - *reinterpret_cast<const double *>(q + sW)              \
                                         + *reinterpret_cast<const double *>(q + sW + 8);         \
                    const double d_old = dp[s * MT];                                      \
                    dp[s * MT] = d_new;                                                   \
                    if (col[s] < 0) continue;                                                     \
                    const double g = factor * (d_new - d_old);                                    \
                    if (SINK == SINK_STORE) grow[(unsigned)col[s]] = (GT)g;                       \
                    else fsum += g * __ldg(mrow + (unsigned)col[s]);                              \
                }                                                                                 \
            }
            if (direct) {
                // output o reads plane o: unrolled with compile-time plane offsets
#pragma unroll
                for (int o = 0; o < MAX_PLANES; ++o) {
                    if (o >= n_out) break;
                    PF_PHASE2_BODY(o, o * LAYER_BYTES)
                    if (SINK == SINK_FORWARD && emit && (PF_FIXED(PROG) ? (o == n_out - 1) : p.outs[o].last)) {
                        fsum = warp_sum(fsum);
                        const int fi = PF_FIXED(PROG) ? 0 : p.outs[o].fi;
                        if ((tid & 31) == 0) wacc[(tid >> 5) * p.n_fwd + fi] += fsum;
                        fsum = 0.0;
                    }
                }
            } else {
                for (int o = 0; o < n_out; ++o) {
                    const int poff = p.outs[o].plane * LAYER_BYTES;
                    PF_PHASE2_BODY(o, poff)
                    if (SINK == SINK_FORWARD && emit && p.outs[o].last) {
                        fsum = warp_sum(fsum);
                        if ((tid & 31) == 0) wacc[(tid >> 5) * p.n_fwd + p.outs[o].fi] += fsum;
                        fsum = 0.0;
                    }
                }
            }
#undef PF_PHASE2_BODY
        };

        // node layers iz_begin .. iz_end; the cells between layers L-1 and L are emitted at L
        for (int L = iz_begin; L <= iz_end; ++L) {
            double *planes = smem + ((n_buf == 2) ? (L & 1) * buf_doubles : 0);
#ifndef PF_ABLATE_P1      // timing experiments only (results are garbage)
            phase1(L, planes);
#endif
            __syncthreads();
#ifndef PF_ABLATE_P2
            if (has_derived) phase2_derived(L, planes);
            else phase2(L, planes);
#endif
            if (n_buf == 1) __syncthreads();
        }
        if (SINK == SINK_FORWARD && PF_FIXED(PROG)) {
            fsum_rx = warp_sum(fsum_rx);
            if ((tid & 31) == 0) wacc[(tid >> 5) * p.n_fwd] += fsum_rx;
        }
        __syncthreads();   // all warps done with the last layer (buffers, dprev and wacc are reused)
        if (SINK == SINK_FORWARD) {
            if (tid < p.n_fwd) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < (MT / 32); ++w) {
                    t += wacc[w * p.n_fwd + tid];
                    wacc[w * p.n_fwd + tid] = 0.0;
                }
                const int64_t n_local = p.n_receivers * p.n_fwd;
                p.partial[(int64_t)blockIdx.x * n_local + r * p.n_fwd + tid] = t;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------
// Node-space G^T v (on-the-fly Jtvec).
//
//   (G^T v)[cell, b] = sum_r v_r f sum_{8 vertices} s_vertex t_b(node - receiver_r)
//                    = sum_{8 vertices} s_vertex N_b[node],
//   N_b[node] = sum_r (v_r f) t_b(node - receiver_r)
//
// because the vertex combination is linear: the kernels are accumulated once
// per UNIQUE node (nN/nC ~ 1.03 evaluations per pair instead of ~1.5 with
// tiles), one thread per node, no shared-memory exchange, no barrier; the
// signed 8-vertex sum (_numba_utils.py:18-45) is applied once at the end.
// Accumulation over receivers is Kahan-compensated (N_b is ~1e4 x larger than
// the vertex sums it cancels to).
// ---------------------------------------------------------------------------
template <int FAM, int PROG>
__global__ void __launch_bounds__(PF_THREADS, 3)
    pf_node_gtv_kernel(const __grid_constant__ LaunchParams p, double *partial /* [chunks][nblk][n_nodes] */)
{
    extern __shared__ double vals[];   // [n_planes][THREADS], thread-private columns
    const int tid = threadIdx.x;
    const int64_t n_nodes = p.plan.n_nodes;
    const int64_t j = (int64_t)blockIdx.x * PF_THREADS + tid;
    const bool valid = j < n_nodes;
    const int64_t jj = valid ? j : 0;
    const double x = p.plan.node_xyz[jj];
    const double y = p.plan.node_xyz[n_nodes + jj];
    const double z = p.plan.node_xyz[2 * n_nodes + jj];
    const int nblk = PF_FIXED(PROG) ? PF_NOUT(PROG) : p.nblk;
    const int n_out = PF_FIXED(PROG) ? PF_NOUT(PROG) : p.n_out;
    const int64_t r_begin = (int64_t)blockIdx.y * p.rc;
    const int64_t r_end = min(r_begin + (int64_t)p.rc, p.n_receivers);
    double a0 = 0, a1 = 0, a2 = 0, c0 = 0, c1 = 0, c2 = 0;   // Kahan sums per column block
    for (int64_t r = r_begin; r < r_end; ++r) {
        const double rx = __ldg(p.receivers + 3 * r);
        const double ry = __ldg(p.receivers + 3 * r + 1);
        const double rz = __ldg(p.receivers + 3 * r + 2);
        if constexpr (PF_FIXED(PROG)) {
            // one component: one row, one coefficient per receiver; the <= 3 plane values stay in
            // registers (no shared-memory round trip, no per-output descriptor reads)
            const int64_t row = p.row_offset + r * p.n_comp_total + p.outs[0].ci;
            const double coef = __ldg(p.vec + row) * p.outs[0].factor;
            double loc[3] = {0.0, 0.0, 0.0};
            eval_node<FAM, PROG, 1>(p, x - rx, y - ry, z - rz, loc, 0);
            { const double yk = coef * loc[0] - c0, t = a0 + yk; c0 = (t - a0) - yk; a0 = t; }
            if (PF_NOUT(PROG) == 3) {
                { const double yk = coef * loc[1] - c1, t = a1 + yk; c1 = (t - a1) - yk; a1 = t; }
                { const double yk = coef * loc[2] - c2, t = a2 + yk; c2 = (t - a2) - yk; a2 = t; }
            }
            continue;
        }
        eval_node<FAM, PROG, PF_THREADS>(p, x - rx, y - ry, z - rz, vals, tid);
        for (int o = 0; o < n_out; ++o) {
            const OutDesc &od = p.outs[o];
            const int64_t row = p.row_offset + r * p.n_comp_total + od.ci;
            const double coef = __ldg(p.vec + row) * od.factor;
            const int plane = PF_FIXED(PROG) ? o : od.plane;
            const int blk = PF_FIXED(PROG) ? o : od.blk;
            const double term = coef * vals[plane * PF_THREADS + tid];
            if (blk == 0) { const double yk = term - c0, t = a0 + yk; c0 = (t - a0) - yk; a0 = t; }
            else if (blk == 1) { const double yk = term - c1, t = a1 + yk; c1 = (t - a1) - yk; a1 = t; }
            else { const double yk = term - c2, t = a2 + yk; c2 = (t - a2) - yk; a2 = t; }
        }
    }
    if (valid) {
        double *out = partial + (int64_t)blockIdx.y * nblk * n_nodes + j;
        out[0] = a0;
        if (nblk == 3) {
            out[n_nodes] = a1;
            out[2 * n_nodes] = a2;
        }
    }
}

// N[i] = sum_chunks partial[chunk][i]
__global__ void __launch_bounds__(256) pf_node_reduce_kernel(const double *partial, int64_t n_chunks,
                                                             int64_t n, double *N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = 0.0;
    for (int64_t c = 0; c < n_chunks; ++c) t += partial[c * n + i];
    N[i] = t;
}

// result[cell + b nC] += -N0+N1+N2-N3+N4-N5-N6+N7 over the cell's vertices
__global__ void __launch_bounds__(256) pf_node_to_cells_kernel(const double *N, int64_t n_nodes,
                                                               const int32_t *cell_nodes, int64_t nC,
                                                               int nblk, double *result)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int4 lo = reinterpret_cast<const int4 *>(cell_nodes)[2 * c];
    const int4 hi = reinterpret_cast<const int4 *>(cell_nodes)[2 * c + 1];
    for (int b = 0; b < nblk; ++b) {
        const double *Nb = N + (int64_t)b * n_nodes;
        const double vs = -Nb[lo.x] + Nb[lo.y] + Nb[lo.z] - Nb[lo.w] + Nb[hi.x] - Nb[hi.y] - Nb[hi.z] + Nb[hi.w];
        result[c + (int64_t)b * nC] += vs;
    }
}

// fields[row_offset + r*n_comp_total + ci[f]] += sum_tiles partial[tile][r*n_fwd + f]
struct ReduceFwdParams {
    const double *partial;
    int64_t n_tiles;
    int64_t n_receivers;
    int n_fwd;
    int n_comp_total;
    int64_t row_offset;
    int ci[PF_MAX_COMPONENTS];
    double *fields;
};

__global__ void __launch_bounds__(256) pf_reduce_forward_kernel(const __grid_constant__ ReduceFwdParams q)
{
    const int64_t n_local = q.n_receivers * q.n_fwd;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    double t = 0.0;
    for (int64_t tile = 0; tile < q.n_tiles; ++tile) t += q.partial[tile * n_local + i];
    const int64_t r = i / q.n_fwd;
    const int f = (int)(i - r * q.n_fwd);
    q.fields[q.row_offset + r * q.n_comp_total + q.ci[f]] += t;
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
// host side: job -> passes -> launches
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
        // gravity/simulation.py:74-85 CHOCLO_KERNELS
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
    // magnetics/simulation.py:45-103 CHOCLO_KERNELS; tmi: _numba/_3d_mesh.py:337-342
    // six-kernel sets are ordered (xx, yy, zz, xy, xz, yz)
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

int need_of_quantity(int family, int q)
{
    if (family == 3) return 1 << q;
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
    PF_CHECK_ARG(plan->n_tiles == 0 ||
                     (plan->tile_cells == PF_TILE_CELLS && plan->tile_nodes == PF_TILE_NODES),
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

// One pass = one launch of the tile kernel: a subset of the job's components of
// one kernel family whose output planes fit in shared memory.
struct Pass {
    int family;
    int prog;
    LaunchParams p;
};

int find_or_add_plane(LaunchParams &p, int kind, const int *q, int nq)
{
    for (int s = 0; s < p.n_planes; ++s) {
        if (p.planes[s].kind != kind) continue;
        bool same = true;
        for (int k = 0; k < nq; ++k) same = same && (p.planes[s].q[k] == q[k]);
        if (same) return s;
    }
    if (p.n_planes == MAX_PLANES) return -1;
    PlaneDesc &pd = p.planes[p.n_planes];
    pd.kind = kind;
    for (int k = 0; k < 6; ++k) pd.q[k] = (k < nq) ? q[k] : 0;
    return p.n_planes++;
}

// Try to add component c of the job to the pass; false if it does not fit.
bool add_component(const pf_job *job, int c, const CompInfo &ci, double amp, LaunchParams &p)
{
    const LaunchParams saved = p;
    const int nblk = p.nblk;
    double factor;
    if (job->physics == PF_GRAVITY) {
        // gravity/simulation.py:101-111 and 580: constants.G * conversion_factor
        factor = PF_NEWTON_G * ((job->components[c] <= PF_GZ) ? 1e8 : 1e12);
    } else {
        // magnetics/simulation.py:958 constant_factor = 1/4/pi, times |B0|
        factor = (1.0 / 4.0 / 3.141592653589793) * amp;
    }
    int plane_of_blk[3] = {-1, -1, -1};
    if (ci.type == CT_SINGLE) {
        plane_of_blk[0] = find_or_add_plane(p, PK_RAW, ci.q, 1);
    } else if (ci.type == CT_MAG3) {
        if (nblk == 1) {
            plane_of_blk[0] = find_or_add_plane(p, PK_DOT3, ci.q, 3);
        } else {
            for (int b = 0; b < 3; ++b) plane_of_blk[b] = find_or_add_plane(p, PK_RAW, &ci.q[b], 1);
        }
    } else {
        const int xx = ci.q[0], yy = ci.q[1], zz = ci.q[2], xy = ci.q[3], xz = ci.q[4], yz = ci.q[5];
        if (nblk == 1) {
            plane_of_blk[0] = find_or_add_plane(p, PK_QUAD, ci.q, 6);
        } else {
            const int rows[3][3] = {{xx, xy, xz}, {xy, yy, yz}, {xz, yz, zz}};
            for (int b = 0; b < 3; ++b) plane_of_blk[b] = find_or_add_plane(p, PK_DOT3, rows[b], 3);
        }
    }
    bool ok = (p.n_out + nblk <= MAX_OUT) && (p.n_fwd < PF_MAX_COMPONENTS);
    for (int b = 0; b < nblk; ++b) ok = ok && (plane_of_blk[b] >= 0);
    if (!ok) {
        p = saved;
        return false;
    }
    for (int b = 0; b < nblk; ++b) {
        OutDesc &od = p.outs[p.n_out++];
        od.plane = plane_of_blk[b];
        od.ci = c;
        od.blk = b;
        od.fi = p.n_fwd;
        od.last = (b == nblk - 1);
        od.role = ROLE_PLAIN;
        od.factor = factor;
    }
    p.n_fwd++;
    for (int k = 0; k < ci.nq; ++k) p.need |= need_of_quantity(ci.family, ci.q[k]);
    return true;
}

void init_pass(const pf_plan *plan, const pf_job *job, double amp, Pass &pass, int family)
{
    std::memset(&pass, 0, sizeof(pass));
    pass.family = family;
    LaunchParams &p = pass.p;
    p.plan = *plan;
    p.receivers = job->receivers;
    p.n_receivers = job->n_receivers;
    p.n_comp_total = job->n_components;
    p.row_offset = job->row_offset;
    p.nblk = n_blocks_of(job);
    if (job->physics == PF_MAGNETICS) {
        // magnetics/_numba/_3d_mesh.py:165-169
        p.fx = job->b0[0] / amp;
        p.fy = job->b0[1] / amp;
        p.fz = job->b0[2] / amp;
    }
}

// Split the job into passes (usually one per kernel family).  allow_derived: gravity gzz is
// produced from the gxx and gyy outputs of the same pass instead of a plane of its own (ROLE_*;
// only the marching kernel understands derived outputs).
int make_passes(const pf_plan *plan, const pf_job *job, Pass *passes, int max_passes, bool allow_derived = false)
{
    int c_gxx = -1, c_gyy = -1, c_gzz = -1;
    if (allow_derived && job->physics == PF_GRAVITY) {
        for (int c = 0; c < job->n_components; ++c) {
            if (job->components[c] == PF_GXX && c_gxx < 0) c_gxx = c;
            if (job->components[c] == PF_GYY && c_gyy < 0) c_gyy = c;
            if (job->components[c] == PF_GZZ && c_gzz < 0) c_gzz = c;
        }
    }
    const bool derive_gzz = c_gxx >= 0 && c_gyy >= 0 && c_gzz >= 0;
    double amp = 1.0;
    if (job->physics == PF_MAGNETICS) {
        const double fx = job->b0[0], fy = job->b0[1], fz = job->b0[2];
        amp = sqrt(fx * fx + fy * fy + fz * fz);
    }
    int n = 0;
    const int families[2] = {12, 3};
    for (int family : families) {
        bool open = false;
        for (int c = 0; c < job->n_components; ++c) {
            CompInfo ci;
            component_info(job->physics, job->components[c], ci);
            if (ci.family != family) continue;
            if (derive_gzz && job->components[c] == PF_GZZ) continue;   // appended below, no plane
            if (!open) {
                if (n == max_passes) return -1;
                init_pass(plan, job, amp, passes[n], family);
                open = true;
            }
            if (!add_component(job, c, ci, amp, passes[n].p)) {
                ++n;
                if (n == max_passes) return -1;
                init_pass(plan, job, amp, passes[n], family);
                if (!add_component(job, c, ci, amp, passes[n].p)) return -1;
            }
        }
        if (open) ++n;
    }
    for (int i = 0; i < n; ++i) {
        LaunchParams &p = passes[i].p;
        {
            // canonical plane order: raw kernels first (increasing q), then the combinations
            PlaneDesc sorted[MAX_PLANES];
            int new_of_old[MAX_PLANES];
            int m = 0;
            p.raw_mask = 0;
            for (int q = 0; q < 10; ++q)
                for (int sl = 0; sl < p.n_planes; ++sl)
                    if (p.planes[sl].kind == PK_RAW && p.planes[sl].q[0] == q) {
                        new_of_old[sl] = m;
                        sorted[m++] = p.planes[sl];
                        p.raw_mask |= 1 << q;
                    }
            p.n_combo = 0;
            for (int sl = 0; sl < p.n_planes; ++sl)
                if (p.planes[sl].kind != PK_RAW) {
                    new_of_old[sl] = m;
                    sorted[m++] = p.planes[sl];
                    p.n_combo++;
                }
            for (int sl = 0; sl < p.n_planes; ++sl) p.planes[sl] = sorted[sl];
            for (int sl = 0; sl < p.n_planes; ++sl) {
                const PlaneDesc &pd = p.planes[sl];
                p.combo_mask[sl] = 0;
                for (int q = 0; q < 10; ++q) p.combo_coef[sl][q] = 0.0;
                const double f[3] = {p.fx, p.fy, p.fz};
                if (pd.kind == PK_DOT3) {
                    for (int a = 0; a < 3; ++a) p.combo_coef[sl][pd.q[a]] += f[a];
                } else if (pd.kind == PK_QUAD) {
                    // f^T K f with K = (xx, yy, zz, xy, xz, yz) = k[q0..q5]
                    p.combo_coef[sl][pd.q[0]] += f[0] * f[0];
                    p.combo_coef[sl][pd.q[1]] += f[1] * f[1];
                    p.combo_coef[sl][pd.q[2]] += f[2] * f[2];
                    p.combo_coef[sl][pd.q[3]] += 2.0 * f[0] * f[1];
                    p.combo_coef[sl][pd.q[4]] += 2.0 * f[0] * f[2];
                    p.combo_coef[sl][pd.q[5]] += 2.0 * f[1] * f[2];
                }
                if (pd.kind != PK_RAW)
                    for (int q = 0; q < 10; ++q)
                        if (p.combo_coef[sl][q] != 0.0) p.combo_mask[sl] |= 1 << q;
            }
            for (int o = 0; o < p.n_out; ++o) p.outs[o].plane = new_of_old[p.outs[o].plane];
            // one plane per output: order the outputs by plane (phase 2 then reads plane o for
            // output o with compile-time offsets) if every component's outputs stay contiguous
            if (p.n_out == p.n_planes) {
                OutDesc by_plane[MAX_OUT];
                bool distinct = true;
                bool seen[MAX_PLANES] = {false};
                for (int o = 0; o < p.n_out; ++o) {
                    if (seen[p.outs[o].plane]) distinct = false;
                    seen[p.outs[o].plane] = true;
                    by_plane[p.outs[o].plane] = p.outs[o];
                }
                bool contiguous = distinct;
                if (distinct) {
                    bool closed[PF_MAX_COMPONENTS] = {false};
                    for (int o = 0; o < p.n_out; ++o) {
                        const int fi = by_plane[o].fi;
                        if (closed[fi]) contiguous = false;
                        if (o + 1 == p.n_out || by_plane[o + 1].fi != fi) closed[fi] = true;
                    }
                }
                if (contiguous) {
                    for (int o = 0; o < p.n_out; ++o) {
                        p.outs[o] = by_plane[o];
                        p.outs[o].last = (o + 1 == p.n_out || by_plane[o + 1].fi != by_plane[o].fi);
                    }
                }
            }
        }
        p.direct = 1;
        for (int o = 0; o < p.n_out; ++o)
            if (p.outs[o].plane != o) p.direct = 0;
        p.n_derived = 0;
        if (derive_gzz && passes[i].family == 12) {
            // the pass that holds gxx also holds gyy (gravity never needs more than one family-12
            // pass: at most 9 planes); every gzz row of the job is derived from them
            int oa = -1, ob = -1;
            for (int o = 0; o < p.n_out; ++o) {
                if (p.outs[o].ci == c_gxx) oa = o;
                if (p.outs[o].ci == c_gyy) ob = o;
            }
            if (oa < 0 || ob < 0) return -1;
            p.outs[oa].role = ROLE_KEEP_A;
            p.outs[ob].role = ROLE_KEEP_B;
            for (int c = 0; c < job->n_components; ++c) {
                if (job->components[c] != PF_GZZ) continue;
                if (p.n_out == MAX_OUT || p.n_fwd == PF_MAX_COMPONENTS) return -1;
                OutDesc &od = p.outs[p.n_out++];
                od.plane = 0;
                od.ci = c;
                od.blk = 0;
                od.fi = p.n_fwd++;
                od.last = 1;
                od.role = ROLE_DERIVED;
                od.factor = p.outs[oa].factor;
                p.n_derived++;
            }
        }
        passes[i].prog = PROG_GENERIC;
        if (passes[i].family == 12 && job->physics == PF_MAGNETICS && p.n_fwd == 1 &&
            job->components[p.outs[0].ci] == PF_TMI)
        {
            // the TMI programs assume output o = plane o = column block o with one common factor
            bool regular = p.direct && p.n_out == p.nblk;
            for (int o = 0; o < p.n_out && regular; ++o)
                regular = p.outs[o].blk == o && p.outs[o].factor == p.outs[0].factor;
            if (regular) {
                passes[i].prog = (p.nblk == 3) ? PROG_TMI_VECTOR : PROG_TMI_SCALAR;
                // the common factor moves into the node evaluation (eval_node)
                const double factor = p.outs[0].factor;
                p.gx = factor * p.fx;
                p.gy = factor * p.fy;
                p.gz = factor * p.fz;
                for (int o = 0; o < p.n_out; ++o) p.outs[o].factor = 1.0;
                p.f2x = 2.0 * p.fx; p.f2y = 2.0 * p.fy; p.f2z = 2.0 * p.fz;
                p.g2x = 2.0 * p.gx; p.g2y = 2.0 * p.gy; p.g2z = 2.0 * p.gz;
            }
        }
        else if (passes[i].family == 12 &&
                 p.need == (NEED_AEE | NEED_ANN | NEED_AUU | NEED_LEN | NEED_LEU | NEED_LNU))
            passes[i].prog = PROG_ALL6;
        else if (passes[i].family == 12 && job->physics == PF_GRAVITY && p.n_out == 1 && p.n_planes == 1 &&
                 p.planes[0].kind == PK_RAW && p.planes[0].q[0] >= Q_E && p.planes[0].q[0] <= Q_U) {
            passes[i].prog = PROG_G1;
            p.g1_axis = p.planes[0].q[0] - Q_E;
        }
    }
    return n;
}

constexpr int MAX_PASSES = 2 * PF_MAX_COMPONENTS;

// receivers per chunk: enough CTAs to fill the machine a few times over
int choose_rc(int64_t n_tiles, int64_t n_receivers, int64_t max_chunks)
{
    const int64_t target_ctas = (int64_t)pf_sm_count() * 2 * 8;
    int64_t chunks = pf_ceil_div(target_ctas, n_tiles > 0 ? n_tiles : 1);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t rc = pf_ceil_div(n_receivers, chunks);
    if (rc < 1) rc = 1;
    if (rc > 65535) rc = 65535;
    return (int)rc;
}

template <int FAM, int PROG, int SINK, typename GT>
int launch_tile_kernel(const LaunchParams &p, cudaStream_t stream)
{
    const size_t smem = ((size_t)(3 + p.n_planes) * PF_TILE_NODES +
                         (SINK == SINK_FORWARD ? WARPS * p.n_fwd : 0)) * sizeof(double);
    auto kernel = pf_tile_kernel<FAM, PROG, SINK, GT>;
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
int launch_pass(const Pass &pass, cudaStream_t stream)
{
    if (pass.family == 12) {
        if (pass.prog == PROG_TMI_VECTOR)
            return launch_tile_kernel<12, PROG_TMI_VECTOR, SINK, GT>(pass.p, stream);
        if (pass.prog == PROG_TMI_SCALAR)
            return launch_tile_kernel<12, PROG_TMI_SCALAR, SINK, GT>(pass.p, stream);
        if (pass.prog == PROG_ALL6) return launch_tile_kernel<12, PROG_ALL6, SINK, GT>(pass.p, stream);
        if (pass.prog == PROG_G1) return launch_tile_kernel<12, PROG_G1, SINK, GT>(pass.p, stream);
        return launch_tile_kernel<12, PROG_GENERIC, SINK, GT>(pass.p, stream);
    }
    return launch_tile_kernel<3, PROG_GENERIC, SINK, GT>(pass.p, stream);
}

// ---- z-marching kernel dispatch (tensor meshes) --------------------------------
constexpr size_t MARCH_SMEM_LIMIT = 110 * 1024;  // keeps >= 2 CTAs per SM

size_t march_smem_n(const LaunchParams &p, int nr, bool forward, int n_buf, int mt = PF_THREADS)
{
    return (((size_t)n_buf * p.n_planes + (p.n_out - p.n_derived)) * nr * mt
            + (forward ? (mt / 32) * p.n_fwd : 0)) * sizeof(double);
}

// double-buffered node planes (one barrier per layer) unless that costs the third CTA per SM and
// a single buffer (two barriers per layer) keeps it
#ifndef PF_MARCH_SMEM_BUDGET
#define PF_MARCH_SMEM_BUDGET (74 * 1024)
#endif
constexpr size_t MARCH_SMEM_3CTAS = PF_MARCH_SMEM_BUDGET;
int march_buffers(const LaunchParams &p, int nr, bool forward, int mt = PF_THREADS)
{
    const size_t budget = MARCH_SMEM_3CTAS * mt / PF_THREADS;     // (6 CTAs per SM at 128 threads)
    if (march_smem_n(p, nr, forward, 2, mt) > budget && march_smem_n(p, nr, forward, 1, mt) <= budget)
        return 1;
    return 2;
}

size_t march_smem(const LaunchParams &p, int nr, bool forward, int mt = PF_THREADS)
{
    return march_smem_n(p, nr, forward, march_buffers(p, nr, forward, mt), mt);
}

// node rounds per layer (4 -> 1024-node footprints, 2 -> 512) or 0 when the pass
// cannot use the marching kernel
int march_rounds(const pf_plan *plan, const LaunchParams &p, bool forward)
{
    if (plan->grid_nx <= 0) return 0;
#ifndef PF_PREFER_NR2
    if (plan->n_footprints4 > 0 && p.n_planes <= 3 && march_smem(p, 4, forward) <= MARCH_SMEM_LIMIT) return 4;
#endif
    if (plan->n_footprints2 > 0 && march_smem(p, 2, forward) <= MARCH_SMEM_LIMIT) return 2;
    return 0;
}

// Threads per CTA of a marching pass.  The three-plane vector program runs in 128-thread CTAs (six
// per SM instead of three, 512-node footprints `footprints4h`) when the plan carries that set: a CTA
// in its cell phase then takes a sixth instead of a third of the partition's warps off the FP64 pipe
// (measured: build +1.5 %, blocked diagonal +7 %; the one-plane programs lose 1.5 % and keep 256).
#ifndef PF_MARCH_HALF_CTAS
#define PF_MARCH_HALF_CTAS 1
#endif
constexpr int MARCH_HALF_THREADS = 128;
int march_threads(const pf_plan *plan, const Pass &pass, int nr)
{
    if (PF_MARCH_HALF_CTAS && pass.family == 12 && pass.prog == PROG_TMI_VECTOR && nr == 4 &&
        plan->n_footprints4h > 0 && plan->footprints4h)
        return MARCH_HALF_THREADS;
    return PF_THREADS;
}
// footprints (= grid.x and forward partials) of a marching pass
int march_parts(const pf_plan *plan, const Pass &pass, int nr)
{
    if (march_threads(plan, pass, nr) != PF_THREADS) return plan->n_footprints4h;
    return (nr == 4) ? plan->n_footprints4 : plan->n_footprints2;
}

template <int FAM, int PROG, int SINK, typename GT, int NR, int MT = PF_THREADS>
int launch_march_kernel(const LaunchParams &p, cudaStream_t stream)
{
    const size_t smem = march_smem(p, NR, SINK == SINK_FORWARD, MT);
    auto kernel = pf_march_kernel<FAM, PROG, SINK, GT, NR, MT>;
    LaunchParams q = p;
    q.n_buf = march_buffers(p, NR, SINK == SINK_FORWARD, MT);
    PF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t chunks = pf_ceil_div(p.n_receivers, p.rc);
    PF_CHECK_ARG(chunks <= 65535, "too many receiver chunks");
    const int n_fp = (MT != PF_THREADS) ? p.plan.n_footprints4h
                                        : ((NR == 4) ? p.plan.n_footprints4 : p.plan.n_footprints2);
    dim3 grid((unsigned)n_fp, (unsigned)chunks);
    kernel<<<grid, MT, smem, stream>>>(q);
    PF_COUNT_LAUNCH();
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

template <int SINK, typename GT, int NR>
int launch_march_nr(const Pass &pass, cudaStream_t stream)
{
    if (pass.family == 12) {
        if (pass.prog == PROG_TMI_VECTOR) {
            if constexpr (NR == 4) {
                if (march_threads(&pass.p.plan, pass, NR) == MARCH_HALF_THREADS)
                    return launch_march_kernel<12, PROG_TMI_VECTOR, SINK, GT, NR, MARCH_HALF_THREADS>(pass.p, stream);
            }
            return launch_march_kernel<12, PROG_TMI_VECTOR, SINK, GT, NR>(pass.p, stream);
        }
        if (pass.prog == PROG_TMI_SCALAR)
            return launch_march_kernel<12, PROG_TMI_SCALAR, SINK, GT, NR>(pass.p, stream);
        if (pass.prog == PROG_ALL6)
            return launch_march_kernel<12, PROG_ALL6, SINK, GT, NR>(pass.p, stream);
        if (pass.prog == PROG_G1)
            return launch_march_kernel<12, PROG_G1, SINK, GT, NR>(pass.p, stream);
        return launch_march_kernel<12, PROG_GENERIC, SINK, GT, NR>(pass.p, stream);
    }
    return launch_march_kernel<3, PROG_GENERIC, SINK, GT, NR>(pass.p, stream);
}

template <int SINK, typename GT>
int launch_march(const Pass &pass, int nr, cudaStream_t stream)
{
    return (nr == 4) ? launch_march_nr<SINK, GT, 4>(pass, stream) : launch_march_nr<SINK, GT, 2>(pass, stream);
}

// Passes of the sinks that can run on the marching kernel (store, forward): derived outputs are
// used when every pass of the job then runs on it, otherwise the plain schedule.
int make_passes_march(const pf_plan *plan, const pf_job *job, Pass *passes, bool forward)
{
    if (plan->grid_nx > 0) {
        const int n = make_passes(plan, job, passes, MAX_PASSES, true);
        bool derived = false, ok = n >= 0;
        for (int i = 0; ok && i < n; ++i) derived = derived || passes[i].p.n_derived > 0;
        if (ok && !derived) return n;
        for (int i = 0; ok && i < n; ++i) ok = march_rounds(plan, passes[i].p, forward) != 0;
        if (ok) return n;
    }
    return make_passes(plan, job, passes, MAX_PASSES, false);
}

int validate_grid(const pf_plan *plan)
{
    if (plan->grid_nx == 0) return PF_OK;
    PF_CHECK_ARG(plan->grid_nx > 0 && plan->grid_ny > 0 && plan->grid_nz > 0, "bad grid shape");
    PF_CHECK_ARG((int64_t)plan->grid_nx * plan->grid_ny * plan->grid_nz < ((int64_t)1 << 31),
                 "tensor mesh with 2^31 or more cells");
    PF_CHECK_ARG(plan->grid_xs && plan->grid_ys && plan->grid_zs, "null grid axis");
    PF_CHECK_ARG(plan->n_footprints4 >= 0 && plan->n_footprints2 >= 0, "negative footprint count");
    PF_CHECK_ARG(plan->n_footprints4 == 0 || plan->footprints4, "null footprints4");
    PF_CHECK_ARG(plan->n_footprints2 == 0 || plan->footprints2, "null footprints2");
    PF_CHECK_ARG(plan->n_footprints4h >= 0 && (plan->n_footprints4h == 0 || plan->footprints4h), "bad footprints4h");
    PF_CHECK_ARG(plan->grid_colmap || (int64_t)plan->grid_nx * plan->grid_ny * plan->grid_nz == plan->n_cells,
                 "grid without colmap must have n_cells == nx ny nz");
    return PF_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int64_t pf_job_columns(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    return plan->n_cells * n_blocks_of(job);
}

extern "C" int pf_job_uses_marching(const pf_plan *plan, const pf_job *job, int forward)
{
    if (validate(plan, job) != PF_OK) return 0;
    if (plan->grid_nx <= 0) return 0;
    static thread_local Pass passes[MAX_PASSES];
    const int n = make_passes_march(plan, job, passes, forward != 0);
    if (n <= 0) return 0;
    for (int i = 0; i < n; ++i)
        if (!march_rounds(plan, passes[i].p, forward != 0)) return 0;
    return 1;
}

extern "C" int pf_build_G(const pf_plan *plan, const pf_job *job, void *G, int64_t ldG,
                          int g_dtype, void *stream)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    PF_CHECK_ARG(g_dtype == PF_F64 || g_dtype == PF_F32, "bad g_dtype");
    if (plan->n_cells == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(G != nullptr, "null G");
    PF_CHECK_ARG(ldG >= pf_job_columns(plan, job), "ldG smaller than the number of columns");
    st = validate_grid(plan);
    if (st != PF_OK) return st;
    static thread_local Pass passes[MAX_PASSES];
    const int n = make_passes_march(plan, job, passes, false);
    PF_CHECK_ARG(n >= 0, "component list cannot be scheduled");
    for (int i = 0; i < n; ++i) {
        LaunchParams &p = passes[i].p;
        p.G = G;
        p.ldG = ldG;
        const int nr = march_rounds(plan, p, false);
        if (nr) {
            const int n_fp = march_parts(plan, passes[i], nr);
            p.rc = choose_rc(n_fp, job->n_receivers, 65535);
            st = (g_dtype == PF_F64) ? launch_march<SINK_STORE, double>(passes[i], nr, (cudaStream_t)stream)
                                     : launch_march<SINK_STORE, float>(passes[i], nr, (cudaStream_t)stream);
        } else {
            PF_CHECK_ARG(plan->n_tiles > 0, "tile arrays are required for this component set (no marching path)");
            p.rc = choose_rc(plan->n_tiles, job->n_receivers, 65535);
            st = (g_dtype == PF_F64) ? launch_pass<SINK_STORE, double>(passes[i], (cudaStream_t)stream)
                                     : launch_pass<SINK_STORE, float>(passes[i], (cudaStream_t)stream);
        }
        if (st != PF_OK) return st;
    }
    return PF_OK;
}

namespace {
// forward partial sums: [parts][receivers of the chunk][components] doubles.  For tree meshes and
// layers `parts` is the number of tiles (~ nC / 512), so the receivers of a job are processed in
// chunks whose partials stay below FWD_WS_CAP (>= FWD_MIN_RX receivers per chunk)
constexpr int64_t FWD_WS_CAP = (int64_t)256 << 20;
constexpr int64_t FWD_MIN_RX = 64;

int64_t forward_parts(const pf_plan *plan)
{
    int64_t parts = plan->n_tiles;
    if (plan->n_footprints4 > parts) parts = plan->n_footprints4;
    if (plan->n_footprints2 > parts) parts = plan->n_footprints2;
    if (plan->n_footprints4h > parts) parts = plan->n_footprints4h;
    return parts > 0 ? parts : 1;
}

int64_t forward_bytes_per_receiver(const pf_plan *plan, const pf_job *job)
{
    return forward_parts(plan) * (int64_t)job->n_components * (int64_t)sizeof(double);
}
}  // namespace

extern "C" int64_t pf_forward_workspace_bytes(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    const int64_t per_rx = forward_bytes_per_receiver(plan, job);
    int64_t rx = FWD_WS_CAP / per_rx;
    if (rx < FWD_MIN_RX) rx = FWD_MIN_RX;
    if (rx > job->n_receivers) rx = job->n_receivers;
    return rx * per_rx;
}

static int pf_forward_chunk(const pf_plan *plan, const pf_job *job, const double *model, double *fields,
                            void *workspace, void *stream);

extern "C" int pf_forward(const pf_plan *plan, const pf_job *job, const double *model,
                          double *fields, void *workspace, int64_t workspace_bytes, void *stream)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    if (plan->n_cells == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(model && fields, "null model or fields");
    st = validate_grid(plan);
    if (st != PF_OK) return st;
    const int64_t per_rx = forward_bytes_per_receiver(plan, job);
    const int64_t rx_chunk = workspace ? workspace_bytes / per_rx : 0;
    if (rx_chunk < 1 || workspace_bytes < pf_forward_workspace_bytes(plan, job)) {
        pf_set_error("pf_forward: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    for (int64_t r0 = 0; r0 < job->n_receivers; r0 += rx_chunk) {
        pf_job sub = *job;
        sub.receivers = job->receivers + 3 * r0;
        sub.n_receivers = (job->n_receivers - r0 < rx_chunk) ? job->n_receivers - r0 : rx_chunk;
        sub.row_offset = job->row_offset + r0 * job->n_components;
        st = pf_forward_chunk(plan, &sub, model, fields, workspace, stream);
        if (st != PF_OK) return st;
    }
    return PF_OK;
}

static int pf_forward_chunk(const pf_plan *plan, const pf_job *job, const double *model, double *fields,
                            void *workspace, void *stream)
{
    int st = PF_OK;
    static thread_local Pass passes[MAX_PASSES];
    const int n = make_passes_march(plan, job, passes, true);
    PF_CHECK_ARG(n >= 0, "component list cannot be scheduled");
    for (int i = 0; i < n; ++i) {
        LaunchParams &p = passes[i].p;
        p.model = model;
        p.partial = (double *)workspace;
        const int nr = march_rounds(plan, p, true);
        int64_t n_parts = plan->n_tiles;
        if (nr) {
            n_parts = march_parts(plan, passes[i], nr);
            p.rc = choose_rc(n_parts, job->n_receivers, 65535);
            st = launch_march<SINK_FORWARD, double>(passes[i], nr, (cudaStream_t)stream);
        } else {
            PF_CHECK_ARG(plan->n_tiles > 0, "tile arrays are required for this component set (no marching path)");
            p.rc = choose_rc(plan->n_tiles, job->n_receivers, 65535);
            st = launch_pass<SINK_FORWARD, double>(passes[i], (cudaStream_t)stream);
        }
        if (st != PF_OK) return st;
        ReduceFwdParams q;
        q.partial = p.partial;
        q.n_tiles = n_parts;
        q.n_receivers = job->n_receivers;
        q.n_fwd = p.n_fwd;
        q.n_comp_total = job->n_components;
        q.row_offset = job->row_offset;
        for (int o = 0; o < p.n_out; ++o) q.ci[p.outs[o].fi] = p.outs[o].ci;
        q.fields = fields;
        const int64_t n_local = job->n_receivers * p.n_fwd;
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
    if (plan->n_cells == 0 || job->n_receivers == 0) return PF_OK;
    PF_CHECK_ARG(vec && out, "null vector or output");
    const int64_t nP = pf_job_columns(plan, job);
    const int64_t chunks = column_chunks(plan, job);
    const bool node_space = (SINK == SINK_GTV && plan->n_nodes > 0);
    if (!workspace || (!node_space && workspace_bytes < chunks * nP * (int64_t)sizeof(double))) {
        pf_set_error("column sink: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    static thread_local Pass passes[MAX_PASSES];
    const int n = make_passes(plan, job, passes, MAX_PASSES);
    PF_CHECK_ARG(n >= 0, "component list cannot be scheduled");
    if (SINK == SINK_GTV && plan->n_nodes > 0) {
        PF_CHECK_ARG(plan->node_xyz && plan->cell_nodes32, "null node table");
        const int nblk = n_blocks_of(job);
        const int64_t nN = plan->n_nodes;
        const int64_t node_ctas = pf_ceil_div(nN, PF_THREADS);
        const int rc = choose_rc(node_ctas, job->n_receivers, MAX_COLUMN_CHUNKS);
        const int64_t nchunks = pf_ceil_div(job->n_receivers, rc);
        if (workspace_bytes < (nchunks + 1) * nblk * nN * (int64_t)sizeof(double)) {
            pf_set_error("pf_gt_dot_v: workspace too small");
            return PF_ERR_WORKSPACE;
        }
        double *partial = (double *)workspace;
        double *N = partial + nchunks * nblk * nN;
        for (int i = 0; i < n; ++i) {
            LaunchParams &p = passes[i].p;
            p.vec = vec;
            p.rc = rc;
            const size_t smem = (size_t)p.n_planes * PF_THREADS * sizeof(double);
            dim3 grid((unsigned)node_ctas, (unsigned)nchunks);
#define PF_LAUNCH_NODE(F, P)                                                                          \
            pf_node_gtv_kernel<F, P><<<grid, PF_THREADS, smem, (cudaStream_t)stream>>>(p, partial)
            if (passes[i].family == 3) PF_LAUNCH_NODE(3, PROG_GENERIC);
            else if (passes[i].prog == PROG_TMI_VECTOR) PF_LAUNCH_NODE(12, PROG_TMI_VECTOR);
            else if (passes[i].prog == PROG_TMI_SCALAR) PF_LAUNCH_NODE(12, PROG_TMI_SCALAR);
            else if (passes[i].prog == PROG_ALL6) PF_LAUNCH_NODE(12, PROG_ALL6);
            else if (passes[i].prog == PROG_G1) PF_LAUNCH_NODE(12, PROG_G1);
            else PF_LAUNCH_NODE(12, PROG_GENERIC);
#undef PF_LAUNCH_NODE
            PF_COUNT_LAUNCH();
            PF_CUDA_OK(cudaGetLastError());
            const double *src = partial;
            if (nchunks > 1) {
                pf_node_reduce_kernel<<<(unsigned)pf_ceil_div(nblk * nN, 256), 256, 0, (cudaStream_t)stream>>>(
                    partial, nchunks, nblk * nN, N);
                PF_COUNT_LAUNCH();
                PF_CUDA_OK(cudaGetLastError());
                src = N;
            }
            pf_node_to_cells_kernel<<<(unsigned)pf_ceil_div(plan->n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
                src, nN, plan->cell_nodes32, plan->n_cells, nblk, out);
            PF_COUNT_LAUNCH();
            PF_CUDA_OK(cudaGetLastError());
        }
        return PF_OK;
    }
    PF_CHECK_ARG(plan->n_tiles > 0, "tile arrays are required for this sink");
    for (int i = 0; i < n; ++i) {
        LaunchParams &p = passes[i].p;
        p.vec = vec;
        p.partial = (double *)workspace;
        p.rc = choose_rc(plan->n_tiles, job->n_receivers, MAX_COLUMN_CHUNKS);
        st = launch_pass<SINK, double>(passes[i], (cudaStream_t)stream);
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
    int64_t bytes = column_chunks(plan, job) * pf_job_columns(plan, job) * (int64_t)sizeof(double);
    if (plan->n_nodes > 0) {
        const int rc = choose_rc(pf_ceil_div(plan->n_nodes, PF_THREADS), job->n_receivers, MAX_COLUMN_CHUNKS);
        const int64_t nchunks = pf_ceil_div(job->n_receivers > 0 ? job->n_receivers : 1, rc);
        const int64_t node_bytes = (nchunks + 1) * n_blocks_of(job) * plan->n_nodes * (int64_t)sizeof(double);
        if (node_bytes > bytes) bytes = node_bytes;
    }
    return bytes;
}

extern "C" int pf_gt_dot_v(const pf_plan *plan, const pf_job *job, const double *v, double *result,
                           void *workspace, int64_t workspace_bytes, void *stream)
{
    return column_sink<SINK_GTV>(plan, job, v, result, workspace, workspace_bytes, stream);
}

// ---- on-the-fly diag(G^T W G) on tensor meshes: blocked build + column norms -----------------
// The tile kernel needs 1.47 node evaluations per pair (bricks) against 1.10 of the z-marching
// kernel, but a marching kernel cannot keep per-cell accumulators (it streams the cells).  So the
// diagonal is formed the way the stored-G path forms it, a block of receivers at a time: the
// marching kernel writes the block's rows into scratch memory and the HBM-bound column kernel folds
// sum_i w_i g_ij^2 into the result (24 B written + read per pair at ~7 TB/s against ~400 FP64
// flop-eq per pair).  Identical arithmetic to pf_build_G + pf_colsumsq.
namespace {
constexpr int64_t DIAG_SCRATCH_TARGET = (int64_t)4 << 30;   // recommended scratch (callers may pass more)
constexpr int64_t DIAG_MIN_RX = 16;

bool diag_blocked(const pf_plan *plan, const pf_job *job)
{
    return validate(plan, job) == PF_OK && plan->n_cells > 0 && job->n_receivers > 0 &&
           pf_job_uses_marching(plan, job, 0) != 0;
}
int64_t diag_ld(const pf_plan *plan, const pf_job *job) { return pf_ceil_div(pf_job_columns(plan, job), 16) * 16; }
}  // namespace

extern "C" int64_t pf_gemv_t_workspace_bytes(int64_t rows, int64_t cols);

namespace {
// workspace of the column kernel for a block of rb receivers (non-decreasing in rb)
int64_t diag_column_ws(const pf_plan *plan, const pf_job *job, int64_t rb)
{
    return pf_gemv_t_workspace_bytes(rb * job->n_components, pf_job_columns(plan, job));
}
}  // namespace

extern "C" int pf_colsumsq_acc(const void *G, int64_t rows, int64_t cols, int64_t ld, int g_dtype,
                               const double *w, double *d, void *workspace, int64_t workspace_bytes,
                               void *stream);

extern "C" int64_t pf_gtg_diag_workspace_bytes(const pf_plan *plan, const pf_job *job)
{
    if (!plan || !job) return -1;
    if (diag_blocked(plan, job)) {
        const int64_t per_rx = (int64_t)job->n_components * diag_ld(plan, job) * (int64_t)sizeof(double);
        int64_t rb = DIAG_SCRATCH_TARGET / per_rx;
        if (rb < DIAG_MIN_RX) rb = DIAG_MIN_RX;
        if (rb > job->n_receivers) rb = job->n_receivers;
        return rb * per_rx + diag_column_ws(plan, job, rb);
    }
    return column_chunks(plan, job) * pf_job_columns(plan, job) * (int64_t)sizeof(double);
}

extern "C" int pf_gtg_diag(const pf_plan *plan, const pf_job *job, const double *w, double *diag,
                           void *workspace, int64_t workspace_bytes, void *stream)
{
    if (!diag_blocked(plan, job))
        return column_sink<SINK_DIAG>(plan, job, w, diag, workspace, workspace_bytes, stream);
    PF_CHECK_ARG(w && diag, "null vector or output");
    const int64_t nP = pf_job_columns(plan, job), ld = diag_ld(plan, job);
    const int64_t per_rx = (int64_t)job->n_components * ld * (int64_t)sizeof(double);
    int64_t rb = workspace ? workspace_bytes / per_rx : 0;
    if (rb > job->n_receivers) rb = job->n_receivers;
    for (int it = 0; it < 4 && rb >= 1; ++it) {     // the column workspace shrinks with the block
        const int64_t fit = (workspace_bytes - diag_column_ws(plan, job, rb)) / per_rx;
        if (fit >= rb) break;
        rb = fit;
    }
    const int64_t ws2 = rb >= 1 ? diag_column_ws(plan, job, rb) : 0;
    if (rb < 1) {
        pf_set_error("pf_gtg_diag: workspace too small");
        return PF_ERR_WORKSPACE;
    }
    char *scratch = (char *)workspace;
    void *colws = scratch + rb * per_rx;
    for (int64_t r0 = 0; r0 < job->n_receivers; r0 += rb) {
        pf_job sub = *job;
        sub.receivers = job->receivers + 3 * r0;
        sub.n_receivers = (job->n_receivers - r0 < rb) ? job->n_receivers - r0 : rb;
        sub.row_offset = 0;
        int st = pf_build_G(plan, &sub, scratch, ld, PF_F64, stream);
        if (st != PF_OK) return st;
        st = pf_colsumsq_acc(scratch, sub.n_receivers * job->n_components, nP, ld, PF_F64,
                             w + job->row_offset + r0 * job->n_components, diag, colws, ws2, stream);
        if (st != PF_OK) return st;
    }
    return PF_OK;
}

namespace {
const char *prog_name(int prog)
{
    switch (prog) {
    case PROG_TMI_SCALAR: return "PROG_TMI_SCALAR";
    case PROG_TMI_VECTOR: return "PROG_TMI_VECTOR";
    case PROG_ALL6: return "PROG_ALL6";
    case PROG_G1: return "PROG_G1";
    default: return "PROG_GENERIC";
    }
}
const char *sink_name(int sink)
{
    switch (sink) {
    case SINK_STORE: return "SINK_STORE";
    case SINK_FORWARD: return "SINK_FORWARD";
    case SINK_GTV: return "SINK_GTV";
    default: return "SINK_DIAG";
    }
}
}  // namespace

// Names of the kernels (template instantiations) the job's passes launch for a sink, joined by
// " + " -- what bench.py prints as roofline.kernel, so that the label follows the dispatch.
extern "C" int pf_describe(const pf_plan *plan, const pf_job *job, int sink, int g_dtype, char *buf,
                           int64_t len)
{
    int st = validate(plan, job);
    if (st != PF_OK) return st;
    PF_CHECK_ARG(buf && len > 0, "null buffer");
    PF_CHECK_ARG(sink >= SINK_STORE && sink <= SINK_DIAG, "bad sink");
    static thread_local Pass passes[MAX_PASSES];
    const bool blocked_diag = sink == SINK_DIAG && diag_blocked(plan, job);
    const int n = (sink == SINK_STORE || sink == SINK_FORWARD || blocked_diag)
                      ? make_passes_march(plan, job, passes, sink == SINK_FORWARD)
                      : make_passes(plan, job, passes, MAX_PASSES);
    PF_CHECK_ARG(n >= 0, "component list cannot be scheduled");
    const char *gt = (sink == SINK_STORE && g_dtype == PF_F32) ? "float" : "double";
    int64_t used = 0;
    buf[0] = 0;
    for (int i = 0; i < n; ++i) {
        const LaunchParams &p = passes[i].p;
        char one[160];
        int nr = 0;
        if (sink == SINK_STORE || sink == SINK_FORWARD) nr = march_rounds(plan, p, sink == SINK_FORWARD);
        if (nr && march_threads(plan, passes[i], nr) != PF_THREADS)
            snprintf(one, sizeof(one), "pf_march_kernel<%d,%s,%s,%s,%d,%d>", passes[i].family,
                     prog_name(passes[i].prog), sink_name(sink), gt, nr, march_threads(plan, passes[i], nr));
        else if (nr)
            snprintf(one, sizeof(one), "pf_march_kernel<%d,%s,%s,%s,%d>", passes[i].family,
                     prog_name(passes[i].prog), sink_name(sink), gt, nr);
        else if (sink == SINK_GTV && plan->n_nodes > 0)
            snprintf(one, sizeof(one), "pf_node_gtv_kernel<%d,%s>", passes[i].family, prog_name(passes[i].prog));
        else if (sink == SINK_DIAG && blocked_diag) {
            const int nrd = march_rounds(plan, p, false), mtd = march_threads(plan, passes[i], nrd);
            if (mtd != PF_THREADS)
                snprintf(one, sizeof(one), "pf_march_kernel<%d,%s,SINK_STORE,double,%d,%d> + column_kernel<double,1>",
                         passes[i].family, prog_name(passes[i].prog), nrd, mtd);
            else
                snprintf(one, sizeof(one), "pf_march_kernel<%d,%s,SINK_STORE,double,%d> + column_kernel<double,1>",
                         passes[i].family, prog_name(passes[i].prog), nrd);
        }
        else
            snprintf(one, sizeof(one), "pf_tile_kernel<%d,%s,%s,%s>", passes[i].family,
                     prog_name(passes[i].prog), sink_name(sink), gt);
        const int64_t need = (int64_t)strlen(one) + (i ? 3 : 0);
        if (used + need + 1 > len) break;
        if (i) { memcpy(buf + used, " + ", 3); used += 3; }
        memcpy(buf + used, one, strlen(one));
        used += (int64_t)strlen(one);
        buf[used] = 0;
    }
    return PF_OK;
}

