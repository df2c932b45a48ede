/*
 * pf_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C/OpenMP restatement of the reference's numba loops for the
 * integral-equation potential-field hot path, plus a restatement of the
 * closed-form prism kernels of the third-party package `choclo` (pinned
 * `choclo>=0.3.0` in the reference's pyproject.toml:54; NOT vendored in the
 * reference tree, so its arithmetic is restated here from its published
 * formulas -- see DESIGN.md "Oracle").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (simpeg_b200/)
 * never does: it fails loudly when its CUDA library is missing.
 *
 * Parity pinning: this file is checked (tests/test_oracle_golden.py,
 * and the committed tests/golden npz files made by oracle/make_golden.py) against
 * the reference's own UNMODIFIED numba files
 *     simpeg/potential_fields/_numba_utils.py
 *     simpeg/potential_fields/gravity/_numba/_3d_mesh.py
 *     simpeg/potential_fields/magnetics/_numba/_3d_mesh.py
 * loaded by path (oracle/ref_loader.py) on top of a numba choclo shim
 * (oracle/choclo_shim).  The choclo formulas themselves are pinned against
 * Gauss-Legendre quadrature of the defining integrals and finite differences
 * (tests/test_choclo_kernels.py).  The singular branches of safe_log /
 * safe_atan2 are restated from memory of choclo/_utils.py: parity UNPINNED
 * for those branches against real choclo (absent from this image).
 *
 * Every loop mirrors the operation order of the reference (no FMA
 * contraction: build with -ffp-contract=off, no -ffast-math), because numba
 * compiles these loops with fastmath=False.
 *
 * Build: see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* choclo restatement (choclo/_utils.py, choclo/prism/_kernels.py)     */
/* ------------------------------------------------------------------ */

/* choclo._utils._safe_atan2: arctan(y/x) on (-pi/2, pi/2], +-pi/2 for x==0 */
static inline double safe_atan2(double y, double x)
{
    if (x != 0.0) return atan(y / x);
    if (y > 0.0) return M_PI / 2;
    if (y < 0.0) return -M_PI / 2;
    return 0.0;
}

/* choclo._utils._safe_log(x, y, z, r) = log(x + r) with a stable form for x<0 */
static inline double safe_log(double x, double y, double z, double r)
{
    if (r == 0.0) return 0.0;
    if (x < 0.0) {
        if (r == -x) return -log(-2 * x);
        return log((y * y + z * z) / (r - x));
    }
    return log(x + r);
}

/* Kernel ids: shared with the product's component tables (values are part
 * of the oracle's own ABI only). */
enum {
    K_E = 0, K_N, K_U,
    K_EE, K_NN, K_UU, K_EN, K_EU, K_NU,
    K_UV,
    K_EEE, K_NNN, K_UUU, K_EEN, K_EEU, K_ENN, K_NNU, K_EUU, K_NUU, K_ENU,
    K_COUNT
};

static inline double kernel_iii(double xi, double xj, double xk, double r)
{
    if ((xi == 0.0 && xj == 0.0) || (xi == 0.0 && xk == 0.0)) return 0.0;
    double xi2 = xi * xi, xj2 = xj * xj, xk2 = xk * xk;
    double numerator = -xj * xk * (2 * xi2 + xj2 + xk2);
    double denominator = (xi2 + xj2) * (xi2 + xk2) * r;
    return numerator / denominator;
}

static inline double kernel_iij(double xi, double xj, double xk, double r)
{
    if (xi == 0.0 && xj == 0.0 && xk < 0.0) return 0.0;
    return -xi / ((xk + r) * r);
}

static inline double eval_kernel(int id, double e, double n, double u, double r)
{
    switch (id) {
    /* first order: choclo.prism.kernel_e / kernel_n / kernel_u */
    case K_E:
        return -(n * safe_log(u, e, n, r) + u * safe_log(n, e, u, r)
                 - e * safe_atan2(n * u, e * r));
    case K_N:
        return -(u * safe_log(e, n, u, r) + e * safe_log(u, e, n, r)
                 - n * safe_atan2(u * e, n * r));
    case K_U:
        return -(e * safe_log(n, e, u, r) + n * safe_log(e, n, u, r)
                 - u * safe_atan2(e * n, u * r));
    /* second order */
    case K_EE: return -safe_atan2(n * u, e * r);
    case K_NN: return -safe_atan2(e * u, n * r);
    case K_UU: return -safe_atan2(e * n, u * r);
    case K_EN: return safe_log(u, e, n, r);
    case K_EU: return safe_log(n, e, u, r);
    case K_NU: return safe_log(e, n, u, r);
    /* simpeg gravity/simulation.py:23-30 kernel_uv */
    case K_UV:
        return 0.5 * (-safe_atan2(e * u, n * r) - (-safe_atan2(n * u, e * r)));
    /* third order */
    case K_EEE: return kernel_iii(e, n, u, r);
    case K_NNN: return kernel_iii(n, u, e, r);
    case K_UUU: return kernel_iii(u, e, n, r);
    case K_EEN: return kernel_iij(e, n, u, r);
    case K_EEU: return kernel_iij(e, u, n, r);
    case K_ENN: return kernel_iij(n, e, u, r);
    case K_NNU: return kernel_iij(n, u, e, r);
    case K_EUU: return kernel_iij(u, e, n, r);
    case K_NUU: return kernel_iij(u, n, e, r);
    case K_ENU: return (r == 0.0) ? 0.0 : -1.0 / r;
    default: return NAN;
    }
}

/* exported for unit tests of the formulas */
double pfo_kernel(int id, double e, double n, double u, double r)
{
    return eval_kernel(id, e, n, u, r);
}
double pfo_safe_log(double x, double y, double z, double r) { return safe_log(x, y, z, r); }
double pfo_safe_atan2(double y, double x) { return safe_atan2(y, x); }

/* simpeg/potential_fields/_numba_utils.py:18-45 kernels_in_nodes_to_cell */
static inline double nodes_to_cell(const double *k, const int64_t *idx)
{
    double result = (-k[idx[0]] + k[idx[1]] + k[idx[2]] - k[idx[3]] + k[idx[4]]
                     - k[idx[5]] - k[idx[6]] + k[idx[7]]);
    return result;
}

int pfo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void pfo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* gravity/_numba/_3d_mesh.py:472-498 _evaluate_kernel, looped over nodes */
static void eval_nodes_1(const double *rx, const double *nodes, int64_t nN,
                         int kid, double *k)
{
    for (int64_t j = 0; j < nN; ++j) {
        double dx = nodes[3 * j + 0] - rx[0];
        double dy = nodes[3 * j + 1] - rx[1];
        double dz = nodes[3 * j + 2] - rx[2];
        double distance = sqrt(dx * dx + dy * dy + dz * dz);
        k[j] = eval_kernel(kid, dx, dy, dz, distance);
    }
}

static void eval_nodes_n(const double *rx, const double *nodes, int64_t nN,
                         const int *kids, int nk, double *k /* nk planes of nN */)
{
    for (int64_t j = 0; j < nN; ++j) {
        double dx = nodes[3 * j + 0] - rx[0];
        double dy = nodes[3 * j + 1] - rx[1];
        double dz = nodes[3 * j + 2] - rx[2];
        double distance = sqrt(dx * dx + dy * dy + dz * dz);
        for (int q = 0; q < nk; ++q)
            k[(int64_t)q * nN + j] = eval_kernel(kids[q], dx, dy, dz, distance);
    }
}

/* ------------------------------------------------------------------ */
/* Gravity (simpeg/potential_fields/gravity/_numba/_3d_mesh.py)        */
/* ------------------------------------------------------------------ */

/* _sensitivity_gravity, lines 97-163.  G is a strided view: row i starts at
 * G + i*row_stride (elements); f32 != 0 stores float32 (sensitivity_dtype). */
void pfo_gravity_sensitivity(const double *receivers, int64_t R,
                             const double *nodes, int64_t nN,
                             const int64_t *cell_nodes, int64_t nC,
                             void *G, int64_t row_stride, int f32,
                             int kid, double constant_factor, int parallel)
{
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_1(receivers + 3 * i, nodes, nN, kid, k);
            if (f32) {
                float *row = (float *)G + i * row_stride;
                for (int64_t c = 0; c < nC; ++c)
                    row[c] = (float)(constant_factor * nodes_to_cell(k, cell_nodes + 8 * c));
            } else {
                double *row = (double *)G + i * row_stride;
                for (int64_t c = 0; c < nC; ++c)
                    row[c] = constant_factor * nodes_to_cell(k, cell_nodes + 8 * c);
            }
        }
        free(k);
    }
}

/* _forward_gravity, lines 21-94: fields[i] += c * rho[k] * vertexsum */
void pfo_gravity_forward(const double *receivers, int64_t R,
                         const double *nodes, int64_t nN,
                         const int64_t *cell_nodes, int64_t nC,
                         const double *densities, double *fields, int64_t fstride,
                         int kid, double constant_factor, int parallel)
{
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_1(receivers + 3 * i, nodes, nN, kid, k);
            double f = fields[i * fstride];
            for (int64_t c = 0; c < nC; ++c)
                f += constant_factor * densities[c] * nodes_to_cell(k, cell_nodes + 8 * c);
            fields[i * fstride] = f;
        }
        free(k);
    }
}

/* helper: numba's array reduction -- every thread owns a zero-initialised
 * private copy, the copies are summed into the shared array at the end. */
static void reduce_private(double *shared, double *priv, int64_t n)
{
#pragma omp critical
    {
        for (int64_t c = 0; c < n; ++c) shared[c] += priv[c];
    }
}

/* _sensitivity_gravity_t_dot_v_{serial,parallel}, lines 320-469 */
void pfo_gravity_gt_dot_v(const double *receivers, int64_t R,
                          const double *nodes, int64_t nN,
                          const int64_t *cell_nodes, int64_t nC,
                          const double *vector, int64_t vstride, double *result,
                          int kid, double constant_factor, int parallel)
{
    if (!parallel) {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN);
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_1(receivers + 3 * i, nodes, nN, kid, k);
            for (int64_t c = 0; c < nC; ++c)
                result[c] += constant_factor * vector[i * vstride]
                             * nodes_to_cell(k, cell_nodes + 8 * c);
        }
        free(k);
        return;
    }
#pragma omp parallel
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN);
        double *priv = (double *)calloc((size_t)nC, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_1(receivers + 3 * i, nodes, nN, kid, k);
            for (int64_t c = 0; c < nC; ++c)
                priv[c] += constant_factor * vector[i * vstride]
                           * nodes_to_cell(k, cell_nodes + 8 * c);
        }
        reduce_private(result, priv, nC);
        free(priv);
        free(k);
    }
}

/* _diagonal_G_T_dot_G_{serial,parallel}, lines 166-317.  NOTE the reference
 * indexes weights with the RECEIVER index i (lines 225, 302) although it is
 * handed the whole (nD,) weights array by the caller (simulation.py:715-723);
 * that is only the data index when the receiver object has ONE component.
 * The oracle mirrors this verbatim. */
void pfo_gravity_diag_gtg(const double *receivers, int64_t R,
                          const double *nodes, int64_t nN,
                          const int64_t *cell_nodes, int64_t nC,
                          const double *weights, double *diagonal,
                          int kid, double constant_factor, int parallel)
{
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN);
        double *priv = (double *)calloc((size_t)nC, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_1(receivers + 3 * i, nodes, nN, kid, k);
            for (int64_t c = 0; c < nC; ++c) {
                double g = constant_factor * nodes_to_cell(k, cell_nodes + 8 * c);
                priv[c] += weights[i] * (g * g);
            }
        }
        reduce_private(diagonal, priv, nC);
        free(priv);
        free(k);
    }
}

/* ------------------------------------------------------------------ */
/* Magnetics (simpeg/potential_fields/magnetics/_numba/_3d_mesh.py)    */
/* ------------------------------------------------------------------ */

/* The three numba families are folded into one body parameterised by the
 * number of node kernels nk: 3 = "magnetic_component" (_sensitivity_mag 21-205,
 * _forward_mag 1913-2064, ...), 6 = "tmi" (kernels hard-wired to
 * ee,nn,uu,en,eu,nu, lines 337-342) and "tmi_derivative" (six kernels passed
 * in as xx,yy,zz,xy,xz,yz, lines 375-557).  cell_vec() returns the per-cell
 * 3-vector that the reference calls (ux,uy,uz) or (bx,by,bz). */
typedef struct {
    double fx, fy, fz, amp;
} field_t;

static field_t unit_field(const double *b0)
{
    field_t f;
    double fx = b0[0], fy = b0[1], fz = b0[2];
    f.amp = sqrt(fx * fx + fy * fy + fz * fz);
    f.fx = fx / f.amp;
    f.fy = fy / f.amp;
    f.fz = fz / f.amp;
    return f;
}

static inline void cell_vec(const double *k, int64_t nN, int nk, const int64_t *idx,
                            const field_t *f, double *bx, double *by, double *bz)
{
    if (nk == 3) {
        *bx = nodes_to_cell(k, idx);
        *by = nodes_to_cell(k + nN, idx);
        *bz = nodes_to_cell(k + 2 * nN, idx);
    } else {
        double uxx = nodes_to_cell(k, idx);
        double uyy = nodes_to_cell(k + nN, idx);
        double uzz = nodes_to_cell(k + 2 * nN, idx);
        double uxy = nodes_to_cell(k + 3 * nN, idx);
        double uxz = nodes_to_cell(k + 4 * nN, idx);
        double uyz = nodes_to_cell(k + 5 * nN, idx);
        *bx = uxx * f->fx + uxy * f->fy + uxz * f->fz;
        *by = uxy * f->fx + uyy * f->fy + uyz * f->fz;
        *bz = uxz * f->fx + uyz * f->fy + uzz * f->fz;
    }
}

/* _sensitivity_mag / _sensitivity_tmi / _sensitivity_tmi_derivative */
void pfo_mag_sensitivity(const double *receivers, int64_t R,
                         const double *nodes, int64_t nN,
                         const int64_t *cell_nodes, int64_t nC,
                         void *G, int64_t row_stride, int f32,
                         const double *b0, const int *kids, int nk,
                         double constant_factor, int scalar_model, int parallel)
{
    field_t f = unit_field(b0);
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN * nk);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_n(receivers + 3 * i, nodes, nN, kids, nk, k);
            for (int64_t c = 0; c < nC; ++c) {
                double bx, by, bz;
                cell_vec(k, nN, nk, cell_nodes + 8 * c, &f, &bx, &by, &bz);
                double g0, g1 = 0, g2 = 0;
                if (scalar_model) {
                    g0 = constant_factor * f.amp * (bx * f.fx + by * f.fy + bz * f.fz);
                } else {
                    g0 = constant_factor * f.amp * bx;
                    g1 = constant_factor * f.amp * by;
                    g2 = constant_factor * f.amp * bz;
                }
                if (f32) {
                    float *row = (float *)G + i * row_stride;
                    row[c] = (float)g0;
                    if (!scalar_model) { row[c + nC] = (float)g1; row[c + 2 * nC] = (float)g2; }
                } else {
                    double *row = (double *)G + i * row_stride;
                    row[c] = g0;
                    if (!scalar_model) { row[c + nC] = g1; row[c + 2 * nC] = g2; }
                }
            }
        }
        free(k);
    }
}

/* _forward_mag / _forward_tmi / _forward_tmi_derivative, lines 1913-2360 */
void pfo_mag_forward(const double *receivers, int64_t R,
                     const double *nodes, int64_t nN,
                     const int64_t *cell_nodes, int64_t nC,
                     const double *model, double *fields, int64_t fstride,
                     const double *b0, const int *kids, int nk,
                     double constant_factor, int scalar_model, int parallel)
{
    field_t f = unit_field(b0);
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN * nk);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_n(receivers + 3 * i, nodes, nN, kids, nk, k);
            double acc = fields[i * fstride];
            for (int64_t c = 0; c < nC; ++c) {
                double bx, by, bz;
                cell_vec(k, nN, nk, cell_nodes + 8 * c, &f, &bx, &by, &bz);
                if (scalar_model)
                    acc += constant_factor * model[c] * f.amp
                           * (bx * f.fx + by * f.fy + bz * f.fz);
                else
                    acc += constant_factor * f.amp
                           * (bx * model[c] + by * model[c + nC] + bz * model[c + 2 * nC]);
            }
            fields[i * fstride] = acc;
        }
        free(k);
    }
}

/* _{mag,tmi,tmi_derivative}_sensitivity_t_dot_v_{serial,parallel}, 560-1273 */
void pfo_mag_gt_dot_v(const double *receivers, int64_t R,
                      const double *nodes, int64_t nN,
                      const int64_t *cell_nodes, int64_t nC,
                      const double *vector, int64_t vstride, double *result,
                      const double *b0, const int *kids, int nk,
                      double constant_factor, int scalar_model, int parallel)
{
    field_t f = unit_field(b0);
    int64_t nP = scalar_model ? nC : 3 * nC;
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN * nk);
        double *priv = (double *)calloc((size_t)nP, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_n(receivers + 3 * i, nodes, nN, kids, nk, k);
            double v = vector[i * vstride];
            for (int64_t c = 0; c < nC; ++c) {
                double bx, by, bz;
                cell_vec(k, nN, nk, cell_nodes + 8 * c, &f, &bx, &by, &bz);
                if (scalar_model) {
                    priv[c] += constant_factor * v * f.amp
                               * (bx * f.fx + by * f.fy + bz * f.fz);
                } else {
                    priv[c] += constant_factor * v * f.amp * bx;
                    priv[c + nC] += constant_factor * v * f.amp * by;
                    priv[c + 2 * nC] += constant_factor * v * f.amp * bz;
                }
            }
        }
        reduce_private(result, priv, nP);
        free(priv);
        free(k);
    }
}

/* _diagonal_G_T_dot_G_{mag,tmi,tmi_deriv}_{serial,parallel}, 1276-1910.
 * Same weights[i]-by-receiver-index quirk as gravity (1364-1369 etc). */
void pfo_mag_diag_gtg(const double *receivers, int64_t R,
                      const double *nodes, int64_t nN,
                      const int64_t *cell_nodes, int64_t nC,
                      const double *weights, double *diagonal,
                      const double *b0, const int *kids, int nk,
                      double constant_factor, int scalar_model, int parallel)
{
    field_t f = unit_field(b0);
    int64_t nP = scalar_model ? nC : 3 * nC;
#pragma omp parallel if (parallel)
    {
        double *k = (double *)malloc(sizeof(double) * (size_t)nN * nk);
        double *priv = (double *)calloc((size_t)nP, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            eval_nodes_n(receivers + 3 * i, nodes, nN, kids, nk, k);
            for (int64_t c = 0; c < nC; ++c) {
                double bx, by, bz;
                cell_vec(k, nN, nk, cell_nodes + 8 * c, &f, &bx, &by, &bz);
                if (scalar_model) {
                    double g = constant_factor * f.amp * (bx * f.fx + by * f.fy + bz * f.fz);
                    priv[c] += weights[i] * (g * g);
                } else {
                    double cst = constant_factor * f.amp;
                    priv[c] += weights[i] * ((cst * bx) * (cst * bx));
                    priv[c + nC] += weights[i] * ((cst * by) * (cst * by));
                    priv[c + 2 * nC] += weights[i] * ((cst * bz) * (cst * bz));
                }
            }
        }
        reduce_private(diagonal, priv, nP);
        free(priv);
        free(k);
    }
}

/* ------------------------------------------------------------------ */
/* Equivalent-source layer (2D mesh + top/bottom per cell)             */
/*   simpeg/potential_fields/gravity/_numba/_2d_mesh.py                */
/*   simpeg/potential_fields/magnetics/_numba/_2d_mesh.py              */
/*   simpeg/potential_fields/_numba_utils.py:48-255                    */
/* Cells are prisms given by bounds (x_min, x_max, y_min, y_max),      */
/* bottom and top; nothing is shared between cells.                    */
/* ------------------------------------------------------------------ */

/* evaluate_kernels_on_cell / evaluate_six_kernels_on_cell (and choclo's
 * _evaluate_kernel): signed sum over the 8 shifted vertices, east / north /
 * top first, sign (-1)^(i+j+k). */
static void eval_cell(const double *rx, const double *b4, double bottom, double top,
                      const int *kids, int nk, double *out)
{
    for (int q = 0; q < nk; ++q) out[q] = 0.0;
    for (int i = 0; i < 2; ++i) {
        double se = (i == 0 ? b4[1] : b4[0]) - rx[0];
        double se2 = se * se;
        for (int j = 0; j < 2; ++j) {
            double sn = (j == 0 ? b4[3] : b4[2]) - rx[1];
            double sn2 = sn * sn;
            for (int k = 0; k < 2; ++k) {
                double su = (k == 0 ? top : bottom) - rx[2];
                double su2 = su * su;
                double radius = sqrt(se2 + sn2 + su2);
                double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                for (int q = 0; q < nk; ++q)
                    out[q] += sign * eval_kernel(kids[q], se, sn, su, radius);
            }
        }
    }
}

#define PFO_GRAVITATIONAL_CONST 6.6743e-11            /* choclo.constants */
#define PFO_VACUUM_MAGNETIC_PERMEABILITY 1.25663706212e-6

/* choclo.prism.gravity_* (density * G * vertex sum); "guv" is
 * gravity/simulation.py:33-70 gravity_uv = 0.5 * (gravity_nn - gravity_ee) */
static double layer_gravity_forward_func(int kid, const double *rx, const double *b4,
                                         double bottom, double top, double density)
{
    double s;
    if (kid == K_UV) {
        int k_nn = K_NN, k_ee = K_EE;
        double a, b;
        eval_cell(rx, b4, bottom, top, &k_nn, 1, &a);
        eval_cell(rx, b4, bottom, top, &k_ee, 1, &b);
        return 0.5 * (PFO_GRAVITATIONAL_CONST * density * a - PFO_GRAVITATIONAL_CONST * density * b);
    }
    eval_cell(rx, b4, bottom, top, &kid, 1, &s);
    return PFO_GRAVITATIONAL_CONST * density * s;
}

/* mode 0: _sensitivity_gravity (97-160), 1: _forward_gravity (21-94),
 * 2: _g_t_dot_v_{serial,parallel} (163-295), 3: _diagonal_G_T_dot_G_* (298-434;
 * same weights[i] quirk as the 3D loops). */
void pfo_layer_gravity(int mode, const double *receivers, int64_t R,
                       const double *bounds, const double *top, const double *bottom, int64_t nC,
                       void *G, int64_t row_stride, int f32,
                       const double *vec, int64_t vstride, double *out, int64_t ostride,
                       int kid, double constant_factor, int parallel)
{
#pragma omp parallel if (parallel)
    {
        double *priv = (mode >= 2) ? (double *)calloc((size_t)nC, sizeof(double)) : NULL;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            const double *rx = receivers + 3 * i;
            double acc = (mode == 1) ? out[i * ostride] : 0.0;
            for (int64_t j = 0; j < nC; ++j) {
                /* forward: density; gt_dot_v: the reference passes vector[i] as the density */
                double dens = (mode == 1) ? vec[j] : (mode == 2 ? vec[i * vstride] : 1.0);
                double g = constant_factor
                           * layer_gravity_forward_func(kid, rx, bounds + 4 * j, bottom[j], top[j], dens);
                if (mode == 0) {
                    if (f32) ((float *)G)[i * row_stride + j] = (float)g;
                    else ((double *)G)[i * row_stride + j] = g;
                } else if (mode == 1) {
                    acc += g;
                } else if (mode == 2) {
                    priv[j] += g;
                } else {
                    priv[j] += vec[i] * (g * g);
                }
            }
            if (mode == 1) out[i * ostride] = acc;
        }
        if (mode >= 2) {
            reduce_private(out, priv, nC);
            free(priv);
        }
    }
}

/* Magnetics.  family 3 = "magnetic_component" (three kernels), 6 = "tmi" /
 * "tmi_derivative" (six kernels xx,yy,zz,xy,xz,yz).  use_choclo_forward != 0
 * selects, for mode 1, the reference's forward loops that call choclo's own
 * forward functions: _forward_mag (24-147, choclo.prism.magnetic_*) and
 * _forward_tmi (150-266, choclo.prism.magnetic_field), which scale by
 * mu0/4pi inside choclo and divide by mu0 afterwards; otherwise
 * _forward_tmi_derivative (269-431). */
void pfo_layer_mag(int mode, const double *receivers, int64_t R,
                   const double *bounds, const double *top, const double *bottom, int64_t nC,
                   void *G, int64_t row_stride, int f32,
                   const double *vec, int64_t vstride, double *out, int64_t ostride,
                   const double *b0, const int *kids, int nk, int scalar_model,
                   int use_choclo_forward, int parallel)
{
    field_t f = unit_field(b0);
    const double constant_factor = 1.0 / 4 / M_PI;
    const int64_t nP = scalar_model ? nC : 3 * nC;
#pragma omp parallel if (parallel)
    {
        double *priv = (mode >= 2) ? (double *)calloc((size_t)nP, sizeof(double)) : NULL;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < R; ++i) {
            const double *rx = receivers + 3 * i;
            double acc = (mode == 1) ? out[i * ostride] : 0.0;
            for (int64_t j = 0; j < nC; ++j) {
                const double *b4 = bounds + 4 * j;
                if (mode == 1 && use_choclo_forward) {
                    double mx, my, mz;
                    if (scalar_model) {
                        mx = vec[j] * f.fx; my = vec[j] * f.fy; mz = vec[j] * f.fz;
                    } else {
                        mx = vec[j]; my = vec[j + nC]; mz = vec[j + 2 * nC];
                    }
                    /* per vertex: M . kernels, summed with the sign; then mu0/4pi */
                    double sums[3] = {0.0, 0.0, 0.0};
                    const int ncomp = (nk == 3) ? 1 : 3;
                    /* kernel triplets of (b_e, b_n, b_u) in terms of xx,yy,zz,xy,xz,yz */
                    for (int a = 0; a < 2; ++a) {
                        double se = (a == 0 ? b4[1] : b4[0]) - rx[0];
                        double se2 = se * se;
                        for (int b = 0; b < 2; ++b) {
                            double sn = (b == 0 ? b4[3] : b4[2]) - rx[1];
                            double sn2 = sn * sn;
                            for (int c = 0; c < 2; ++c) {
                                double su = (c == 0 ? top[j] : bottom[j]) - rx[2];
                                double su2 = su * su;
                                double radius = sqrt(se2 + sn2 + su2);
                                double sign = ((a + b + c) & 1) ? -1.0 : 1.0;
                                if (ncomp == 1) {
                                    sums[0] += sign * (mx * eval_kernel(kids[0], se, sn, su, radius)
                                                       + my * eval_kernel(kids[1], se, sn, su, radius)
                                                       + mz * eval_kernel(kids[2], se, sn, su, radius));
                                } else {
                                    double kxx = eval_kernel(kids[0], se, sn, su, radius);
                                    double kyy = eval_kernel(kids[1], se, sn, su, radius);
                                    double kzz = eval_kernel(kids[2], se, sn, su, radius);
                                    double kxy = eval_kernel(kids[3], se, sn, su, radius);
                                    double kxz = eval_kernel(kids[4], se, sn, su, radius);
                                    double kyz = eval_kernel(kids[5], se, sn, su, radius);
                                    sums[0] += sign * (mx * kxx + my * kxy + mz * kxz);
                                    sums[1] += sign * (mx * kxy + my * kyy + mz * kyz);
                                    sums[2] += sign * (mx * kxz + my * kyz + mz * kzz);
                                }
                            }
                        }
                    }
                    const double c_m = PFO_VACUUM_MAGNETIC_PERMEABILITY / 4 / M_PI;
                    if (ncomp == 1) {
                        acc += f.amp * (c_m * sums[0]) / PFO_VACUUM_MAGNETIC_PERMEABILITY;
                    } else {
                        double bx = c_m * sums[0], by = c_m * sums[1], bz = c_m * sums[2];
                        acc += f.amp * (bx * f.fx + by * f.fy + bz * f.fz)
                               / PFO_VACUUM_MAGNETIC_PERMEABILITY;
                    }
                    continue;
                }
                double u[6];
                eval_cell(rx, b4, bottom[j], top[j], kids, nk, u);
                double bx, by, bz;
                if (mode == 1 && nk == 6 && !scalar_model) {
                    /* _forward_tmi_derivative, vector model: b = U m, then f . b */
                    double m_x = vec[j], m_y = vec[j + nC], m_z = vec[j + 2 * nC];
                    bx = u[0] * m_x + u[3] * m_y + u[4] * m_z;
                    by = u[3] * m_x + u[1] * m_y + u[5] * m_z;
                    bz = u[4] * m_x + u[5] * m_y + u[2] * m_z;
                    acc += f.amp * (bx * f.fx + by * f.fy + bz * f.fz) / 4 / M_PI;
                    continue;
                }
                if (nk == 3) {
                    bx = u[0]; by = u[1]; bz = u[2];
                } else {
                    bx = u[0] * f.fx + u[3] * f.fy + u[4] * f.fz;
                    by = u[3] * f.fx + u[1] * f.fy + u[5] * f.fz;
                    bz = u[4] * f.fx + u[5] * f.fy + u[2] * f.fz;
                }
                if (mode == 0) {
                    double g0, g1 = 0, g2 = 0;
                    if (scalar_model) {
                        g0 = constant_factor * f.amp * (bx * f.fx + by * f.fy + bz * f.fz);
                    } else {
                        g0 = constant_factor * f.amp * bx;
                        g1 = constant_factor * f.amp * by;
                        g2 = constant_factor * f.amp * bz;
                    }
                    if (f32) {
                        float *row = (float *)G + i * row_stride;
                        row[j] = (float)g0;
                        if (!scalar_model) { row[j + nC] = (float)g1; row[j + 2 * nC] = (float)g2; }
                    } else {
                        double *row = (double *)G + i * row_stride;
                        row[j] = g0;
                        if (!scalar_model) { row[j + nC] = g1; row[j + 2 * nC] = g2; }
                    }
                } else if (mode == 1) {
                    /* _forward_tmi_derivative, scalar model (405-421) */
                    acc += vec[j] * f.amp * (f.fx * bx + f.fy * by + f.fz * bz) / (4 * M_PI);
                } else if (mode == 2) {
                    double v = vec[i * vstride];
                    if (scalar_model) {
                        priv[j] += constant_factor * v * f.amp * (bx * f.fx + by * f.fy + bz * f.fz);
                    } else {
                        priv[j] += constant_factor * v * f.amp * bx;
                        priv[j + nC] += constant_factor * v * f.amp * by;
                        priv[j + 2 * nC] += constant_factor * v * f.amp * bz;
                    }
                } else {
                    if (scalar_model) {
                        double g = constant_factor * f.amp * (bx * f.fx + by * f.fy + bz * f.fz);
                        priv[j] += vec[i] * (g * g);
                    } else {
                        double cst = constant_factor * f.amp;
                        priv[j] += vec[i] * ((cst * bx) * (cst * bx));
                        priv[j + nC] += vec[i] * ((cst * by) * (cst * by));
                        priv[j + 2 * nC] += vec[i] * ((cst * bz) * (cst * bz));
                    }
                }
            }
            if (mode == 1) out[i * ostride] = acc;
        }
        if (mode >= 2) {
            reduce_private(out, priv, nP);
            free(priv);
        }
    }
}
