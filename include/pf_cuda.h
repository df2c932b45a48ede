/*
 * pf_cuda.h -- C ABI of libpf_cuda.so: the B200 (sm_100a) engine for SimPEG's
 * integral-equation potential-field hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI on this path: the
 * hot loops are numba-jitted Python functions selected through the dict
 * NUMBA_FUNCTIONS_3D and called from four private methods of
 * gravity.Simulation3DIntegral / magnetics.Simulation3DIntegral.  Each entry
 * point below names the reference function(s) it replaces (paths relative to
 * the reference root, simpeg/potential_fields/...).  INTEGRATION.md shows the
 * ctypes binding a SimPEG maintainer would add for engine="cuda".
 *
 * Conventions
 *  - every function returns 0 on success, a negative pf_status otherwise; the
 *    message of the last error of the calling thread is pf_last_error();
 *  - nothing throws or exits across the ABI;
 *  - all array pointers are DEVICE pointers owned by the caller (torch
 *    tensors on the host side); the library allocates nothing that outlives a
 *    call; scratch memory is passed in as `workspace` (size from the matching
 *    *_workspace_bytes query);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - all arithmetic is IEEE fp64; G may be stored as fp64 or fp32
 *    (sensitivity_dtype, base.py:196-215).
 */
#ifndef PF_CUDA_H
#define PF_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 6

typedef enum {
    PF_OK = 0,
    PF_ERR_INVALID = -1,     /* bad argument (null pointer, bad enum, bad size) */
    PF_ERR_UNSUPPORTED = -2, /* component set not supported on this physics      */
    PF_ERR_WORKSPACE = -3,   /* workspace too small                              */
    PF_ERR_CUDA = -4         /* CUDA runtime error (message has the cudaError)   */
} pf_status;

/* Component codes.  Gravity: gravity/receivers.py:56-71 and the kernel table
 * gravity/simulation.py:74-85.  Magnetics: magnetics/receivers.py:46-62 and
 * magnetics/simulation.py:29-103 (CHOCLO_SUPPORTED_COMPONENTS). */
typedef enum {
    PF_GX = 0, PF_GY = 1, PF_GZ = 2,
    PF_GXX = 3, PF_GYY = 4, PF_GZZ = 5, PF_GXY = 6, PF_GXZ = 7, PF_GYZ = 8,
    PF_GUV = 9,
    PF_TMI = 16, PF_BX = 17, PF_BY = 18, PF_BZ = 19,
    PF_BXX = 20, PF_BYY = 21, PF_BZZ = 22, PF_BXY = 23, PF_BXZ = 24, PF_BYZ = 25,
    PF_TMI_X = 26, PF_TMI_Y = 27, PF_TMI_Z = 28
} pf_component;

typedef enum { PF_GRAVITY = 0, PF_MAGNETICS = 1 } pf_physics;
typedef enum { PF_F64 = 0, PF_F32 = 1 } pf_dtype;

#define PF_MAX_COMPONENTS 16

/*
 * Tile plan: the device-resident form of what the reference obtains from
 * BasePFSimulation._get_active_nodes() (base.py:371-399): the active nodes
 * and the (nC, 8) cell -> node table.  Active cells are partitioned into
 * tiles of at most `tile_cells` cells whose union of vertices has at most
 * `tile_nodes` nodes; a CTA evaluates the prism kernels once per (receiver,
 * tile node) in shared memory and combines them per cell (the same
 * node-sharing the reference gets from its per-receiver `kernels` array,
 * gravity/_numba/_3d_mesh.py:147-163).
 */
typedef struct {
    int64_t n_tiles;
    int64_t n_cells;               /* nC: active cells = columns per block of G */
    int32_t tile_cells;            /* TC: slots per tile (multiple of 256)      */
    int32_t tile_nodes;            /* TN: node slots per tile (multiple of 256) */
    const double *tile_xyz;        /* [n_tiles][3][TN] node coordinates (SoA)   */
    const uint16_t *tile_cell_nodes; /* [n_tiles][TC][8] local node ids, F-order  */
    const int32_t *tile_cell_col;  /* [n_tiles][TC] global cell index, -1 = pad */
    const int32_t *tile_n_nodes;   /* [n_tiles] nodes used in each tile         */
    /*
     * Optional tensor-mesh description (grid_nx == 0: absent).  When present,
     * pf_build_G and pf_forward use the z-marching kernel: a CTA owns an (x, y)
     * footprint of the mesh and sweeps the node layers bottom-up, keeping the
     * last two layers of node kernels in shared memory, so that every node is
     * evaluated once per footprint instead of once per brick.  Cell (ix,iy,iz)
     * has mesh index ix + nx (iy + ny iz); its G column is colmap[index]
     * (colmap == NULL: all cells active, column = index; -1 = inactive).
     * Footprints are given for two node capacities per layer: 1024 ("4") and
     * 512 ("2") nodes; each is (ix0, iy0, width, height, iz_begin, iz_end): an
     * (x, y) rectangle of cells with (width + 1)(height + 1) <= capacity,
     * width + 1 <= 256 (a thread of the 256-thread CTA keeps one node column; a
     * wider footprint traps the kernel) and the range of cell layers
     * [iz_begin, iz_end) that holds its active cells.  A footprint costs
     * ceil((height + 1) / floor(256 / (width + 1))) node rounds per layer.
     * footprints4h (optional, n_footprints4h == 0: absent) is a set for the
     * 128-thread variant of the kernel, which the three-plane programs prefer
     * (twice as many, smaller CTAs per SM): (width + 1)(height + 1) <= 512,
     * width + 1 <= 128, node rounds counted with 128.
     */
    int32_t grid_nx, grid_ny, grid_nz;
    int32_t n_footprints4, n_footprints2;
    int32_t n_footprints4h;
    const double *grid_xs;         /* [nx+1] node coordinates                   */
    const double *grid_ys;         /* [ny+1]                                    */
    const double *grid_zs;         /* [nz+1] ascending                          */
    const int32_t *grid_colmap;    /* [nx ny nz] or NULL                        */
    const int32_t *footprints4;    /* [n_footprints4][6]                        */
    const int32_t *footprints2;    /* [n_footprints2][6]                        */
    /*
     * Optional global node table (n_nodes == 0: absent): the active nodes and
     * the (nC, 8) cell -> node table exactly as _get_active_nodes returns them
     * (base.py:371-399), node coordinates as SoA.  When present, pf_gt_dot_v
     * works in node space: sum_r v_r k(node - receiver_r) is accumulated once per
     * unique node and the 8-vertex combination is applied once at the end (the
     * vertex sum is linear), instead of once per (receiver, cell).
     */
    int64_t n_nodes;
    const double *node_xyz;        /* [3][n_nodes]                              */
    const int32_t *cell_nodes32;   /* [n_cells][8], F-order vertices            */
    const int32_t *footprints4h;   /* [n_footprints4h][6] (see grid_nx above)   */
} pf_plan;

/*
 * One receiver object = one (components, locations) pair of
 * BasePFSimulation._get_components_and_receivers() (base.py:401-408).
 * Rows are location-major, component-minor:
 *     row = row_offset + i_location * n_components + i_component
 * (vector_slice / matrix_slice of gravity/simulation.py:570-572, 617-619).
 */
typedef struct {
    int32_t physics;               /* pf_physics */
    int32_t n_components;
    int32_t components[PF_MAX_COMPONENTS]; /* pf_component, in receiver order */
    int32_t scalar_model;          /* magnetics: 1 = scalar chi, 0 = vector (3 nC columns) */
    int32_t reserved;
    double b0[3];                  /* magnetics: inducing field (survey.source_field.b0)   */
    const double *receivers;       /* [n_receivers][3] (x, y, z), device                   */
    int64_t n_receivers;
    int64_t row_offset;            /* first row of this object in G / dpred / v / w        */
} pf_job;

int pf_abi_version(void);
const char *pf_last_error(void);
/* number of CUDA kernels this library has launched in this process so far */
int64_t pf_launch_count(void);

/* 1 if every pass of the job runs on the z-marching kernel (tensor-mesh description present
 * and the output planes fit), so that the tile arrays of the plan are not needed for
 * pf_build_G (forward == 0) / pf_forward (forward != 0); 0 otherwise. */
int pf_job_uses_marching(const pf_plan *plan, const pf_job *job, int forward);

/* Names of the kernels (template instantiations) the passes of a job launch for a sink
 * (0 = store G, 1 = forward, 2 = G^T v, 3 = diag), joined by " + ", written to buf (NUL-terminated,
 * truncated at a kernel boundary): the label bench.py reports as roofline.kernel. */
int pf_describe(const pf_plan *plan, const pf_job *job, int sink, int g_dtype, char *buf, int64_t len);

/* number of columns of G for a job: nC (gravity, scalar) or 3 nC (vector) */
int64_t pf_job_columns(const pf_plan *plan, const pf_job *job);

/*
 * pf_build_G -- fill the rows of G that belong to one receiver object.
 * Replaces: _sensitivity_gravity (gravity/_numba/_3d_mesh.py:97-163) as
 * dispatched by gravity/simulation.py:585-629, and _sensitivity_mag /
 * _sensitivity_tmi / _sensitivity_tmi_derivative
 * (magnetics/_numba/_3d_mesh.py:21-557) as dispatched by
 * magnetics/simulation.py:931-1026.  All components of the object are produced
 * in one pass.  G is row-major with leading dimension ldG (elements).
 */
int pf_build_G(const pf_plan *plan, const pf_job *job, void *G, int64_t ldG,
               int g_dtype, void *stream);

/*
 * pf_forward -- fields[row] += sum_k G[row,k] * model[k] without storing G
 * (store_sensitivities="forward_only").
 * Replaces: _forward_gravity (gravity/_numba/_3d_mesh.py:21-94; dispatch
 * gravity/simulation.py:541-583) and _forward_mag/_forward_tmi/
 * _forward_tmi_derivative (magnetics/_numba/_3d_mesh.py:1913-2360; dispatch
 * magnetics/simulation.py:829-929).  `fields` is the full (nD,) vector and is
 * accumulated into, like the reference.
 */
int64_t pf_forward_workspace_bytes(const pf_plan *plan, const pf_job *job);
int pf_forward(const pf_plan *plan, const pf_job *job, const double *model,
               double *fields, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * pf_gt_dot_v -- result[k] += sum_rows G[row,k] * v[row] without storing G.
 * Replaces: _sensitivity_gravity_t_dot_v_{serial,parallel}
 * (gravity/_numba/_3d_mesh.py:320-469; dispatch gravity/simulation.py:631-671)
 * and _{mag,tmi,tmi_derivative}_sensitivity_t_dot_v_*
 * (magnetics/_numba/_3d_mesh.py:560-1273; dispatch magnetics/simulation.py:
 * 1046-1141).  `v` is the full (nD,) vector.
 */
int64_t pf_gt_dot_v_workspace_bytes(const pf_plan *plan, const pf_job *job);
int pf_gt_dot_v(const pf_plan *plan, const pf_job *job, const double *v,
                double *result, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * pf_gtg_diag -- diag[k] += sum_rows w[row] * G[row,k]^2 without storing G.
 * Replaces: _diagonal_G_T_dot_G_{serial,parallel}
 * (gravity/_numba/_3d_mesh.py:166-317; dispatch gravity/simulation.py:690-724)
 * and _diagonal_G_T_dot_G_{mag,tmi,tmi_deriv}_* (magnetics/_numba/_3d_mesh.py:
 * 1276-1910; dispatch magnetics/simulation.py:1143-1232).  `w` is the full
 * (nD,) vector of squared data weights and is indexed by ROW (the reference
 * indexes it by receiver, which is only right for one-component objects; the
 * stored-G definition np.einsum("i,ij,ij->j", w, G, G), gravity/simulation.py:
 * 313-315, is what this implements).
 */
int64_t pf_gtg_diag_workspace_bytes(const pf_plan *plan, const pf_job *job);
int pf_gtg_diag(const pf_plan *plan, const pf_job *job, const double *w,
                double *diag, void *workspace, int64_t workspace_bytes, void *stream);

/*
 * Stored-G products (store_sensitivities="ram").  G is rows x cols, row-major,
 * leading dimension ld, dtype fp64 or fp32; vectors are fp64; accumulation is
 * fp64.  Replace the numpy/BLAS expressions
 *   pf_gemv      y = G @ x        gravity/simulation.py:220,387  magnetics/simulation.py:252,489,540
 *   pf_gemv_t    z = G.T @ y      gravity/simulation.py:418      magnetics/simulation.py:533
 *   pf_colsumsq  d = einsum("i,ij,ij->j", w, G, G)   gravity/simulation.py:313-315  magnetics/simulation.py:441-443
 * All three OVERWRITE their output.
 */
int64_t pf_gemv_workspace_bytes(int64_t rows, int64_t cols);
int pf_gemv(const void *G, int64_t rows, int64_t cols, int64_t ld, int g_dtype,
            const double *x, double *y, void *workspace, int64_t workspace_bytes,
            void *stream);
int64_t pf_gemv_t_workspace_bytes(int64_t rows, int64_t cols);
int pf_gemv_t(const void *G, int64_t rows, int64_t cols, int64_t ld, int g_dtype,
              const double *y, double *z, void *workspace, int64_t workspace_bytes,
              void *stream);
int pf_colsumsq(const void *G, int64_t rows, int64_t cols, int64_t ld, int g_dtype,
                const double *w, double *d, void *workspace, int64_t workspace_bytes,
                void *stream);
/* d += einsum("i,ij,ij->j", w, G, G): accumulating variant (workspace as pf_gemv_t).  pf_gtg_diag uses
 * it on tensor meshes, where the diagonal is formed a block of receivers at a time from rows the
 * z-marching kernel writes into scratch memory. */
int pf_colsumsq_acc(const void *G, int64_t rows, int64_t cols, int64_t ld, int g_dtype,
                    const double *w, double *d, void *workspace, int64_t workspace_bytes,
                    void *stream);

/*
 * Amplitude data (is_amplitude_data=True, magnetics/simulation.py:444-456,
 * 491-495, 529-565).  Rows come in (bx, by, bz) triplets per location.
 *   pf_amp_normalize   amp[i] = |B_i|, ampderiv[c][i] = B_i[c] / |B_i|   (normalized_fields, 546-565)
 *   pf_amp_combine     out[i] = sum_c ampderiv[c][i] * jv[3 i + c]       (Jvec, 491-495)
 *   pf_amp_spread      out[3 i + c] = ampderiv[c][i] * v[i]              (Jtvec, 529-532)
 *   pf_colsumsq_amp    d[k] = sum_i w[i] (sum_c ampderiv[c][i] G[3i+c,k])^2   (444-456)
 * ampderiv is stored (3, n_loc) row-major like the reference's array.
 */
int pf_amp_normalize(const double *b_xyz, int64_t n_loc, double *amp, double *ampderiv,
                     void *stream);
int pf_amp_combine(const double *ampderiv, const double *jv, int64_t n_loc, double *out,
                   void *stream);
int pf_amp_spread(const double *ampderiv, const double *v, int64_t n_loc, double *out,
                  void *stream);
int pf_colsumsq_amp(const void *G, int64_t n_loc, int64_t cols, int64_t ld, int g_dtype,
                    const double *w, const double *ampderiv, double *d, void *workspace,
                    int64_t workspace_bytes, void *stream);

/*
 * Host-side tile planner for tensor meshes (HOST pointers; no CUDA calls).
 * Produces the arrays of pf_plan from the 1-D node coordinates and the list of
 * active cells (mesh indices, x fastest, ascending; NULL = all active), i.e.
 * what BasePFSimulation._get_active_nodes (base.py:371-399) feeds the numba
 * loops with, already cut into bricks of (bx, by, bz) cells.  Call
 * pf_plan_tensor_count first (it marks the bricks that hold active cells in
 * `used`, one byte per brick, and returns how many: n_tiles), allocate
 * tile_xyz[n_tiles][3][768], tile_cell_nodes[n_tiles][512][8],
 * tile_cell_col[n_tiles][512], tile_n_nodes[n_tiles], then call
 * pf_plan_tensor_fill.  Column k of G is the k-th active cell.
 */
int64_t pf_plan_tensor_count(int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by,
                             int64_t bz, const int64_t *active, int64_t n_active, uint8_t *used);
int pf_plan_tensor_fill(int64_t nx, int64_t ny, int64_t nz, const double *xs, const double *ys,
                        const double *zs, int64_t bx, int64_t by, int64_t bz,
                        const int64_t *active, int64_t n_active, const uint8_t *used,
                        int64_t n_tiles, double *tile_xyz, uint16_t *tile_cell_nodes,
                        int32_t *tile_cell_col, int32_t *tile_n_nodes);

/*
 * Device-resident inner loop of the Gauss-Newton step: the vector algebra of
 * ProjectedGNCG.findSearchDirection (simpeg/optimization.py:1625-1711) around the two products
 * over G that make up L2DataMisfit.deriv2 (simpeg/data_misfit.py:341-343,
 * 2 J^T W^T W J v).  All vectors have n = nP entries and stay on the device; the CG scalars
 * live in `scal` (>= 4 doubles): scal[0] = sold, scal[1] = p.q, scal[2] = r.r, scal[3] = snew.
 * Dot products are reduced in a fixed order (deterministic, identical on every rank).
 *   pf_cg_init         step = 0; r = inactive * (-g); p = minv * r; scal[2] = r.r; scal[0] = r.p   (1638-1652)
 *   pf_cg_hvp_combine  q = inactive * (scale * d * z + beta * R p); scal[1] = p.q                (1664)
 *                      z = G^T W^2 G (d * p) (already summed over ranks); d = diagonal of the
 *                      model-map derivative or NULL; R = regularisation Hessian as CSR
 *                      (indptr, indices, values; int32) or diagonal `rdiag`, or both NULL
 *   pf_cg_update       alpha = scal[0] / scal[1]; step += alpha p; r -= alpha q; h = minv * r;
 *                      scal[2] = r.r; scal[3] = r.h                                              (1666-1676)
 *   pf_cg_direction    p = h + (scal[3] / scal[0]) p; scal[0] = scal[3]                          (1678-1682)
 *   pf_vec_mul         out = a * b (row weights W^2, chain rule with a diagonal derivative)
 * `inactive` is one byte per entry (1 = free variable, 0 = on a bound) or NULL; `minv` is the
 * diagonal of approxHinv (directives/_directives.py:2205-2233 UpdatePreconditioner) or NULL.
 */
int64_t pf_cg_workspace_bytes(int64_t n);
int pf_cg_init(const double *g, const uint8_t *inactive, const double *minv, int64_t n,
               double *step, double *r, double *p, double *scal, void *workspace,
               int64_t workspace_bytes, void *stream);
int pf_cg_hvp_combine(const double *z, const double *d, const double *p, const int32_t *indptr,
                      const int32_t *indices, const double *values, const double *rdiag,
                      double scale, double beta, const uint8_t *inactive, int64_t n, double *q,
                      double *scal, void *workspace, int64_t workspace_bytes, void *stream);
int pf_cg_update(const double *p, const double *q, const double *minv, int64_t n, double *step,
                 double *r, double *h, double *scal, void *workspace, int64_t workspace_bytes,
                 void *stream);
int pf_cg_direction(const double *h, int64_t n, double *p, double *scal, void *stream);
int pf_vec_mul(const double *a, const double *b, int64_t n, double *out, void *stream);
/* y = A x, A in CSR (int32 indptr [n_rows + 1] / indices, fp64 values): the sparse model-map derivative
 * (maps.TileMap, maps.Wires, any LinearMap; gravity/simulation.py:386-387, 418-419 apply it on the host)
 * and its transpose around pf_gemv / pf_gemv_t inside the device CG loop. */
int pf_csr_spmv(const int32_t *indptr, const int32_t *indices, const double *values, int64_t n_rows,
                const double *x, double *y, void *stream);

/*
 * Host-side generic tile planner (HOST pointers; no CUDA calls): recursive coordinate bisection of
 * the (nodes, cell_nodes) tables BasePFSimulation._get_active_nodes returns (base.py:371-399) --
 * tree meshes with hanging nodes, equivalent-source layers.  pf_plan_generic_assign writes the
 * tile of every cell (<= 512 cells and <= 768 distinct nodes per tile) and returns the number of
 * tiles; pf_plan_generic_fill then fills tile_xyz[n_tiles][3][768], tile_cell_nodes[n_tiles][512][8],
 * tile_cell_col[n_tiles][512] (-1 = padding) and tile_n_nodes[n_tiles].
 */
int64_t pf_plan_generic_assign(const double *nodes, int64_t n_nodes, const int64_t *cell_nodes,
                               int64_t n_cells, int32_t *tile_of_cell);
int pf_plan_generic_fill(const double *nodes, int64_t n_nodes, const int64_t *cell_nodes,
                         int64_t n_cells, const int32_t *tile_of_cell, int64_t n_tiles,
                         double *tile_xyz, uint16_t *tile_cell_nodes, int32_t *tile_cell_col,
                         int32_t *tile_n_nodes);

/* Host staging copy (HOST pointers, a few threads): numpy vector <-> pinned staging buffer of the
 * H2D / D2H transfers around the products (the boundary cost of SURVEY.md 8(b) "Callers across
 * the boundary"). */
int pf_host_copy(void *dst, const void *src, int64_t bytes);

/* FP64 pipe peak probe: runs a dependent-free DFMA loop and returns the
 * achieved FLOP/s (2 per DFMA) in *flops; used by bench.py for the roofline
 * denominator of the compute-bound build kernels. */
int pf_probe_fp64_peak(double *flops, void *stream);

/* Element-wise probe of the fp64 building blocks of the prism kernels (tests/test_gpu_math.py):
 * kind 0 = a / b, 1 = atan(a / b) (choclo _safe_atan2(a, b)), 2 = log(a / b) for a, b > 0,
 * 3 = choclo _safe_log(a, b, c, r) with r = sqrt(a^2 + b^2 + c^2), 4 = sqrt(a) for a == 0 or a normal
 * double (bit-identical to the IEEE root).  `c` may be NULL for every kind but 3. */
int pf_math_probe(int kind, const double *a, const double *b, const double *c, double *out, int64_t n,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PF_CUDA_H */
