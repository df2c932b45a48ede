// pf_plan.cu -- host-side tile planner for tensor meshes (no CUDA calls).
//
// The generic planner lives in Python (simpeg_b200/tiling.py: recursive
// coordinate bisection for tree meshes / arbitrary cell_nodes tables).  For a
// tensor mesh the plan is pure index arithmetic; doing it natively keeps the
// one-off cost of `sim.G` at a few milliseconds per million cells.
//
// Replaces the mesh -> (nodes, cell_nodes) extraction of
// BasePFSimulation._get_active_nodes (base.py:371-399) for TensorMesh inputs:
// the tile arrays are produced straight from the 1-D node coordinates and the
// list of active cells, without materialising (nodes, cell_nodes).
#include <algorithm>
#include <vector>

#include "pf_common.cuh"

namespace {
struct Grid {
    int64_t nx, ny, nz, bx, by, bz, nbx, nby, nbz;
};

Grid make_grid(int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by, int64_t bz)
{
    Grid g{nx, ny, nz, bx, by, bz, 0, 0, 0};
    if (bx > 0 && by > 0 && bz > 0) {
        g.nbx = pf_ceil_div(nx, bx);
        g.nby = pf_ceil_div(ny, by);
        g.nbz = pf_ceil_div(nz, bz);
    }
    return g;
}

bool valid(const Grid &g)
{
    return g.nx > 0 && g.ny > 0 && g.nz > 0 && g.bx > 0 && g.by > 0 && g.bz > 0 &&
           g.bx * g.by * g.bz <= PF_TILE_CELLS &&
           (g.bx + 1) * (g.by + 1) * (g.bz + 1) <= PF_TILE_NODES;
}
}  // namespace

// Marks the bricks that contain active cells (used[b] = 1) and returns their
// number, or a negative pf_status.  `active` may be NULL (all cells active).
extern "C" int64_t pf_plan_tensor_count(int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by,
                                        int64_t bz, const int64_t *active, int64_t n_active,
                                        uint8_t *used)
{
    const Grid g = make_grid(nx, ny, nz, bx, by, bz);
    if (!valid(g) || !used) {
        pf_set_error("pf_plan_tensor_count: bad grid / brick dimensions");
        return PF_ERR_INVALID;
    }
    const int64_t nb = g.nbx * g.nby * g.nbz;
    if (!active) {
        for (int64_t b = 0; b < nb; ++b) used[b] = 1;
        return nb;
    }
    for (int64_t b = 0; b < nb; ++b) used[b] = 0;
    const int64_t n_mesh = nx * ny * nz;
    for (int64_t k = 0; k < n_active; ++k) {
        const int64_t c = active[k];
        if (c < 0 || c >= n_mesh) {
            pf_set_error("pf_plan_tensor_count: active cell index out of range");
            return PF_ERR_INVALID;
        }
        const int64_t ix = c % nx, iy = (c / nx) % ny, iz = c / (nx * ny);
        used[(ix / bx) + g.nbx * ((iy / by) + g.nby * (iz / bz))] = 1;
    }
    int64_t n = 0;
    for (int64_t b = 0; b < nb; ++b) n += used[b];
    return n;
}

// Fills the host arrays of the plan (layouts of pf_plan).  `used` comes from
// pf_plan_tensor_count; n_tiles is its return value.
extern "C" int pf_plan_tensor_fill(int64_t nx, int64_t ny, int64_t nz, const double *xs,
                                   const double *ys, const double *zs, int64_t bx, int64_t by,
                                   int64_t bz, const int64_t *active, int64_t n_active,
                                   const uint8_t *used, int64_t n_tiles, double *tile_xyz,
                                   uint16_t *tile_cell_nodes, int32_t *tile_cell_col,
                                   int32_t *tile_n_nodes)
{
    const Grid g = make_grid(nx, ny, nz, bx, by, bz);
    PF_CHECK_ARG(valid(g), "bad grid / brick dimensions");
    PF_CHECK_ARG(xs && ys && zs && used && tile_xyz && tile_cell_nodes && tile_cell_col && tile_n_nodes,
                 "null pointer");
    const int64_t nb = g.nbx * g.nby * g.nbz;
    std::vector<int64_t> tile_of_brick(nb, -1);
    std::vector<int64_t> brick_of_tile(n_tiles);
    int64_t t = 0;
    for (int64_t b = 0; b < nb; ++b)
        if (used[b]) {
            PF_CHECK_ARG(t < n_tiles, "n_tiles does not match `used`");
            tile_of_brick[b] = t;
            brick_of_tile[t++] = b;
        }
    PF_CHECK_ARG(t == n_tiles, "n_tiles does not match `used`");

#pragma omp parallel for schedule(static) num_threads(4)
    for (int64_t i = 0; i < n_tiles * (int64_t)PF_TILE_CELLS; ++i) tile_cell_col[i] = -1;

    // node coordinates and node counts per tile
#pragma omp parallel for schedule(static) num_threads(4)
    for (int64_t tt = 0; tt < n_tiles; ++tt) {
        const int64_t b = brick_of_tile[tt];
        const int64_t tx = b % g.nbx, ty = (b / g.nbx) % g.nby, tz = b / (g.nbx * g.nby);
        const int64_t bw = std::min(bx, nx - tx * bx), bh = std::min(by, ny - ty * by),
                      bd = std::min(bz, nz - tz * bz);
        const int64_t w1 = bw + 1, h1 = bh + 1, d1 = bd + 1;
        tile_n_nodes[tt] = (int32_t)(w1 * h1 * d1);
        double *X = tile_xyz + tt * 3 * PF_TILE_NODES;
        for (int64_t ln = 0; ln < PF_TILE_NODES; ++ln) {
            double x = 0.0, y = 0.0, z = 0.0;
            if (ln < w1 * h1 * d1) {
                x = xs[tx * bx + ln % w1];
                y = ys[ty * by + (ln / w1) % h1];
                z = zs[tz * bz + ln / (w1 * h1)];
            }
            X[ln] = x;
            X[PF_TILE_NODES + ln] = y;
            X[2 * PF_TILE_NODES + ln] = z;
        }
    }

    // cells: slot = position inside the brick (ascending with the G column)
    const int64_t n_cells = active ? n_active : nx * ny * nz;
#pragma omp parallel for schedule(static) num_threads(4)
    for (int64_t k = 0; k < n_cells; ++k) {
        const int64_t c = active ? active[k] : k;
        const int64_t ix = c % nx, iy = (c / nx) % ny, iz = c / (nx * ny);
        const int64_t tx = ix / bx, ty = iy / by, tz = iz / bz;
        const int64_t lx = ix - tx * bx, ly = iy - ty * by, lz = iz - tz * bz;
        const int64_t tt = tile_of_brick[tx + g.nbx * (ty + g.nby * tz)];
        const int64_t bw = std::min(bx, nx - tx * bx), bh = std::min(by, ny - ty * by);
        const int64_t slot = tt * PF_TILE_CELLS + lx + bw * (ly + bh * lz);
        tile_cell_col[slot] = (int32_t)k;
        uint16_t *cn = tile_cell_nodes + slot * 8;
        const int64_t w1 = bw + 1, h1 = bh + 1;
        for (int v = 0; v < 8; ++v) {
            const int64_t dx = v & 1, dy = (v >> 1) & 1, dz = v >> 2;
            cn[v] = (uint16_t)((lx + dx) + w1 * ((ly + dy) + h1 * (lz + dz)));
        }
    }
    return PF_OK;
}


// Host staging copy with a few threads (HOST pointers): numpy vector <-> pinned staging buffer of
// the H2D / D2H transfers around every Jvec / Jtvec call.  A single-threaded copy of a 24 MB
// vector (and the first touch of a fresh 24 MB result array) cost as much as the 60 GB product
// they bracket; 4 threads bring both under a millisecond.
extern "C" int pf_host_copy(void *dst, const void *src, int64_t bytes)
{
    if (bytes <= 0) return PF_OK;
    PF_CHECK_ARG(dst != nullptr && src != nullptr, "null pointer");
    constexpr int64_t CHUNK = 1 << 20;
    const int64_t n_chunks = (bytes + CHUNK - 1) / CHUNK;
    if (n_chunks == 1) {
        memcpy(dst, src, (size_t)bytes);
        return PF_OK;
    }
#pragma omp parallel for schedule(static) num_threads(4)
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int64_t o = c * CHUNK;
        const int64_t len = (o + CHUNK <= bytes) ? CHUNK : bytes - o;
        memcpy((char *)dst + o, (const char *)src + o, (size_t)len);
    }
    return PF_OK;
}
