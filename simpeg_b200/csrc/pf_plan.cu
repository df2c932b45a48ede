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
#include <cmath>
#include <utility>
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


// ---------------------------------------------------------------------------
// Generic tile planner: recursive coordinate bisection of arbitrary (nodes, cell_nodes) tables --
// tree meshes with hanging nodes, equivalent-source layers, anything _get_active_nodes
// (base.py:371-399) can return.  Same rule as the numpy version it replaces
// (simpeg_b200/tiling.py::_rcb_tile_ids): a region is a tile when it has <= 512 cells whose
// vertices are <= 768 distinct nodes; otherwise it is cut, across its longest axis, BETWEEN two
// distinct cell-centre coordinates (cells in one plane stay together) -- the numpy statement of
// the rule (simpeg_b200/tiling.py::_rcb_tile_ids) cuts as close to the median as possible, this one
// as close as possible to the position that leaves a whole number of expected tiles on either side
// (leaf_capacity), which fills the tiles to ~90 % instead of ~65 %.  Every level is a handful of O(n)
// passes (selection, not sorting).
// ---------------------------------------------------------------------------
namespace {

struct RcbPlanner {
    const int64_t *cn;
    int64_t n_nodes;
    std::vector<double> c[3];         // cell centres
    std::vector<int32_t> ids;         // permutation of the cells, regions are sub-ranges
    std::vector<int32_t> stamp;       // last region id that touched a node
    std::vector<double> vals;         // scratch: coordinates of a region
    int32_t stamp_id = 0;
    int32_t *tile_of_cell;
    int32_t n_tiles = 0;

    bool fits(int64_t lo, int64_t hi)
    {
        if (hi - lo > PF_TILE_CELLS) return false;
        ++stamp_id;
        int count = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t *v = cn + 8 * (int64_t)ids[i];
            for (int k = 0; k < 8; ++k) {
                if (stamp[v[k]] != stamp_id) {
                    stamp[v[k]] = stamp_id;
                    if (++count > PF_TILE_NODES) return false;
                }
            }
        }
        return true;
    }

    // position m in (lo, hi) such that [lo, m) / [m, hi) are the two sides of the cut, or -1 when
    // every centre of the region has the same coordinate along `ax`
    int64_t cut(int64_t lo, int64_t hi, int ax, int64_t pos)
    {
        const int64_t n = hi - lo;
        const std::vector<double> &x = c[ax];
        vals.resize((size_t)n);
        double vmin = x[ids[lo]], vmax = vmin;
        for (int64_t i = 0; i < n; ++i) {
            const double v = x[ids[lo + i]];
            vals[i] = v;
            vmin = std::min(vmin, v);
            vmax = std::max(vmax, v);
        }
        if (!(vmax > vmin)) return -1;
        // the coordinate at sorted position `pos` (pos = n / 2: the median; np.median's mean of the
        // two middle elements for even n)
        std::nth_element(vals.begin(), vals.begin() + pos, vals.end());
        double med = vals[pos];
        if (pos == n / 2 && n % 2 == 0) med = 0.5 * (med + *std::max_element(vals.begin(), vals.begin() + pos));
        // distinct neighbours of the median: a = largest value <= med, a_prev = largest < a,
        // b = smallest > med
        double a = vmin, b = vmax;
        bool has_b = false;
        for (int64_t i = 0; i < n; ++i) {
            const double v = vals[i];
            if (v <= med) { if (v > a) a = v; }
            else if (!has_b || v < b) { b = v; has_b = true; }
        }
        double a_prev = 0.0;
        bool has_prev = false;
        for (int64_t i = 0; i < n; ++i) {
            const double v = vals[i];
            if (v < a && (!has_prev || v > a_prev)) { a_prev = v; has_prev = true; }
        }
        double best = 0.0;
        bool have = false;
        if (has_prev) { best = 0.5 * (a_prev + a); have = true; }      // lower cut first (argmin ties)
        if (has_b) {
            const double c2 = 0.5 * (a + b);
            if (!have || std::fabs(c2 - med) < std::fabs(best - med)) best = c2;
            have = true;
        }
        if (!have) return -1;
        const double cutv = best;
        int32_t *first = ids.data() + lo, *last = ids.data() + hi;
        int32_t *mid = std::stable_partition(first, last, [&](int32_t id) { return x[id] < cutv; });
        const int64_t m = lo + (mid - first);
        return (m > lo && m < hi) ? m : -1;
    }

    // distinct nodes of a region
    int64_t count_nodes(int64_t lo, int64_t hi)
    {
        ++stamp_id;
        int64_t count = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t *v = cn + 8 * (int64_t)ids[i];
            for (int k = 0; k < 8; ++k)
                if (stamp[v[k]] != stamp_id) { stamp[v[k]] = stamp_id; ++count; }
        }
        return count;
    }

    // Cells a tile of this region can be expected to hold.  A compact block of m cells has about
    // r_int (1 + m^(-1/3))^3 nodes per cell (interior ratio r_int, plus the faces it does not share);
    // r_int is read off the region itself with the same formula, and 8 nodes per cell is the ceiling
    // (a draped equivalent-source layer shares no vertices at all).
    static int64_t leaf_capacity(int64_t n, int64_t nodes)
    {
        const double surf_n = std::pow(1.0 + 1.0 / std::cbrt((double)n), 3.0);
        const double r_int = (double)nodes / (double)n / surf_n;
        for (int64_t m = PF_TILE_CELLS; m > 32; m -= 2) {
            const double r = std::min(8.0, r_int * std::pow(1.0 + 1.0 / std::cbrt((double)m), 3.0));
            if (m * r <= 0.97 * PF_TILE_NODES) return m;
        }
        return 32;
    }

    void run(int64_t n_cells)
    {
        std::vector<std::pair<int64_t, int64_t>> stack;
        stack.emplace_back(0, n_cells);
        while (!stack.empty()) {
            const auto [lo, hi] = stack.back();
            stack.pop_back();
            if (hi <= lo) continue;
            if (fits(lo, hi)) {
                for (int64_t i = lo; i < hi; ++i) tile_of_cell[ids[i]] = n_tiles;
                ++n_tiles;
                continue;
            }
            // axes by decreasing extent
            double ext[3];
            for (int ax = 0; ax < 3; ++ax) {
                double mn = c[ax][ids[lo]], mx = mn;
                for (int64_t i = lo; i < hi; ++i) {
                    const double v = c[ax][ids[i]];
                    mn = std::min(mn, v);
                    mx = std::max(mx, v);
                }
                ext[ax] = mx - mn;
            }
            int order[3] = {0, 1, 2};
            std::stable_sort(order, order + 3, [&](int p, int q) { return ext[p] > ext[q]; });
            // Split so that both sides are whole numbers of expected tiles (k tiles in the region:
            // floor(k / 2) of them on the low side) instead of always at the median: bisection by
            // halves leaves tiles anywhere between half full and full (308 of ~440 possible cells on
            // an octree with hanging nodes), and every tile costs the same node and cell rounds.
            const int64_t n = hi - lo;
            const int64_t cap = leaf_capacity(n, count_nodes(lo, hi));
            const int64_t k = std::max<int64_t>(2, pf_ceil_div(n, cap));
            const int64_t pos = std::min(n - 1, std::max<int64_t>(1, n * (k / 2) / k));
            int64_t m = -1;
            for (int t = 0; t < 3 && m < 0; ++t) m = cut(lo, hi, order[t], pos);
            if (m < 0) m = lo + pos;                 // all centres coincide: split by count
            stack.emplace_back(m, hi);
            stack.emplace_back(lo, m);
        }
    }
};

}  // namespace

// tile_of_cell[c] = tile of cell c; returns the number of tiles (or a negative pf_status)
extern "C" int64_t pf_plan_generic_assign(const double *nodes /*[nN][3]*/, int64_t n_nodes,
                                          const int64_t *cell_nodes /*[nC][8]*/, int64_t n_cells,
                                          int32_t *tile_of_cell)
{
    if (n_cells < 0 || n_nodes < 0 || (n_cells > 0 && (!nodes || !cell_nodes || !tile_of_cell))) {
        pf_set_error("pf_plan_generic_assign: bad arguments");
        return PF_ERR_INVALID;
    }
    if (n_cells >= ((int64_t)1 << 31) || n_nodes >= ((int64_t)1 << 31)) {
        pf_set_error("pf_plan_generic_assign: more than 2^31-1 cells or nodes");
        return PF_ERR_INVALID;
    }
    RcbPlanner r;
    r.cn = cell_nodes;
    r.n_nodes = n_nodes;
    r.tile_of_cell = tile_of_cell;
    for (int ax = 0; ax < 3; ++ax) r.c[ax].resize((size_t)n_cells);
    r.ids.resize((size_t)n_cells);
    r.stamp.assign((size_t)n_nodes, 0);
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(4) reduction(| : bad)
    for (int64_t cidx = 0; cidx < n_cells; ++cidx) {
        const int64_t *v = cell_nodes + 8 * cidx;
        double s[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < 8; ++k) {
            if (v[k] < 0 || v[k] >= n_nodes) { bad = 1; continue; }
            for (int ax = 0; ax < 3; ++ax) s[ax] += nodes[3 * v[k] + ax];
        }
        for (int ax = 0; ax < 3; ++ax) r.c[ax][cidx] = 0.125 * s[ax];
        r.ids[cidx] = (int32_t)cidx;
    }
    if (bad) {
        pf_set_error("pf_plan_generic_assign: cell_nodes holds indices outside the nodes array");
        return PF_ERR_INVALID;
    }
    r.run(n_cells);
    return r.n_tiles;
}

// Fills the pf_plan arrays for a given assignment: cells of a tile in ascending index order
// (contiguous G writes), local node ids in ascending node order.
extern "C" int pf_plan_generic_fill(const double *nodes, int64_t n_nodes, const int64_t *cell_nodes,
                                    int64_t n_cells, const int32_t *tile_of_cell, int64_t n_tiles,
                                    double *tile_xyz, uint16_t *tile_cell_nodes, int32_t *tile_cell_col,
                                    int32_t *tile_n_nodes)
{
    PF_CHECK_ARG(n_cells >= 0 && n_tiles >= 0, "negative size");
    if (n_cells == 0 || n_tiles == 0) return PF_OK;
    PF_CHECK_ARG(nodes && cell_nodes && tile_of_cell && tile_xyz && tile_cell_nodes && tile_cell_col && tile_n_nodes,
                 "null pointer");
    std::vector<int64_t> start((size_t)n_tiles + 1, 0);
    for (int64_t cidx = 0; cidx < n_cells; ++cidx) {
        const int32_t t = tile_of_cell[cidx];
        PF_CHECK_ARG(t >= 0 && t < n_tiles, "tile id out of range");
        ++start[(size_t)t + 1];
    }
    for (int64_t t = 0; t < n_tiles; ++t) {
        PF_CHECK_ARG(start[(size_t)t + 1] <= PF_TILE_CELLS, "tile with too many cells");
        start[(size_t)t + 1] += start[(size_t)t];
    }
    std::vector<int32_t> cells((size_t)n_cells);
    {
        std::vector<int64_t> pos(start.begin(), start.end() - 1);
        for (int64_t cidx = 0; cidx < n_cells; ++cidx) cells[(size_t)pos[(size_t)tile_of_cell[cidx]]++] = (int32_t)cidx;
    }
    int failed = 0;
#pragma omp parallel num_threads(4)
    {
        std::vector<int32_t> local((size_t)n_nodes, -1);      // per thread: global node -> local id
        std::vector<int64_t> uniq;
        uniq.reserve(PF_TILE_NODES + 8);
#pragma omp for schedule(dynamic, 16)
        for (int64_t t = 0; t < n_tiles; ++t) {
            uniq.clear();
            for (int64_t i = start[(size_t)t]; i < start[(size_t)t + 1]; ++i) {
                const int64_t *v = cell_nodes + 8 * (int64_t)cells[(size_t)i];
                for (int k = 0; k < 8; ++k)
                    if (local[(size_t)v[k]] < 0) { local[(size_t)v[k]] = 0; uniq.push_back(v[k]); }
            }
            if ((int64_t)uniq.size() > PF_TILE_NODES) {
                failed = 1;
                for (int64_t g : uniq) local[(size_t)g] = -1;
                continue;
            }
            std::sort(uniq.begin(), uniq.end());
            double *xyz = tile_xyz + t * 3 * (int64_t)PF_TILE_NODES;
            for (size_t j = 0; j < uniq.size(); ++j) {
                local[(size_t)uniq[j]] = (int32_t)j;
                for (int ax = 0; ax < 3; ++ax) xyz[ax * PF_TILE_NODES + (int64_t)j] = nodes[3 * uniq[j] + ax];
            }
            for (size_t j = uniq.size(); j < (size_t)PF_TILE_NODES; ++j)
                for (int ax = 0; ax < 3; ++ax) xyz[ax * PF_TILE_NODES + (int64_t)j] = 0.0;
            tile_n_nodes[t] = (int32_t)uniq.size();
            int64_t slot = 0;
            for (int64_t i = start[(size_t)t]; i < start[(size_t)t + 1]; ++i, ++slot) {
                const int64_t cidx = cells[(size_t)i];
                const int64_t *v = cell_nodes + 8 * cidx;
                uint16_t *out = tile_cell_nodes + (t * PF_TILE_CELLS + slot) * 8;
                for (int k = 0; k < 8; ++k) out[k] = (uint16_t)local[(size_t)v[k]];
                tile_cell_col[t * PF_TILE_CELLS + slot] = (int32_t)cidx;
            }
            for (; slot < PF_TILE_CELLS; ++slot) {
                uint16_t *out = tile_cell_nodes + (t * PF_TILE_CELLS + slot) * 8;
                for (int k = 0; k < 8; ++k) out[k] = 0;
                tile_cell_col[t * PF_TILE_CELLS + slot] = -1;
            }
            for (int64_t g : uniq) local[(size_t)g] = -1;
        }
    }
    PF_CHECK_ARG(!failed, "tile with too many nodes");
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
