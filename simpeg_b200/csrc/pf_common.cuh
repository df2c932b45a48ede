// pf_common.cuh -- shared helpers of libpf_cuda.so (error reporting, constants).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "pf_cuda.h"

// Tile geometry is fixed at compile time (the host planner must match; it is
// validated against pf_plan.tile_cells / tile_nodes on every call).
constexpr int PF_THREADS = 256;
constexpr int PF_TILE_CELLS = 512;  // 2 cells per thread
constexpr int PF_TILE_NODES = 768;  // 3 nodes per thread (8x8x8 brick = 729 nodes)
constexpr int PF_CPT = PF_TILE_CELLS / PF_THREADS;
constexpr int PF_NPT = PF_TILE_NODES / PF_THREADS;

// scipy.constants.G (CODATA 2018/2022): the reference multiplies it by the
// unit conversion factor, gravity/simulation.py:101-111, 580.
constexpr double PF_NEWTON_G = 6.6743e-11;

void pf_set_error(const char *fmt, ...);
int pf_sm_count(void);   // SMs of the current device (pf_abi.cu)
extern std::atomic<int64_t> g_pf_launches;
#define PF_COUNT_LAUNCH() (g_pf_launches.fetch_add(1, std::memory_order_relaxed))

#define PF_CHECK_ARG(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            pf_set_error("%s: %s", __func__, msg);                \
            return PF_ERR_INVALID;                                \
        }                                                         \
    } while (0)

#define PF_CUDA_OK(expr)                                                        \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            pf_set_error("%s: CUDA error %s at %s:%d", __func__,                \
                         cudaGetErrorString(_e), __FILE__, __LINE__);           \
            return PF_ERR_CUDA;                                                 \
        }                                                                       \
    } while (0)

static inline int64_t pf_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
