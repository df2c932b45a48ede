// pf_abi.cu -- ABI version and per-thread error reporting of libpf_cuda.so.
#include <atomic>
#include <cstdarg>

#include "pf_common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void pf_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

std::atomic<int64_t> g_pf_launches{0};

extern "C" int64_t pf_launch_count(void) { return g_pf_launches.load(); }

extern "C" int pf_abi_version(void) { return PF_ABI_VERSION; }

extern "C" const char *pf_last_error(void) { return g_error; }

// SMs of the current device (cached per device; 148 on B200).  Grid sizes and workspace
// queries are derived from it; without a usable device (the CPU-side ABI tests) the B200
// count is returned so that the size queries stay callable.
int pf_sm_count(void)
{
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return 148;
    }
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;
    }
    cache[dev].store(n, std::memory_order_relaxed);
    return n;
}
