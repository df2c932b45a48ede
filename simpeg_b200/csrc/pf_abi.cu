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
