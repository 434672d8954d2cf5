// Error plumbing and bookkeeping shared by every libstv entry point.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "stv_common.cuh"

namespace stv {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return STV_E_CUDA;
    }
    return STV_OK;
}

}  // namespace stv

extern "C" int stv_version(void) { return 100; }
extern "C" const char* stv_last_error(void) { return stv::g_err; }
extern "C" unsigned long long stv_launch_count(void) { return stv::g_launches.load(std::memory_order_relaxed); }
