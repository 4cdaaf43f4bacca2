// Error plumbing, launch counter and version for the C ABI (include/mofa_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// called right after every kernel launch: counts it and turns launch-time errors into return codes
int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return MOFA_ERR_CUDA;
    }
    return MOFA_OK;
}

}  // namespace mofa

extern "C" const char* mofa_last_error(void) { return mofa::g_err; }
extern "C" int mofa_version(void) { return 100; }
extern "C" int64_t mofa_launch_count(void) { return mofa::g_launches.load(); }
extern "C" void mofa_launch_count_reset(void) { mofa::g_launches.store(0); }
