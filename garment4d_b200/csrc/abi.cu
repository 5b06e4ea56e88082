// C-ABI plumbing shared by all entry points: error text, version, device queries.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace g4d {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace g4d

G4D_API const char* g4d_last_error(void) { return g4d::g_err; }
G4D_API int g4d_abi_version(void) { return 1; }
G4D_API int g4d_sm_count(void) { return g4d::sm_count(); }
G4D_API unsigned long long g4d_launch_count(void) { return __atomic_load_n(&g4d::g_launches, __ATOMIC_RELAXED); }
