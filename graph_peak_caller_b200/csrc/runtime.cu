// Error text, version string, device properties.
#include <stdarg.h>

#include "common.cuh"

namespace gpc {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace gpc

extern "C" const char* gpc_last_error_string(void) { return gpc::g_error; }
extern "C" const char* gpc_version_string(void) { return "gpc_b200 0.1 sm_100a"; }
