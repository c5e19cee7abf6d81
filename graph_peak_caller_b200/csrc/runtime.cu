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

// LinearMap.from_graph / find_starts / find_ends (control/linearmap.py:102-154)
// on HOST arrays: longest distance from the source to every node start, and
// linear length minus the longest distance from every node end to the sink.
// Sequential by nature (each node depends on its predecessors); one pass each
// way over the CSR.  Preprocessing step, all pointers are host pointers.
extern "C" int gpc_linear_map_from_graph_host(const int64_t* node_len, const int64_t* succ_ptr,
                                              const int32_t* succ, const int64_t* pred_ptr,
                                              const int32_t* pred, uint64_t n_nodes,
                                              double* starts, double* ends) {
    if (n_nodes == 0) return GPC_OK;
    for (uint64_t v = 0; v < n_nodes; ++v) starts[v] = 0.0;
    for (uint64_t v = 0; v < n_nodes; ++v) {
        double reach = starts[v] + (double)node_len[v];
        for (int64_t e = succ_ptr[v]; e < succ_ptr[v + 1]; ++e)
            if (reach > starts[succ[e]]) starts[succ[e]] = reach;
    }
    for (uint64_t v = 0; v < n_nodes; ++v) ends[v] = 0.0;       // used as "back" first
    for (uint64_t v = n_nodes; v-- > 0;) {
        double reach = ends[v] + (double)node_len[v];
        for (int64_t e = pred_ptr[v]; e < pred_ptr[v + 1]; ++e)
            if (reach > ends[pred[e]]) ends[pred[e]] = reach;
    }
    double length = ends[0] + (double)node_len[0];               // linearmap.py:153
    for (uint64_t v = 0; v < n_nodes; ++v) ends[v] = length - ends[v];
    return GPC_OK;
}
