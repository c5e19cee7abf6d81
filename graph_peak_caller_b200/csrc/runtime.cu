// Error text, version string, device properties.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace gpc {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---- launch counter and per-kernel event timing -------------------------------
struct ProfRec { const char* name; cudaEvent_t a, b; double bytes; };
static std::map<std::string, double> g_extra_bytes;
static std::vector<ProfRec> g_recs;
static bool g_prof_on = false;
static std::atomic<unsigned long long> g_launches{0};
static thread_local const char* g_pending = nullptr;
static thread_local cudaStream_t g_pending_stream = nullptr;
static thread_local cudaEvent_t g_pending_a = nullptr;
static thread_local double g_pending_bytes = 0.0;
static std::mutex g_prof_mutex;        // stage functions may run on several host threads

void prof_add_bytes(const char* name, double bytes) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_extra_bytes[name] += bytes;
}

void prof_begin(const char* name, cudaStream_t stream, double bytes) {
    if (!g_prof_on) return;
    g_pending = name;
    g_pending_bytes = bytes;
    g_pending_stream = stream;
    cudaEventCreate(&g_pending_a);
    cudaEventRecord(g_pending_a, stream);
}

void prof_end() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_pending) return;
    ProfRec r{g_pending, g_pending_a, nullptr, g_pending_bytes};
    cudaEventCreate(&r.b);
    cudaEventRecord(r.b, g_pending_stream);
    {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        g_recs.push_back(r);
    }
    g_pending = nullptr;
}

// The stream-ordered pool gives freed scratch back to the OS at every
// synchronisation unless told otherwise; keep it (the path allocates and frees
// the same few hundred MB every pass).
static void keep_pool_memory(int dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        keep_pool_memory(dev);
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

extern "C" void gpc_profile_enable(int on) { gpc::g_prof_on = on != 0; }
extern "C" uint64_t gpc_launch_count(void) { return gpc::g_launches.load(); }

// Writes "name calls total_ms algorithmic_bytes\n" lines for every kernel timed since the last
// report (synchronises the device) and clears the records.  Returns the
// number of bytes needed (excluding the terminating 0).
extern "C" uint64_t gpc_profile_report(char* buf, uint64_t capacity) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(gpc::g_prof_mutex);
    struct Acc { unsigned long long calls = 0; double ms = 0, bytes = 0; };
    std::map<std::string, Acc> acc;
    for (auto& r : gpc::g_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = acc[r.name];
        e.calls += 1;
        e.ms += ms;
        e.bytes += r.bytes;
    }
    for (auto& kv : gpc::g_extra_bytes) acc[kv.first].bytes += kv.second;
    if (buf && capacity) {       // a size query (buf == NULL) keeps the records
        for (auto& r : gpc::g_recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        gpc::g_recs.clear();
        gpc::g_extra_bytes.clear();
    }
    std::string out;
    char line[256];
    for (auto& kv : acc) {
        snprintf(line, sizeof(line), "%s %llu %.6f %.0f\n", kv.first.c_str(), kv.second.calls,
                 kv.second.ms, kv.second.bytes);
        out += line;
    }
    if (buf && capacity) {
        size_t nbytes = out.size() < capacity - 1 ? out.size() : capacity - 1;
        memcpy(buf, out.data(), nbytes);
        buf[nbytes] = 0;
    }
    return out.size();
}
