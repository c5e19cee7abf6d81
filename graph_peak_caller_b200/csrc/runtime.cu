// Error text, version string, device properties.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

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

// events are recycled: cudaEventCreate costs more than the record itself
static std::vector<cudaEvent_t> g_free_events;
static cudaEvent_t take_event() {
    {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        if (!g_free_events.empty()) {
            cudaEvent_t e = g_free_events.back();
            g_free_events.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(const char* name, cudaStream_t stream, double bytes) {
    if (!g_prof_on) return;
    g_pending = name;
    g_pending_bytes = bytes;
    g_pending_stream = stream;
    g_pending_a = take_event();
    cudaEventRecord(g_pending_a, stream);
}

void prof_end() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_pending) return;
    ProfRec r{g_pending, g_pending_a, take_event(), g_pending_bytes};
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

// Per device (a process normally drives one GPU, but one that switches devices gets each
// device's own SM count, and each device's memory pool is configured on first use).
int sm_count() {
    constexpr int MAX_DEVICES = 64;
    static std::atomic<int> known[MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148;
    int n = known[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        keep_pool_memory(dev);
        known[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- guard mode ----------------------------------------------------------------
static std::atomic<unsigned long long> g_guard_violations{0};

bool guard_mode() {
    static const bool on = [] { const char* e = getenv("GPC_GUARD"); return e && atoi(e) != 0; }();
    return on;
}

// base: start of the padded allocation; bytes: the size the caller asked for
void guard_check(const void* base, size_t bytes, cudaStream_t stream) {
    const size_t body = (bytes + 15) & ~(size_t)15;
    unsigned char head[GUARD_PAD], tail[GUARD_PAD + 16];
    const char* b = static_cast<const char*>(base);
    // the unused tail of the last 16 bytes belongs to the canary as well
    const size_t slack = body - bytes;
    if (cudaMemcpyAsync(head, b, GUARD_PAD, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaMemcpyAsync(tail, b + GUARD_PAD + bytes, GUARD_PAD + slack, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
        cudaGetLastError();
        return;                        // a faulted stream is reported by whoever launched on it
    }
    size_t bad_head = 0, bad_tail = 0;
    for (size_t i = 0; i < GUARD_PAD; ++i) bad_head += head[i] != 0xA5;
    for (size_t i = 0; i < GUARD_PAD + slack; ++i) bad_tail += tail[i] != 0xA5;
    if (bad_head || bad_tail) {
        g_guard_violations.fetch_add(1);
        fprintf(stderr, "gpc_b200 GUARD: scratch of %zu bytes written outside its bounds "
                        "(%zu bytes before, %zu bytes behind)\n", bytes, bad_head, bad_tail);
    }
}

// ---- read_back ---------------------------------------------------------------
struct RbArgs { const unsigned char* src[4]; unsigned bytes[4]; int n; };
constexpr unsigned RB_PAYLOAD = 160;

// One thread, byte by byte.  (Measured and dropped: one thread per payload byte with a block
// barrier before the sequence number -- all loads in flight together, a warp's stores leaving
// as one write -- took LONGER, 10.6 instead of 8.7 us per read-back: the time is the
// system-scope fence and the posted write reaching the host, not the copy loop.)
__global__ void readback_kernel(RbArgs a, volatile unsigned char* slot, unsigned long long seq) {
    unsigned off = 0;
    for (int i = 0; i < a.n; ++i) {
        for (unsigned b = 0; b < a.bytes[i]; ++b) slot[off + b] = a.src[i][b];
        off += (a.bytes[i] + 7u) & ~7u;
    }
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(slot + RB_PAYLOAD) = seq;
}

struct RbSlot {
    unsigned char* host = nullptr;
    unsigned char* dev = nullptr;
    unsigned long long seq = 0;
};
static thread_local RbSlot g_rb;

int read_back(const RbItem* items, int n_items, cudaStream_t stream) {
    if (n_items <= 0) return GPC_OK;
    if (!g_rb.host) {
        GPC_CUDA(cudaHostAlloc((void**)&g_rb.host, 256, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(g_rb.host, 0, 256);
        GPC_CUDA(cudaHostGetDevicePointer((void**)&g_rb.dev, g_rb.host, 0));
    }
    RbArgs a;
    unsigned total = 0;
    a.n = n_items;
    if (n_items > 4) { set_error("read_back: too many items"); return GPC_ERR_INTERNAL; }
    for (int i = 0; i < 4; ++i) {
        a.src[i] = i < n_items ? static_cast<const unsigned char*>(items[i].dev) : nullptr;
        a.bytes[i] = i < n_items ? items[i].bytes : 0;
        total += (a.bytes[i] + 7u) & ~7u;
    }
    if (total > RB_PAYLOAD) { set_error("read_back: payload too large"); return GPC_ERR_INTERNAL; }
    const unsigned long long seq = ++g_rb.seq;
    prof_begin("readback", stream);
    readback_kernel<<<1, 1, 0, stream>>>(a, g_rb.dev, seq);
    GPC_LAUNCH_CHECK();
    volatile unsigned long long* flag =
        reinterpret_cast<volatile unsigned long long*>(g_rb.host + RB_PAYLOAD);
    for (unsigned spins = 1; *flag != seq; ++spins) {
        if ((spins & 0x3fff) == 0) {              // a faulting kernel must not hang the host
            cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaErrorNotReady) continue;
            GPC_CUDA(q);
            if (*flag != seq) { set_error("read_back: value never arrived"); return GPC_ERR_INTERNAL; }
        }
    }
    __sync_synchronize();
    unsigned off = 0;
    for (int i = 0; i < n_items; ++i) {
        memcpy(items[i].host, g_rb.host + off, items[i].bytes);
        off += (items[i].bytes + 7u) & ~7u;
    }
    return GPC_OK;
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

extern "C" uint64_t gpc_guard_violations(void) { return gpc::g_guard_violations.load(); }
extern "C" int gpc_guard_mode(void) { return gpc::guard_mode() ? 1 : 0; }

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
    // debugging aid: GPC_PROF_TIMELINE=<file> gets "name start_ms duration_ms" per launch
    if (const char* path = getenv("GPC_PROF_TIMELINE")) {
        if (buf && capacity && !gpc::g_recs.empty()) {
            if (FILE* fh = fopen(path, "w")) {
                for (auto& r : gpc::g_recs) {
                    float t0 = 0.f, d = 0.f;
                    cudaEventElapsedTime(&t0, gpc::g_recs[0].a, r.a);
                    cudaEventElapsedTime(&d, r.a, r.b);
                    fprintf(fh, "%s %.6f %.6f\n", r.name, t0, d);
                }
                fclose(fh);
            }
        }
    }
    if (buf && capacity) {       // a size query (buf == NULL) keeps the records
        for (auto& r : gpc::g_recs) {
            gpc::g_free_events.push_back(r.a);
            gpc::g_free_events.push_back(r.b);
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
