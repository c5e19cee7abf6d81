// Shared device/host helpers for the sm_100a callpeaks kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/gpc_b200.h"

namespace gpc {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);

#define GPC_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            gpc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,       \
                           cudaGetErrorString(e__));                           \
            return GPC_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

#define GPC_TRY(call)                                                          \
    do {                                                                       \
        int s__ = (call);                                                      \
        if (s__ != GPC_OK) return s__;                                         \
    } while (0)

// Every kernel launch is followed by GPC_LAUNCH_CHECK(); a launch preceded by
// GPC_PROF("name") is timed with a CUDA event pair on its own stream while
// profiling is enabled (gpc_profile_enable), and counted always.
void prof_begin(const char* name, cudaStream_t stream, double algorithmic_bytes = 0.0);
void prof_end();
void prof_add_bytes(const char* name, double algorithmic_bytes);
#define GPC_PROF(name) gpc::prof_begin(name, stream)
// same, with the ALGORITHMIC bytes of this launch (compulsory input + output,
// DESIGN.md "roofline accounting") for the achieved-bandwidth figure
#define GPC_PROF_BYTES(name, bytes) gpc::prof_begin(name, stream, (double)(bytes))
#define GPC_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        gpc::prof_end();                                                       \
        GPC_CUDA(cudaGetLastError());                                          \
    } while (0)

int sm_count();

// One NVTX range per C-ABI stage (tracing): a timeline tool (nsys, ncu
// --nvtx) shows the kernels of a pass grouped under gpc_unique_reads, gpc_sample_pileup,
// gpc_control_linear_track, ...  Header-only NVTX v3: without a tool attached a push is a
// load and a branch.
struct StageRange {
    explicit StageRange(const char* name) { nvtxRangePushA(name); }
    ~StageRange() { nvtxRangePop(); }
    StageRange(const StageRange&) = delete;
    StageRange& operator=(const StageRange&) = delete;
};
#define GPC_STAGE() gpc::StageRange stage_range__(__func__)

// Small device -> host read-backs (counts, flags) without cudaMemcpy +
// cudaStreamSynchronize: a one-thread kernel copies the words into a
// host-mapped slot and then bumps a sequence number the host polls.  Measured
// on B200: ~10 us from kernel end to the next launch instead of ~24 us.
// At most 4 items, 160 bytes in total.  Returns after the values have arrived,
// i.e. everything enqueued on `stream` before the call has finished.
struct RbItem { const void* dev; void* host; unsigned bytes; };
int read_back(const RbItem* items, int n_items, cudaStream_t stream);
#define GPC_READ_BACK(...)                                                     \
    do {                                                                       \
        const gpc::RbItem items__[] = {__VA_ARGS__};                           \
        GPC_TRY(gpc::read_back(items__, (int)(sizeof(items__) / sizeof(items__[0])), stream)); \
    } while (0)

// Stream-ordered scratch that frees itself (cudaMallocAsync pool).
//
// Guard mode (environment GPC_GUARD=1; compute-sanitizer is closed on the GPU pool this was
// developed on): every scratch buffer is poisoned with 0xA5 when it is handed out -- a kernel
// that reads what nobody wrote then computes garbage and the parity tests see it -- and sits
// between two 256-byte canaries that are checked when it is given back; a kernel that wrote
// outside its buffer is counted in gpc_guard_violations() and named on stderr.
bool guard_mode();
void guard_check(const void* base, size_t bytes, cudaStream_t stream);
constexpr size_t GUARD_PAD = 256;

struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    size_t guarded_bytes = 0;          // > 0: guard mode, p = base + GUARD_PAD
    Scratch() {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { release(); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        release();
        sm_count();   // first call configures the memory pool of this device
        s = stream;
        if (bytes == 0) bytes = 16;
        if (guard_mode()) {
            const size_t padded = ((bytes + 15) & ~(size_t)15) + 2 * GUARD_PAD;
            void* base = nullptr;
            cudaError_t e = cudaMallocAsync(&base, padded, stream);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(base, 0xA5, padded, stream);
            if (e != cudaSuccess) return e;
            p = static_cast<char*>(base) + GUARD_PAD;
            guarded_bytes = bytes;
            return cudaSuccess;
        }
        return cudaMallocAsync(&p, bytes, stream);
    }
    void release() {
        if (p && guarded_bytes) {
            void* base = static_cast<char*>(p) - GUARD_PAD;
            guard_check(base, guarded_bytes, s);
            cudaFreeAsync(base, s);
        } else if (p) {
            cudaFreeAsync(p, s);
        }
        p = nullptr;
        guarded_bytes = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- warp / block primitives ------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_incl_sum(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(FULL, v, d);
        if (lane_id() >= (unsigned)d) v += o;
    }
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Block-wide exclusive sum.  `smem` needs 33 T's.  Returns the exclusive
// prefix of `v` for this thread; `total` receives the block sum.  Contains
// __syncthreads, so every thread of the block must call it.
template <typename T>
__device__ __forceinline__ T block_excl_sum(T v, T* smem, T& total) {
    T incl = warp_incl_sum(v);
    unsigned w = warp_id(), l = lane_id();
    unsigned nw = (blockDim.x + 31) >> 5;
    if (l == 31) smem[w] = incl;
    __syncthreads();
    if (w == 0) {
        T x = (l < nw) ? smem[l] : T(0);
        T xi = warp_incl_sum(x);
        smem[l] = xi - x;
        if (l == 31) smem[32] = xi;
    }
    __syncthreads();
    T base = smem[w];
    total = smem[32];
    __syncthreads();
    return base + incl - v;
}

// ---- chained scan (decoupled look-back) over tiles --------------------------
// One 64-bit word per tile: status in the top 2 bits, 62-bit payload below.
// Tiles must be claimed in launch order (atomic ticket) so a tile only ever
// waits on tiles that are already running.
constexpr unsigned long long LB_EMPTY = 0ull;
constexpr unsigned long long LB_AGG = 1ull << 62;
constexpr unsigned long long LB_PREFIX = 2ull << 62;
constexpr unsigned long long LB_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long lb_load(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void lb_store(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// The walk back is warp-wide (32 predecessors per step) and comes in two halves,
// so that a block can put other work (the
// loads of its next tile) between publishing a tile's aggregate and needing
// its prefix: by then the predecessors have published too and the walk back
// costs one or two round trips instead of a spin.
// lb_publish_agg: ONE thread.  lb_resolve_warp: all 32 lanes of ONE warp, same
// arguments; 32 x GPC_LB_WIDE predecessors are inspected per step.
__device__ __forceinline__ void lb_publish_agg(unsigned long long* status, unsigned tile,
                                               unsigned long long agg) {
    lb_store(&status[tile], (tile == 0 ? LB_PREFIX : LB_AGG) | (agg & LB_MASK));
}
// LB_WIDE windows of 32 predecessors are loaded per round trip (the loads of a round are
// independent: one L2 latency for all of them).  With ~600 tiles in flight that arrive at
// their look-back as a wave, the tile at position i of the wave has to walk back over i
// aggregates before it meets a prefix; at 32 per round trip that was up to 18 dependent L2
// round trips (ncu of the difference-array scan: barrier 13 + long scoreboard 12 warps per
// issue, nothing busy above 41 %), at 128 it is five.
#ifndef GPC_LB_WIDE
#define GPC_LB_WIDE 4
#endif
__device__ __forceinline__ unsigned long long lb_resolve_warp(unsigned long long* status,
                                                              unsigned tile,
                                                              unsigned long long agg) {
    constexpr int LB_WIDE = GPC_LB_WIDE;
    const unsigned lane = threadIdx.x & 31;
    agg &= LB_MASK;
    if (tile == 0) return 0;
    unsigned long long excl = 0;
    int newest = (int)tile - 1;
    while (true) {
        unsigned long long s[LB_WIDE];
#pragma unroll
        for (int k = 0; k < LB_WIDE; ++k) {
            const int t = newest - (int)lane - 32 * k;     // lane 0 of window 0 = nearest predecessor
            s[k] = (t >= 0) ? lb_load(&status[t]) : LB_PREFIX;
        }
        bool done = false;
        int complete = 0;                           // windows of this round that were all ready
#pragma unroll
        for (int k = 0; k < LB_WIDE; ++k) {
            if (done || complete != k) continue;    // (uniform: every lane sees the same ballots)
            const unsigned long long st = s[k] & ~LB_MASK;
            const unsigned empty = __ballot_sync(0xffffffffu, st == LB_EMPTY);
            const unsigned pref = __ballot_sync(0xffffffffu, st == LB_PREFIX);
            const int first_pref = pref ? (__ffs(pref) - 1) : 32;
            const unsigned need = (first_pref >= 31) ? 0xffffffffu : ((2u << first_pref) - 1u);
            if (empty & need) continue;             // someone in this window is not ready: reload from here
            unsigned long long v = ((int)lane <= first_pref) ? (s[k] & LB_MASK) : 0ull;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            excl += v;
            complete = k + 1;
            if (first_pref < 32) done = true;
        }
        if (done) break;
        newest -= 32 * complete;                    // aggregates already summed stay summed
    }
    // sums are taken modulo 2^62, so packed signed payloads work as well
    if (lane == 0) lb_store(&status[tile], LB_PREFIX | ((excl + agg) & LB_MASK));
    return excl & LB_MASK;
}
// Both halves back to back (all 32 lanes of ONE warp, same arguments).
__device__ __forceinline__ unsigned long long lookback_excl_warp(unsigned long long* status,
                                                                 unsigned tile,
                                                                 unsigned long long agg) {
    if ((threadIdx.x & 31) == 0) lb_publish_agg(status, tile, agg);
    return lb_resolve_warp(status, tile, agg);
}

__device__ __forceinline__ uint4 node_record(const gpc_graph_t& g, uint32_t v) {
    return __ldg(reinterpret_cast<const uint4*>(g.node_rec) + v);
}

// Order-preserving maps between IEEE doubles and unsigned 64-bit keys.
__host__ __device__ __forceinline__ unsigned long long f64_to_key(double d) {
    unsigned long long u;
#ifdef __CUDA_ARCH__
    u = (unsigned long long)__double_as_longlong(d);
#else
    memcpy(&u, &d, 8);
#endif
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_to_f64(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}

}  // namespace gpc
