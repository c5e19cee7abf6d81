// Device-wide exclusive prefix sum (single pass, chained scan) and the stream
// compaction helpers built on it.
#pragma once
#include "common.cuh"

namespace gpc {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

// Claims the next tile index for this block (launch-order ticket).
__device__ __forceinline__ unsigned claim_tile(unsigned* ticket, unsigned* s_slot) {
    if (threadIdx.x == 0) *s_slot = atomicAdd(ticket, 1u);
    __syncthreads();
    unsigned t = *s_slot;
    __syncthreads();
    return t;
}

// N (a multiple of 4) consecutive uint32 items of one thread as 16-byte accesses where the
// array allows it (16-byte aligned base, whole group inside the array), else item by item.
template <int N>
__device__ __forceinline__ void load_items_u32(const uint32_t* __restrict__ in, unsigned base, unsigned n,
                                               unsigned (&v)[N]) {
    if (base + N <= n && (reinterpret_cast<size_t>(in) & 15) == 0) {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const uint4 a = p[q];
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = (base + j < n) ? in[base + j] : 0u;
    }
}
// the same for uint8 items (N a multiple of 8)
template <int N>
__device__ __forceinline__ void load_items_u8(const uint8_t* __restrict__ in, unsigned base, unsigned n,
                                              unsigned (&v)[N]) {
    if (base + N <= n && (reinterpret_cast<size_t>(in) & 15) == 0) {
        const uint2* p = reinterpret_cast<const uint2*>(in + base);
#pragma unroll
        for (int q = 0; q < N / 8; ++q) {
            const uint2 a = p[q];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * q + j] = (a.x >> (8 * j)) & 0xffu;
                v[8 * q + 4 + j] = (a.y >> (8 * j)) & 0xffu;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = (base + j < n) ? (unsigned)in[base + j] : 0u;
    }
}
// out[base + j] = run + v[0] + ... + v[j - 1]
template <int N>
__device__ __forceinline__ void store_running(uint32_t* __restrict__ out, unsigned base, unsigned n,
                                              unsigned run, const unsigned (&v)[N]) {
    if (base + N <= n && (reinterpret_cast<size_t>(out) & 15) == 0) {
        uint4* p = reinterpret_cast<uint4*>(out + base);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            uint4 o;
            o.x = run; run += v[4 * q];
            o.y = run; run += v[4 * q + 1];
            o.z = run; run += v[4 * q + 2];
            o.w = run; run += v[4 * q + 3];
            p[q] = o;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
}

// ---- reduce, then scan ------------------------------------------------------------------
// The chained scans above make every tile wait for the tiles before it; with several hundred
// tiles in flight that wait was 45 % of the difference-array scan (measured with the look-back
// switched off, csrc/sample.cu) and the top stall of every compaction (barrier 35-42 warps per
// issue).  Above RTS_MIN_ITEMS items the drivers below therefore run two launches instead:
// per-tile totals (one more read of the input), and the scan / compaction kernel, whose
// tiles look their prefix up in the totals -- tiles independent.
constexpr size_t RTS_MIN_ITEMS = 1u << 20;

// Tile totals come in two levels: sums[tile], and supers[tile / 64] accumulated by one atomic
// per tile.  A tile's exclusive prefix is then the supers before its group plus the tiles
// before it inside the group: at most a few hundred independent loads spread over the block
// and one block reduction -- no serial pass over the totals, no launch of its own (a single
// block scanning 13 100 totals took 25-45 us, as a kernel or as the tail of the last block).
constexpr int RTS_GROUP_LOG2 = 6;
__device__ __forceinline__ void tile_total_publish(unsigned long long* sums, unsigned long long* supers,
                                                   unsigned tile, unsigned long long t) {   // ONE thread
    sums[tile] = t;
    atomicAdd(supers + (tile >> RTS_GROUP_LOG2), t);
}
// every thread of the block calls it (contains barriers); s_red: 33 entries
__device__ __forceinline__ unsigned long long tile_prefix_lookup(const unsigned long long* __restrict__ sums,
                                                                 const unsigned long long* __restrict__ supers,
                                                                 unsigned tile, unsigned long long* s_red) {
    const unsigned group = tile >> RTS_GROUP_LOG2, first = group << RTS_GROUP_LOG2;
    unsigned long long v = 0;
    for (unsigned i = threadIdx.x; i < group; i += blockDim.x) v += supers[i];
    for (unsigned i = first + threadIdx.x; i < tile; i += blockDim.x) v += sums[i];
    v = warp_sum(v);
    if (lane_id() == 0) s_red[warp_id()] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
        s_red[32] = t;
    }
    __syncthreads();
    const unsigned long long r = s_red[32];
    __syncthreads();
    return r;
}

// out[i] = sum_{j<i} in[j]  (uint32 in, uint64 running sum stored as uint32:
// callers guarantee the total fits).  total[0] receives the grand total.
// Sixteen items per thread in 16-byte accesses: eight scalar loads at a stride of eight
// items replayed every sector eight times, and 2048-item tiles left the per-tile costs
// (ticket, two block scans, look-back) uncovered -- node-sized scans ran at a tenth of
// the copy rate.
constexpr int ES_ITEMS = 16;
constexpr int ES_TILE = SC_THREADS * ES_ITEMS;

template <typename InT>
__global__ void __launch_bounds__(SC_THREADS)
excl_scan_sums_kernel(const InT* __restrict__ in, unsigned n, unsigned long long* __restrict__ sums,
                      unsigned long long* __restrict__ supers /* zeroed */) {
    __shared__ unsigned s_red[SC_THREADS / 32];
    const unsigned base = blockIdx.x * ES_TILE + threadIdx.x * ES_ITEMS;
    unsigned v[ES_ITEMS], sum = 0;
    if (sizeof(InT) == 4)
        load_items_u32<ES_ITEMS>(reinterpret_cast<const uint32_t*>(in), base, n, v);
    else
        load_items_u8<ES_ITEMS>(reinterpret_cast<const uint8_t*>(in), base, n, v);
#pragma unroll
    for (int j = 0; j < ES_ITEMS; ++j) sum += v[j];
    sum = warp_sum(sum);
    if (lane_id() == 0) s_red[warp_id()] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < SC_THREADS / 32; ++w) t += s_red[w];
        tile_total_publish(sums, supers, blockIdx.x, t);
    }
}

// tile_prefix == NULL: chained scan over the tiles (status, ticket); else the exclusive
// prefix of every tile is handed in and the tiles are independent
template <typename InT>
__global__ void __launch_bounds__(SC_THREADS)
excl_scan_kernel(const InT* __restrict__ in, uint32_t* __restrict__ out, unsigned n,
                 unsigned long long* __restrict__ status, unsigned* __restrict__ ticket,
                 unsigned long long* __restrict__ total,
                 const unsigned long long* __restrict__ tile_sums = nullptr,
                 const unsigned long long* __restrict__ tile_supers = nullptr) {
    __shared__ unsigned s_scan[33];
    __shared__ unsigned s_slot;
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_red[33];
    const unsigned tile = tile_sums ? blockIdx.x : claim_tile(ticket, &s_slot);
    const unsigned base = tile * ES_TILE + threadIdx.x * ES_ITEMS;
    unsigned v[ES_ITEMS];
    unsigned sum = 0;
    if (sizeof(InT) == 4)
        load_items_u32<ES_ITEMS>(reinterpret_cast<const uint32_t*>(in), base, n, v);
    else
        load_items_u8<ES_ITEMS>(reinterpret_cast<const uint8_t*>(in), base, n, v);
#pragma unroll
    for (int j = 0; j < ES_ITEMS; ++j) sum += v[j];
    unsigned tile_total;
    unsigned ex = block_excl_sum<unsigned>(sum, s_scan, tile_total);
    if (tile_sums) {
        const unsigned long long p = tile_prefix_lookup(tile_sums, tile_supers, tile, s_red);
        if (threadIdx.x == 0 && tile == gridDim.x - 1 && total) *total = p + tile_total;
        store_running<ES_ITEMS>(out, base, n, (unsigned)p + ex, v);
        return;
    }
    if (threadIdx.x < 32) {
        unsigned long long p = lookback_excl_warp(status, tile, tile_total);
        if (threadIdx.x == 0) {
            s_prefix = p;
            if (tile == gridDim.x - 1 && total) *total = p + tile_total;
        }
    }
    __syncthreads();
    store_running<ES_ITEMS>(out, base, n, (unsigned)s_prefix + ex, v);
}

// Host driver: scratch for status/ticket is taken from the stream pool.
template <typename InT>
int exclusive_scan(const InT* in, uint32_t* out, size_t n, unsigned long long* total_dev,
                   cudaStream_t stream) {
    if (n == 0) {
        if (total_dev) GPC_CUDA(cudaMemsetAsync(total_dev, 0, 8, stream));
        return GPC_OK;
    }
    static_assert(sizeof(InT) == 4 || sizeof(InT) == 1, "exclusive_scan: uint32 or uint8 input");
    unsigned tiles = div_up(n, ES_TILE);
    const unsigned groups = (tiles >> RTS_GROUP_LOG2) + 1;
    Scratch sc;
    GPC_CUDA(sc.alloc(((size_t)tiles + groups) * 8 + 16, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, ((size_t)tiles + groups) * 8 + 16, stream));
    unsigned long long* status = sc.as<unsigned long long>() + 2;
    unsigned long long* supers = status + tiles;
    unsigned* ticket = sc.as<unsigned>();
    if (n >= RTS_MIN_ITEMS) {
        GPC_PROF_BYTES("excl_scan_sums", (double)n * sizeof(InT));
        excl_scan_sums_kernel<InT><<<tiles, SC_THREADS, 0, stream>>>(in, (unsigned)n, status, supers);
        GPC_LAUNCH_CHECK();
        GPC_PROF_BYTES("excl_scan", (double)n * (sizeof(InT) + 4));
        excl_scan_kernel<InT><<<tiles, SC_THREADS, 0, stream>>>(in, out, (unsigned)n, nullptr, nullptr,
                                                                total_dev, status, supers);
        GPC_LAUNCH_CHECK();
        return GPC_OK;
    }
    GPC_PROF_BYTES("excl_scan", (double)n * (sizeof(InT) + 4));
    excl_scan_kernel<InT><<<tiles, SC_THREADS, 0, stream>>>(in, out, (unsigned)n, status, ticket, total_dev);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// Single-pass stream compaction (flag + scan + scatter in one kernel): item i is
// kept when op.keep(i); every kept item is handed to op.emit(i, rank) with its
// rank among the kept items; op.finish(count) runs once.  *total (device)
// receives the number kept.
template <typename Op>
__global__ void __launch_bounds__(SC_THREADS)
select_count_kernel(Op op, unsigned n_max, const unsigned long long* __restrict__ n_dev,
                    unsigned long long* __restrict__ counts, unsigned long long* __restrict__ supers /* zeroed */) {
    __shared__ unsigned s_red[SC_THREADS / 32];
    const unsigned n = n_dev ? (unsigned)min((unsigned long long)n_max, *n_dev) : n_max;
    const unsigned base = blockIdx.x * SC_TILE + threadIdx.x;
    unsigned c = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        const unsigned i = base + j * SC_THREADS;
        c += (i < n && op.keep(i)) ? 1u : 0u;
    }
    c = warp_sum(c);
    if (lane_id() == 0) s_red[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < SC_THREADS / 32; ++w) t += s_red[w];
        tile_total_publish(counts, supers, blockIdx.x, (unsigned long long)t);
    }
}

// tile_prefix == NULL: chained scan over the tiles (status, ticket); else the number of
// items kept before every tile is handed in and the tiles are independent
template <typename Op>
__global__ void __launch_bounds__(SC_THREADS)
select_kernel(Op op, unsigned n_max, const unsigned long long* __restrict__ n_dev,
              unsigned long long* __restrict__ status, unsigned* __restrict__ ticket,
              unsigned long long* __restrict__ total,
              const unsigned long long* __restrict__ tile_sums = nullptr,
              const unsigned long long* __restrict__ tile_supers = nullptr) {
    // striped items (row j of the tile = 256 consecutive items, one per thread):
    // loads and stores of a warp are contiguous; ranks come from ballots
    constexpr int WARPS = SC_THREADS / 32;
    __shared__ unsigned s_cnt[SC_ITEMS * WARPS + 1];
    __shared__ unsigned s_slot;
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned long long s_red[33];
    const unsigned tile = tile_sums ? blockIdx.x : claim_tile(ticket, &s_slot);
    // the item count may still sit in device memory (produced by the previous kernel)
    const unsigned n = n_dev ? (unsigned)min((unsigned long long)n_max, *n_dev) : n_max;
    if ((unsigned long long)tile * SC_TILE >= n) return;       // (whole block, nobody waits on it)
    const unsigned last_tile = (n - 1) / SC_TILE;
    const unsigned base = tile * SC_TILE + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned keep = 0, before[SC_ITEMS];
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        const unsigned i = base + j * SC_THREADS;
        const bool k = i < n && op.keep(i);
        const unsigned b = __ballot_sync(0xffffffffu, k);
        before[j] = __popc(b & ((1u << lane) - 1u));
        if (k) keep |= 1u << j;
        if (lane == 0) s_cnt[j * WARPS + w] = __popc(b);
    }
    __syncthreads();
    unsigned long long known = 0;                    // (uniform over the block)
    if (tile_sums) known = tile_prefix_lookup(tile_sums, tile_supers, tile, s_red);
    if (threadIdx.x < 32) {
        // exclusive scan of the SC_ITEMS x WARPS (= 64) counts, two per lane
        unsigned a0 = s_cnt[2 * lane], a1 = s_cnt[2 * lane + 1];
        unsigned incl = warp_incl_sum(a0 + a1);
        unsigned tile_total = __shfl_sync(0xffffffffu, incl, 31);
        s_cnt[2 * lane] = incl - a0 - a1;
        s_cnt[2 * lane + 1] = incl - a1;
        unsigned long long p = tile_sums ? known
                                         : lookback_excl_warp(status, tile, (unsigned long long)tile_total);
        if (lane == 0) {
            s_prefix = p;
            if (tile == last_tile) {
                if (total) *total = p + tile_total;
                op.finish(p + tile_total);          // number kept, known to the last tile
            }
        }
    }
    __syncthreads();
    const unsigned long long p = s_prefix;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j)
        if ((keep >> j) & 1u)
            op.emit(base + j * SC_THREADS, p + s_cnt[j * WARPS + w] + before[j]);
}

// n_dev (optional): the real item count, <= n, read by the kernel from device memory.
template <typename Op>
int select_if(const char* name, Op op, size_t n, unsigned long long* total_dev, cudaStream_t stream,
              const unsigned long long* n_dev = nullptr, bool keep_is_cheap = true) {
    // keep_is_cheap == false: op.keep() is a dependent gather (the duplicate filter's table
    // probe); evaluating it twice costs more than the chained scan it would save
    if (n == 0) {
        if (total_dev) GPC_CUDA(cudaMemsetAsync(total_dev, 0, 8, stream));
        return GPC_OK;
    }
    unsigned tiles = div_up(n, SC_TILE);
    const unsigned groups = (tiles >> RTS_GROUP_LOG2) + 1;
    Scratch sc;
    GPC_CUDA(sc.alloc(((size_t)tiles + groups) * 8 + 16, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, ((size_t)tiles + groups) * 8 + 16, stream));
    unsigned long long* status = sc.as<unsigned long long>() + 2;
    unsigned long long* supers = status + tiles;
    unsigned* ticket = sc.as<unsigned>();
    if (n >= RTS_MIN_ITEMS && keep_is_cheap) {
        // (tiles beyond a device-side item count contribute zero and leave at once)
        GPC_PROF("select_count");
        select_count_kernel<Op><<<tiles, SC_THREADS, 0, stream>>>(op, (unsigned)n, n_dev, status, supers);
        GPC_LAUNCH_CHECK();
        GPC_PROF(name);
        select_kernel<Op><<<tiles, SC_THREADS, 0, stream>>>(op, (unsigned)n, n_dev, nullptr, nullptr,
                                                            total_dev, status, supers);
        GPC_LAUNCH_CHECK();
        return GPC_OK;
    }
    GPC_PROF(name);
    select_kernel<Op><<<tiles, SC_THREADS, 0, stream>>>(op, (unsigned)n, n_dev, status, ticket,
                                                        total_dev);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// Up to four independent exclusive scans of uint32 arrays of the same length in
// one launch (one chained scan each); totals[k] receives the sum of array k.
struct MultiScanArgs {
    const uint32_t* in[4];
    uint32_t* out[4];
    unsigned long long* status[4];
    int k;
};

static __global__ void __launch_bounds__(SC_THREADS)
multi_scan_kernel(MultiScanArgs a, unsigned n, unsigned long long* __restrict__ totals,
                  unsigned* __restrict__ ticket) {
    __shared__ unsigned s_scan[33];
    __shared__ unsigned s_slot;
    __shared__ unsigned long long s_prefix[4];
    const unsigned tile = claim_tile(ticket, &s_slot);
    const unsigned base = tile * SC_TILE + threadIdx.x * SC_ITEMS;
    for (int k = 0; k < a.k; ++k) {
        unsigned v[SC_ITEMS], sum = 0;
        load_items_u32<SC_ITEMS>(a.in[k], base, n, v);
#pragma unroll
        for (int j = 0; j < SC_ITEMS; ++j) sum += v[j];
        unsigned tile_total;
        unsigned ex = block_excl_sum<unsigned>(sum, s_scan, tile_total);
        if (threadIdx.x < 32) {
            unsigned long long p = lookback_excl_warp(a.status[k], tile, tile_total);
            if (threadIdx.x == 0) {
                s_prefix[k] = p;
                if (tile == gridDim.x - 1) totals[k] = p + tile_total;
            }
        }
        __syncthreads();
        store_running<SC_ITEMS>(a.out[k], base, n, (unsigned)s_prefix[k] + ex, v);
    }
}

// Host driver; totals_host[k] are written after one synchronisation.
inline int multi_scan(int k, const uint32_t* const* in, uint32_t* const* out, size_t n,
                      unsigned long long* totals_host, cudaStream_t stream) {
    for (int i = 0; i < k; ++i) totals_host[i] = 0;
    if (n == 0) return GPC_OK;
    unsigned tiles = div_up(n, SC_TILE);
    Scratch sc;
    size_t bytes = ((size_t)tiles * k + 8) * 8;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    MultiScanArgs a;
    a.k = k;
    unsigned long long* totals = sc.as<unsigned long long>();
    for (int i = 0; i < 4; ++i) {
        a.in[i] = i < k ? in[i] : nullptr;
        a.out[i] = i < k ? out[i] : nullptr;
        a.status[i] = totals + 8 + (size_t)tiles * (i < k ? i : 0);
    }
    unsigned* ticket = reinterpret_cast<unsigned*>(totals + 7);       // zeroed with the rest
    GPC_PROF("multi_scan");
    multi_scan_kernel<<<tiles, SC_THREADS, 0, stream>>>(a, (unsigned)n, totals, ticket);
    GPC_LAUNCH_CHECK();
    GPC_READ_BACK({totals, totals_host, (unsigned)(8 * k)});
    return GPC_OK;
}

}  // namespace gpc
