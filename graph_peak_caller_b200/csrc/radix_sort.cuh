// LSD radix sort, 8- or 9-bit digits, "onesweep" structure:
//   1. one histogram kernel counts every digit of every pass in a single read
//      of the keys;
//   2. per pass, one scatter kernel: each tile (4096-6144 keys) counts its
//      digits and publishes them through a chained scan (decoupled look-back,
//      several predecessors per round trip) BEFORE ranking its keys (stable,
//      one ballot per digit bit), then writes the keys out grouped by digit,
//      staged through shared memory so that stores are coalesced runs.
// Traffic per pass: keys read once, written once (+1-2 KB of tile status).
// Passes whose digit is the same for every key are skipped (64-bit keys).
// Inputs of at most 4096 keys are sorted by one block in shared memory.
#pragma once
#include "common.cuh"

namespace gpc {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

// keys per thread: bounded by the 48 KB of static shared memory a tile's keys
// (and values) are staged in
template <typename KeyT, bool HAS_VALUE = false> struct RsTraits;
#ifndef GPC_RS_ITEMS32
#define GPC_RS_ITEMS32 24
#endif
#ifndef GPC_RS_ITEMS64
#define GPC_RS_ITEMS64 16
#endif
template <> struct RsTraits<uint32_t, false> { static constexpr int ITEMS = GPC_RS_ITEMS32; static constexpr int MAX_PASSES = 4; };
template <> struct RsTraits<uint32_t, true> { static constexpr int ITEMS = 16; static constexpr int MAX_PASSES = 4; };
template <> struct RsTraits<unsigned long long, false> { static constexpr int ITEMS = GPC_RS_ITEMS64; static constexpr int MAX_PASSES = 8; };
template <> struct RsTraits<unsigned long long, true> { static constexpr int ITEMS = 8; static constexpr int MAX_PASSES = 8; };

template <typename KeyT, int RS_BITS>
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const KeyT* __restrict__ keys, size_t n, int begin_bit, int n_passes,
               unsigned* __restrict__ hist /* [n_passes][RADIX] */) {
    constexpr int RS_RADIX = 1 << RS_BITS;
    __shared__ unsigned sh[RsTraits<KeyT>::MAX_PASSES * RS_RADIX];
    for (int i = threadIdx.x; i < n_passes * RS_RADIX; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * RS_THREADS;
    for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        KeyT k = keys[i];
        for (int p = 0; p < n_passes; ++p) {
            unsigned d = (unsigned)(k >> (begin_bit + p * RS_BITS)) & (RS_RADIX - 1);
            atomicAdd(&sh[p * RS_RADIX + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * RS_RADIX; i += RS_THREADS) {
        unsigned v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// One block per pass: exclusive scan of its RADIX counts (in place) and a flag
// telling whether one digit holds every key (pass can be skipped).  Thread t
// owns the DPT consecutive digits t*DPT .. t*DPT+DPT-1.
template <int RS_BITS>
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_hist_kernel(unsigned* __restrict__ hist, unsigned n, unsigned* __restrict__ trivial) {
    constexpr int RS_RADIX = 1 << RS_BITS;
    constexpr int DPT = RS_RADIX / RS_THREADS;
    __shared__ unsigned sm[33];
    unsigned* h = hist + blockIdx.x * RS_RADIX;
    unsigned v[DPT], sum = 0;
#pragma unroll
    for (int t = 0; t < DPT; ++t) {
        v[t] = h[threadIdx.x * DPT + t];
        if (v[t] == n && n > 0) trivial[blockIdx.x] = 1;
        sum += v[t];
    }
    unsigned total;
    unsigned ex = block_excl_sum<unsigned>(sum, sm, total);
#pragma unroll
    for (int t = 0; t < DPT; ++t) {
        h[threadIdx.x * DPT + t] = ex;
        ex += v[t];
    }
}

constexpr unsigned RS_ST_AGG = 1u << 30;
constexpr unsigned RS_ST_PREFIX = 2u << 30;
constexpr unsigned RS_ST_MASK = (1u << 30) - 1;

#ifndef GPC_RS_MIN_BLOCKS
#define GPC_RS_MIN_BLOCKS 3
#endif
template <typename KeyT, bool HAS_VALUE, int RS_BITS>
__global__ void __launch_bounds__(RS_THREADS, GPC_RS_MIN_BLOCKS)
rs_onesweep_kernel(const KeyT* __restrict__ keys_in, KeyT* __restrict__ keys_out,
                   const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                   unsigned n, int shift, const unsigned* __restrict__ digit_base,
                   unsigned* __restrict__ status /* [tiles][RADIX], zeroed */,
                   unsigned* __restrict__ ticket /* zeroed */) {
    constexpr int RS_RADIX = 1 << RS_BITS;
    constexpr int DPT = RS_RADIX / RS_THREADS;
    constexpr int ITEMS = RsTraits<KeyT, HAS_VALUE>::ITEMS;
    constexpr int TILE = RS_THREADS * ITEMS;
    __shared__ unsigned warp_hist[RS_WARPS][RS_RADIX];
    __shared__ unsigned digit_start[RS_RADIX];   // exclusive offsets inside the tile
    __shared__ unsigned out_base[RS_RADIX];      // global index of tile-sorted slot 0 of digit d
    __shared__ KeyT s_keys[TILE];
    __shared__ uint32_t s_vals[HAS_VALUE ? TILE : 1];
    __shared__ unsigned s_scan[33];
    __shared__ unsigned s_tile;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS)
        (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const unsigned tile_base = tile * TILE;
    const unsigned count = min((unsigned)TILE, n - tile_base);
    const unsigned w = warp_id(), l = lane_id();
    const unsigned warp_base = w * (ITEMS * 32);

    KeyT key[ITEMS];
    unsigned short rank[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        unsigned idx = warp_base + j * 32 + l;
        key[j] = (idx < count) ? keys_in[tile_base + idx] : ~KeyT(0);
    }
    // ---- early count: the tile's digit histogram, published before the ranking so
    // that later tiles never wait for a tile that is still ranking its keys ----------
    for (int i = threadIdx.x; i < RS_RADIX; i += RS_THREADS) digit_start[i] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        atomicAdd(&digit_start[(unsigned)(key[j] >> shift) & (RS_RADIX - 1)], 1u);
    __syncthreads();
    {
        unsigned tot[DPT], thread_sum = 0;
#pragma unroll
        for (int t = 0; t < DPT; ++t) {
            tot[t] = digit_start[threadIdx.x * DPT + t];
            thread_sum += tot[t];
        }
        unsigned tile_total;
        unsigned dstart = block_excl_sum<unsigned>(thread_sum, s_scan, tile_total);
        // chained scan of every digit over the tiles; the DPT digits of a thread walk
        // back together so that their loads are in flight at the same time
        unsigned excl[DPT];
        volatile unsigned* st = status;
#pragma unroll
        for (int t = 0; t < DPT; ++t) excl[t] = 0;
        if (tile == 0) {
#pragma unroll
            for (int t = 0; t < DPT; ++t)
                st[tile * RS_RADIX + threadIdx.x * DPT + t] = RS_ST_PREFIX | tot[t];
        } else {
#pragma unroll
            for (int t = 0; t < DPT; ++t)
                st[tile * RS_RADIX + threadIdx.x * DPT + t] = RS_ST_AGG | tot[t];
            // LB_WIDE predecessors per round trip.  With one predecessor per round
            // the walk is as slow as the tiles are spaced and its depth never
            // shrinks (16 round trips per tile were measured); several at once
            // make the walk outrun the wavefront, so the depth collapses.
            constexpr int LB_WIDE = 8 / DPT;
            int tt[DPT];
            bool done[DPT];
#pragma unroll
            for (int t = 0; t < DPT; ++t) { tt[t] = (int)tile - 1; done[t] = false; }
            bool all_done = false;
            while (!all_done) {
                unsigned sv[DPT][LB_WIDE];
#pragma unroll
                for (int t = 0; t < DPT; ++t)
#pragma unroll
                    for (int k = 0; k < LB_WIDE; ++k)
                        sv[t][k] = done[t] ? 0u
                                           : st[max(tt[t] - k, 0) * RS_RADIX + threadIdx.x * DPT + t];
                all_done = true;
#pragma unroll
                for (int t = 0; t < DPT; ++t) {
                    bool open = !done[t];                // false once an empty slot was met
#pragma unroll
                    for (int k = 0; k < LB_WIDE; ++k) {
                        unsigned flag = sv[t][k] & ~RS_ST_MASK;
                        open = open && flag != 0;
                        if (open) {
                            excl[t] += sv[t][k] & RS_ST_MASK;
                            --tt[t];
                            if (flag == RS_ST_PREFIX) { done[t] = true; open = false; }
                        }
                    }
                    all_done &= done[t];
                }
            }
#pragma unroll
            for (int t = 0; t < DPT; ++t)
                st[tile * RS_RADIX + threadIdx.x * DPT + t] = RS_ST_PREFIX | (excl[t] + tot[t]);
        }
#pragma unroll
        for (int t = 0; t < DPT; ++t) {
            const unsigned d = threadIdx.x * DPT + t;
            digit_start[d] = dstart;                 // exclusive offset of digit d inside the tile
            out_base[d] = digit_base[d] + excl[t] - dstart;
            dstart += tot[t];
        }
    }
    // stable ranking: items in (j, lane) order == memory order inside the warp chunk
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        unsigned d = (unsigned)(key[j] >> shift) & (RS_RADIX - 1);
        // lanes holding the same digit, from one ballot per digit bit
        unsigned peers = FULL;
#pragma unroll
        for (int b = 0; b < RS_BITS; ++b) {
            unsigned vote = __ballot_sync(FULL, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? vote : ~vote;
        }
        unsigned leader = __ffs(peers) - 1;
        unsigned before = __popc(peers & ((1u << l) - 1));
        unsigned old = 0;
        if (l == leader) {
            old = warp_hist[w][d];
            warp_hist[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(FULL, old, leader);
        rank[j] = (unsigned short)(old + before);
        __syncwarp();
    }
    __syncthreads();
    // per digit (thread t owns digits t*DPT ..): exclusive scan over the warps
#pragma unroll
    for (int t = 0; t < DPT; ++t) {
        const unsigned d = threadIdx.x * DPT + t;
        unsigned acc = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            unsigned c = warp_hist[ww][d];
            warp_hist[ww][d] = acc;
            acc += c;
        }
    }
    __syncthreads();
    // tile-local sorted position of every key
    unsigned pos[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        unsigned dd = (unsigned)(key[j] >> shift) & (RS_RADIX - 1);
        pos[j] = digit_start[dd] + warp_hist[w][dd] + rank[j];
        s_keys[pos[j]] = key[j];
    }
    if (HAS_VALUE) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            unsigned idx = warp_base + j * 32 + l;
            if (idx < count) s_vals[pos[j]] = vals_in[tile_base + idx];
        }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < count; i += RS_THREADS) {
        KeyT k = s_keys[i];
        unsigned dd = (unsigned)(k >> shift) & (RS_RADIX - 1);
        unsigned dst = out_base[dd] + i;
        keys_out[dst] = k;
        if (HAS_VALUE) vals_out[dst] = s_vals[i];
    }
}

// Host driver.  Sorts bits [begin_bit, end_bit) of `keys` (n < 2^30).
// `keys`/`vals` and `keys_tmp`/`vals_tmp` ping-pong; *result_in_tmp says where
// the sorted data ended up.  `skip_trivial` costs one 32-byte D2H read.
template <typename KeyT, bool HAS_VALUE, int RS_BITS>
int radix_sort_bits(KeyT* keys, KeyT* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, size_t n,
                    int begin_bit, int end_bit, bool skip_trivial, bool* result_in_tmp,
                    cudaStream_t stream) {
    constexpr int RS_RADIX = 1 << RS_BITS;
    *result_in_tmp = false;
    if (n <= 1) return GPC_OK;
    if (n >= (1ull << 30)) {
        set_error("radix_sort: %zu keys exceed the 2^30 limit", n);
        return GPC_ERR_ARG;
    }
    constexpr int TILE = RS_THREADS * RsTraits<KeyT, HAS_VALUE>::ITEMS;
    const int n_passes = (end_bit - begin_bit + RS_BITS - 1) / RS_BITS;
    if (n_passes <= 0) return GPC_OK;
    if (n_passes > RsTraits<KeyT>::MAX_PASSES) {
        set_error("radix_sort: too many passes");
        return GPC_ERR_ARG;
    }
    const unsigned tiles = div_up(n, TILE);
    // scratch: hist[n_passes*256] | trivial[8] | ticket[8] | status[tiles*256] per pass
    size_t head = (size_t)n_passes * RS_RADIX + 16;
    size_t per_pass = (size_t)tiles * RS_RADIX;
    Scratch sc;
    GPC_CUDA(sc.alloc((head + per_pass * n_passes) * sizeof(unsigned), stream));
    unsigned* hist = sc.as<unsigned>();
    unsigned* trivial = hist + (size_t)n_passes * RS_RADIX;
    unsigned* ticket = trivial + 8;
    unsigned* status = hist + head;
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, (head + per_pass * n_passes) * sizeof(unsigned), stream));
    unsigned hist_blocks = min(div_up(n, RS_THREADS * 8), (unsigned)(sm_count() * 8));
    GPC_PROF_BYTES("rs_hist", (double)n * sizeof(KeyT));
    rs_hist_kernel<KeyT, RS_BITS><<<hist_blocks, RS_THREADS, 0, stream>>>(keys, n, begin_bit, n_passes, hist);
    GPC_LAUNCH_CHECK();
    GPC_PROF("rs_scan_hist");
    rs_scan_hist_kernel<RS_BITS><<<n_passes, RS_THREADS, 0, stream>>>(hist, (unsigned)n, trivial);
    GPC_LAUNCH_CHECK();
    unsigned h_trivial[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (skip_trivial) {
        GPC_READ_BACK({trivial, h_trivial, (unsigned)(sizeof(h_trivial))});
    }
    bool in_tmp = false;
    for (int p = 0; p < n_passes; ++p) {
        if (h_trivial[p]) continue;
        const KeyT* kin = in_tmp ? keys_tmp : keys;
        KeyT* kout = in_tmp ? keys : keys_tmp;
        const uint32_t* vin = in_tmp ? vals_tmp : vals;
        uint32_t* vout = in_tmp ? vals : vals_tmp;
        GPC_PROF_BYTES("rs_onesweep", 2.0 * n * (sizeof(KeyT) + (HAS_VALUE ? 4 : 0)));
        rs_onesweep_kernel<KeyT, HAS_VALUE, RS_BITS><<<tiles, RS_THREADS, 0, stream>>>(
            kin, kout, vin, vout, (unsigned)n, begin_bit + p * RS_BITS,
            hist + (size_t)p * RS_RADIX, status + per_pass * p, ticket + p);
        GPC_LAUNCH_CHECK();
        in_tmp = !in_tmp;
    }
    *result_in_tmp = in_tmp;
    return GPC_OK;
}

// ---- small inputs: the whole sort in one block ------------------------------------
// Up to RS_SMALL_N keys are sorted by ONE block entirely in shared memory (all
// passes in one launch, no histogram kernels, no tile status, no read-back): a
// 1000-key sort used to cost 8 passes x ~10 us of launch + latency.
constexpr int RS_SMALL_THREADS = 512;
constexpr int RS_SMALL_WARPS = RS_SMALL_THREADS / 32;
constexpr int RS_SMALL_ITEMS = 8;
constexpr int RS_SMALL_N = RS_SMALL_THREADS * RS_SMALL_ITEMS;      // 4096

template <typename KeyT, bool HAS_VALUE>
constexpr size_t rs_small_smem() {
    return 2 * RS_SMALL_N * sizeof(KeyT) + (HAS_VALUE ? 2 * RS_SMALL_N * 4 : 0) +
           RS_SMALL_WARPS * 256 * 4 + 256 * 4 + 64 * 4;
}

template <typename KeyT, bool HAS_VALUE>
__global__ void __launch_bounds__(RS_SMALL_THREADS)
rs_small_kernel(KeyT* __restrict__ keys, uint32_t* __restrict__ vals, unsigned n, int begin_bit,
                int end_bit) {
    extern __shared__ __align__(16) unsigned char rs_small_raw[];
    KeyT* kbuf = reinterpret_cast<KeyT*>(rs_small_raw);                           // [2][N]
    uint32_t* vbuf = reinterpret_cast<uint32_t*>(kbuf + 2 * RS_SMALL_N);          // [2][N] if HAS_VALUE
    unsigned* warp_hist = vbuf + (HAS_VALUE ? 2 * RS_SMALL_N : 0);                // [WARPS][256]
    unsigned* digit_start = warp_hist + RS_SMALL_WARPS * 256;                      // [256]
    unsigned* s_scan = digit_start + 256;                                          // [33]
    KeyT* s_or = reinterpret_cast<KeyT*>(s_scan + 34);                             // [2], 8-byte aligned
    const unsigned w = warp_id(), l = lane_id();
    // load (padding = all ones: largest digit in every pass, behind all real keys)
    KeyT v_or = 0, v_and = ~KeyT(0);
    for (unsigned i = threadIdx.x; i < RS_SMALL_N; i += RS_SMALL_THREADS) {
        KeyT k = i < n ? keys[i] : ~KeyT(0);
        kbuf[i] = k;
        if (HAS_VALUE) vbuf[i] = i < n ? vals[i] : 0u;
        if (i < n) { v_or |= k; v_and &= k; }
    }
    // bits that differ between keys: passes over constant digits are skipped
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        v_or |= __shfl_xor_sync(FULL, v_or, d);
        v_and &= __shfl_xor_sync(FULL, v_and, d);
    }
    if (threadIdx.x == 0) { s_or[0] = 0; s_or[1] = ~KeyT(0); }
    __syncthreads();
    if (l == 0) {
        if (sizeof(KeyT) == 8) {
            atomicOr(reinterpret_cast<unsigned long long*>(&s_or[0]), (unsigned long long)v_or);
            atomicAnd(reinterpret_cast<unsigned long long*>(&s_or[1]), (unsigned long long)v_and);
        } else {
            atomicOr(reinterpret_cast<unsigned*>(&s_or[0]), (unsigned)v_or);
            atomicAnd(reinterpret_cast<unsigned*>(&s_or[1]), (unsigned)v_and);
        }
    }
    __syncthreads();
    const KeyT varying = s_or[0] ^ s_or[1];
    int cur = 0;
    for (int bit = begin_bit; bit < end_bit; bit += 8) {
        const int width = min(8, end_bit - bit);
        const unsigned mask = (1u << width) - 1u;
        if ((((unsigned)(varying >> bit)) & mask) == 0) continue;        // uniform decision
        KeyT* kin = kbuf + cur * RS_SMALL_N;
        KeyT* kout = kbuf + (cur ^ 1) * RS_SMALL_N;
        uint32_t* vin = vbuf + cur * RS_SMALL_N;
        uint32_t* vout = vbuf + (cur ^ 1) * RS_SMALL_N;
        for (int i = threadIdx.x; i < RS_SMALL_WARPS * 256; i += RS_SMALL_THREADS) warp_hist[i] = 0;
        __syncthreads();
        KeyT key[RS_SMALL_ITEMS];
        unsigned rank[RS_SMALL_ITEMS];
        const unsigned warp_base = w * (RS_SMALL_ITEMS * 32);
#pragma unroll
        for (int j = 0; j < RS_SMALL_ITEMS; ++j) {
            key[j] = kin[warp_base + j * 32 + l];
            unsigned d = (unsigned)(key[j] >> bit) & mask;
            unsigned peers = FULL;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                unsigned vote = __ballot_sync(FULL, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? vote : ~vote;
            }
            unsigned leader = __ffs(peers) - 1;
            unsigned old = 0;
            if (l == leader) {
                old = warp_hist[w * 256 + d];
                warp_hist[w * 256 + d] = old + __popc(peers);
            }
            old = __shfl_sync(FULL, old, leader);
            rank[j] = old + __popc(peers & ((1u << l) - 1u));
            __syncwarp();
        }
        __syncthreads();
        unsigned total = 0;
        if (threadIdx.x < 256) {
            for (int ww = 0; ww < RS_SMALL_WARPS; ++ww) {
                unsigned c = warp_hist[ww * 256 + threadIdx.x];
                warp_hist[ww * 256 + threadIdx.x] = total;
                total += c;
            }
        }
        unsigned all;
        unsigned start = block_excl_sum<unsigned>(total, s_scan, all);
        if (threadIdx.x < 256) digit_start[threadIdx.x] = start;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_SMALL_ITEMS; ++j) {
            unsigned d = (unsigned)(key[j] >> bit) & mask;
            unsigned pos = digit_start[d] + warp_hist[w * 256 + d] + rank[j];
            kout[pos] = key[j];
            if (HAS_VALUE) vout[pos] = vin[warp_base + j * 32 + l];
        }
        __syncthreads();
        cur ^= 1;
    }
    for (unsigned i = threadIdx.x; i < n; i += RS_SMALL_THREADS) {
        keys[i] = kbuf[cur * RS_SMALL_N + i];
        if (HAS_VALUE) vals[i] = vbuf[cur * RS_SMALL_N + i];
    }
}

template <typename KeyT, bool HAS_VALUE>
int radix_sort_small(KeyT* keys, uint32_t* vals, size_t n, int begin_bit, int end_bit,
                     cudaStream_t stream) {
    static bool configured = false;          // opt in to > 48 KB of dynamic shared memory once
    if (!configured) {
        GPC_CUDA(cudaFuncSetAttribute(rs_small_kernel<KeyT, HAS_VALUE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)rs_small_smem<KeyT, HAS_VALUE>()));
        configured = true;
    }
    GPC_PROF("rs_small");
    rs_small_kernel<KeyT, HAS_VALUE><<<1, RS_SMALL_THREADS, rs_small_smem<KeyT, HAS_VALUE>(), stream>>>(
        keys, vals, (unsigned)n, begin_bit, end_bit);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

// Picks the digit width: 9-bit digits when that saves a pass for 32-bit keys
// (e.g. 26 significant bits: 3 passes instead of 4), else 8-bit digits.
template <typename KeyT, bool HAS_VALUE>
int radix_sort(KeyT* keys, KeyT* keys_tmp, uint32_t* vals, uint32_t* vals_tmp, size_t n,
               int begin_bit, int end_bit, bool skip_trivial, bool* result_in_tmp,
               cudaStream_t stream) {
    const int range = end_bit - begin_bit;
    if (n <= (size_t)RS_SMALL_N && n > 1 && range > 0) {
        *result_in_tmp = false;
        return radix_sort_small<KeyT, HAS_VALUE>(keys, vals, n, begin_bit, end_bit, stream);
    }
    if constexpr (sizeof(KeyT) == 4 && !HAS_VALUE) {
        if ((range + 8) / 9 < (range + 7) / 8)
            return radix_sort_bits<KeyT, HAS_VALUE, 9>(keys, keys_tmp, vals, vals_tmp, n, begin_bit,
                                                       end_bit, skip_trivial, result_in_tmp, stream);
    }
    return radix_sort_bits<KeyT, HAS_VALUE, 8>(keys, keys_tmp, vals, vals_tmp, n, begin_bit, end_bit,
                                               skip_trivial, result_in_tmp, stream);
}

}  // namespace gpc
