// Background (lambda) track on sm_100a.
//
// Replaces the reference's
//   LinearMap.map_interval_collection / graph_position_to_linear   control/linearmap.py:51-74
//   SparseControl.create                                           control/controlgenerator.py:21-44
//   SparseDiffs.from_starts_and_ends / clip_min / maximum          sparsediffs.py:130-178
//   LinearPileup.sanitize_*, get_event_sorter, from_event_sorter   control/linearpileup.py:56-143
//   LinearMap.to_sparse_pileup                                     control/linearmap.py:76-99
//
// Formulation.  Every control read start is projected to the linear axis with
// the reference's exact float64 operations and the R projected starts are
// radix sorted once (64-bit order-preserving keys) and run-length compressed
// (distinct x, multiplicity).  Each extension window w turns a start x into a
// +1 at x-e_w and a -1 at x+e_w, so the 2W event streams are all shifted views
// of the same sorted array.  They are merged without ever materialising or
// sorting the 2WR events: every T-th element of every view is a splitter, the
// sorted splitters cut the value axis into tiles that hold at most ~T elements
// of each view, and one CTA per tile merges its <= 2W(T+2) elements in shared
// memory (rank by binary search), then scans integer counts per window.  The
// counts at a tile's left edge follow directly from the split indexes, so tiles
// are independent; only the output offset goes through a chained scan.
// The linear lambda is
//     max( min_value , max_w count_w * weight_w )
// (before the first event of window 0 the clip and window 0 do not apply: that
// is what clip_min on the first track followed by maximum() does).  Counts are
// integers, so lambda = count * weight is exact where the reference accumulates
// +-weight steps in float64; the two agree to ~1e-13 relative.
// Back-projection to graph positions is two binary searches per node plus a
// segmented copy, with numpy's float floor-division replicated bit for bit.
#include <stdlib.h>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace gpc {

constexpr int MAX_WINDOWS = 3;

struct WindowParams {
    int n_windows;
    double half[MAX_WINDOWS];     // e_w = extension // 2
    double weight[MAX_WINDOWS];   // 1.0 / (2 e_w / fragment_length)
    double min_value;
};

// x = node_start + scale * offset   (forward)   /   node_end - scale * offset (reverse)
// with scale = (node_end - node_start) / node_size; no FMA contraction.
__global__ void project_keys_kernel(gpc_graph_t graph, const int32_t* __restrict__ start_node,
                                    const int32_t* __restrict__ start_off,
                                    const uint32_t* __restrict__ read_index, unsigned n_reads,
                                    unsigned long long* __restrict__ keys,
                                    uint32_t* __restrict__ keys32,
                                    double* __restrict__ x_out, int* __restrict__ err) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    unsigned r = read_index ? read_index[i] : i;
    int id = start_node[r];
    long long v = (long long)(id < 0 ? -id : id) - graph.min_node;
    if (v < 0 || v >= (long long)graph.n_nodes) {
        atomicExch(err, GPC_ERR_INVALID_INTERVAL);
        v = 0;
    }
    double ns = graph.lin_start[v], ne = graph.lin_end[v];
    double size = (double)node_record(graph, (uint32_t)v).w;
    double scale = __ddiv_rn(__dsub_rn(ne, ns), size);
    double so = __dmul_rn(scale, (double)start_off[r]);
    double x = id > 0 ? __dadd_rn(ns, so) : __dsub_rn(ne, so);
    if (x_out) x_out[i] = x;
    if (keys32) {
        // the linear map is flagged all-integer (GPC_GRAPH_INTEGER_LINEAR): the start IS a
        // small non-negative integer and is sorted as one (three 9-bit passes instead of five
        // 8-bit passes over an order-preserving image of the double)
        if (!(x == floor(x)) || !(x >= 0.0) || !(x < 1073741824.0)) { err[1] = 1; x = 0.0; }
        keys32[i] = (uint32_t)x;
    } else {
        keys[i] = f64_to_key(x);
    }
}

// distinct x and the index of their first occurrence (== weight prefix)
// distinct projected starts with their multiplicities (as prefix counts), plus
// the check that every start is a non-negative integer below 2^30
struct UniqueStartsOp {
    const unsigned long long* keys;
    unsigned n;
    double* xu;             // [n]  distinct starts, ascending
    uint32_t* wpre;         // [n + 1]  index of the first copy of each; wpre[count] = n
    int* not_integer;
    __device__ __forceinline__ bool keep(unsigned i) const {
        double x = key_to_f64(keys[i]);
        if (!(x == floor(x)) || !(x >= 0.0) || !(x < 1073741824.0)) *not_integer = 1;
        return i == 0 || keys[i] != keys[i - 1];
    }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        xu[w] = key_to_f64(keys[i]);
        wpre[w] = i;
    }
    __device__ __forceinline__ void finish(unsigned long long count) const { wpre[count] = n; }
};

struct UniqueStartsOp32 {                       // the same over integer keys
    const uint32_t* keys;
    unsigned n;
    double* xu;
    uint32_t* wpre;
    __device__ __forceinline__ bool keep(unsigned i) const { return i == 0 || keys[i] != keys[i - 1]; }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        xu[w] = (double)keys[i];
        wpre[w] = i;
    }
    __device__ __forceinline__ void finish(unsigned long long count) const { wpre[count] = n; }
};

// ---- dense path: every projected start is a non-negative integer -----------------
// The linear axis is cut into tiles of DENSE_W positions.  The events of a tile
// are contiguous slices of the sorted starts in every view (two binary searches
// per view); they are scattered into per-window difference arrays in shared
// memory, one block scan turns those into counts per position, and a breakpoint
// is written wherever lambda differs from the position before.  Counts at the
// tile's left edge come from the slice starts, so tiles are independent.
constexpr int DENSE_W = 2048;
constexpr int DENSE_THREADS = 256;
constexpr int DENSE_PER = DENSE_W / DENSE_THREADS;

__device__ __forceinline__ unsigned lower_bound_xu(const double* xu, unsigned n, double x) {
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (xu[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}



// view q of the sorted starts: q = 2*w + is_end  ->  x -/+ half[w]
__device__ __forceinline__ double view_value(double x, int q, const WindowParams& wp) {
    double h = wp.half[q >> 1];
    return (q & 1) ? __dadd_rn(x, h) : __dsub_rn(x, h);
}

constexpr int MERGE_T = 128;                       // splitter stride per view
constexpr int MERGE_THREADS = 256;
constexpr int MERGE_MAXQ = 2 * MAX_WINDOWS;
constexpr int MERGE_CAP = 3072;                    // elements per tile (shared memory)
constexpr int MERGE_PER_THREAD = MERGE_CAP / MERGE_THREADS;
constexpr size_t MERGE_SMEM = (size_t)MERGE_CAP * (8 + 8 + 4 + 1);

__global__ void splitters_kernel(const double* __restrict__ xu, unsigned n_unique, WindowParams wp,
                                 unsigned per_view, unsigned long long* __restrict__ sp) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nq = 2 * wp.n_windows;
    if (i >= per_view * nq) return;
    unsigned q = i / per_view, j = i % per_view;
    sp[i] = f64_to_key(view_value(xu[(size_t)j * MERGE_T], (int)q, wp));
}

// bounds[k][q] = first index i with view_q(i) >= splitter k-1   (k = 1..ns);
// bounds[0][q] = 0, bounds[ns+1][q] = n_unique.  Tile k = [bounds[k], bounds[k+1]).
__global__ void tile_bounds_kernel(const double* __restrict__ xu, unsigned n_unique, WindowParams wp,
                                   const unsigned long long* __restrict__ sp, unsigned ns,
                                   uint32_t* __restrict__ bounds) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nq = 2 * wp.n_windows;
    if (i >= (ns + 2) * nq) return;
    unsigned k = i / nq, q = i % nq;
    uint32_t r;
    if (k == 0) r = 0;
    else if (k == ns + 1) r = n_unique;
    else {
        unsigned long long s = sp[k - 1];
        unsigned lo = 0, hi = n_unique;
        while (lo < hi) {
            unsigned mid = (lo + hi) >> 1;
            if (f64_to_key(view_value(xu[mid], (int)q, wp)) < s) lo = mid + 1; else hi = mid;
        }
        r = lo;
    }
    bounds[i] = r;
}

// merged rank of every splitter boundary, then tile edges: tile t starts at the
// last boundary whose rank is <= t * stride.  Consecutive boundaries are at most
// nq * (T + 2) elements apart, so a tile holds at most stride + nq * (T + 2).
__global__ void boundary_rank_kernel(const uint32_t* __restrict__ bounds, unsigned n_bounds,
                                     unsigned nq, unsigned long long* __restrict__ rank) {
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_bounds) return;
    unsigned long long r = 0;
    for (unsigned q = 0; q < nq; ++q) r += bounds[(size_t)k * nq + q];
    rank[k] = r;
}
__global__ void pick_tiles_kernel(const unsigned long long* __restrict__ rank, unsigned n_bounds,
                                  unsigned n_tiles, unsigned long long stride,
                                  uint32_t* __restrict__ pick) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) { pick[t] = n_bounds - 1; return; }
    unsigned long long target = (unsigned long long)t * stride;
    unsigned lo = 0, hi = n_bounds;            // first boundary with rank > target
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (rank[mid] <= target) lo = mid + 1; else hi = mid;
    }
    pick[t] = lo - 1;                          // rank[0] == 0 <= target, so lo >= 1
}

constexpr int NC = 4;            // count per window (3) + "window-0 starts seen"

// exact double of a non-negative int below 2^31 with one DADD: I2F.F64 runs at a
// fraction of the FP64 rate and was 22 % of the dense kernel's stall samples
__device__ __forceinline__ double count_to_f64(int c) {
    return __hiloint2double(0x43300000, c) - 4503599627370496.0;
}

__device__ __forceinline__ double lambda_of(const int c[NC], const WindowParams& wp) {
    double best = 0.0;
    bool clipped = c[3] > 0;
#pragma unroll
    for (int w = 0; w < MAX_WINDOWS; ++w)      // static indices: wp stays in the constant bank
        if (w < wp.n_windows && (w > 0 || clipped))
            best = fmax(best, __dmul_rn(count_to_f64(c[w]), wp.weight[w]));
    if (clipped) best = fmax(best, wp.min_value);
    return best;
}

// slice edges of every tile in every view: bounds[t][q] = first start whose event
// in view q lies at or after the left edge of tile t (t = 0 .. n_tiles)
__global__ void dense_bounds_kernel(const double* __restrict__ xu, unsigned n_unique,
                                    WindowParams wp, long long origin, unsigned n_tiles,
                                    uint32_t* __restrict__ bounds) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned nq = 2 * wp.n_windows;
    if (i >= (n_tiles + 1) * nq) return;
    unsigned t = i / nq, q = i % nq;
    double h = wp.half[q >> 1];
    double x = (double)(origin + (long long)t * DENSE_W) + ((q & 1) ? -h : h);
    bounds[i] = lower_bound_xu(xu, n_unique, x);
}

__global__ void __launch_bounds__(DENSE_THREADS)
linear_dense_kernel(const double* __restrict__ xu, const uint32_t* __restrict__ wpre,
                    unsigned n_unique, WindowParams wp, long long origin, unsigned n_tiles,
                    const uint32_t* __restrict__ bounds,
                    double* __restrict__ tmp_pos, double* __restrict__ tmp_val,
                    uint32_t* __restrict__ tile_base, uint32_t* __restrict__ tile_cnt,
                    unsigned long long* __restrict__ out_count) {
    __shared__ int s_diff[NC][DENSE_W];
    __shared__ uint32_t s_lo[MERGE_MAXQ], s_hi[MERGE_MAXQ];
    __shared__ int s_warp[DENSE_THREADS / 32][NC];
    __shared__ int s_scan[33];
    __shared__ unsigned long long s_obase;
    const unsigned tile = blockIdx.x;
    const int nq = 2 * wp.n_windows;
    const long long a = origin + (long long)tile * DENSE_W;      // first position of the tile
    if ((int)threadIdx.x < nq) {
        s_lo[threadIdx.x] = bounds[(size_t)tile * nq + threadIdx.x];
        s_hi[threadIdx.x] = bounds[(size_t)(tile + 1) * nq + threadIdx.x];
    }
    for (int i = threadIdx.x; i < NC * DENSE_W; i += DENSE_THREADS) (&s_diff[0][0])[i] = 0;
    __syncthreads();
    // scatter the events: the slices of all views as one index space, so that the
    // loads of different views are in flight together
    {
        unsigned pre[MERGE_MAXQ + 1];
        pre[0] = 0;
#pragma unroll
        for (int q = 0; q < MERGE_MAXQ; ++q)
            pre[q + 1] = pre[q] + (q < nq ? s_hi[q] - s_lo[q] : 0u);
        for (unsigned f = threadIdx.x; f < pre[MERGE_MAXQ]; f += DENSE_THREADS) {
            int q = 0;
            unsigned first = 0;                // pre[q], without indexing the array
#pragma unroll
            for (int qq = 1; qq < MERGE_MAXQ; ++qq)
                if (f >= pre[qq]) { q = qq; first = pre[qq]; }
            const unsigned i = s_lo[q] + (f - first);
            const int w = q >> 1;
            const double h = w == 0 ? wp.half[0] : (w == 1 ? wp.half[1] : wp.half[2]);
            double t = (q & 1) ? xu[i] + h : xu[i] - h;            // exact: integers
            int p = (int)((long long)t - a);
            int mul = (int)(wpre[i + 1] - wpre[i]);
            atomicAdd(&s_diff[w][p], (q & 1) ? -mul : mul);
            if (q == 0) atomicAdd(&s_diff[3][p], mul);
        }
    }
    __syncthreads();
    // counts per position: thread owns DENSE_PER consecutive positions
    const int p0 = threadIdx.x * DENSE_PER;
    int tot[NC] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int j = 0; j < DENSE_PER; ++j) tot[k] += s_diff[k][p0 + j];
    int run[NC];
    {
        int incl[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            incl[k] = warp_incl_sum(tot[k]);
            if (lane_id() == 31) s_warp[warp_id()][k] = incl[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            // counts in force before the tile, straight from the slice starts
            int pre = (k < wp.n_windows) ? (int)wpre[s_lo[2 * k]] - (int)wpre[s_lo[2 * k + 1]]
                                         : (k == 3 ? (int)wpre[s_lo[0]] : 0);
            for (unsigned ww = 0; ww < warp_id(); ++ww) pre += s_warp[ww][k];
            run[k] = pre + incl[k] - tot[k];
        }
    }
    double was = lambda_of(run, wp);
    double e_val[DENSE_PER];
    unsigned emit = 0;                         // bit j: position p0 + j is a breakpoint
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < DENSE_PER; ++j) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            int d = s_diff[k][p0 + j];
            run[k] += d;
            any |= d != 0;
        }
        e_val[j] = 0.0;
        if (any) {
            double now = lambda_of(run, wp);
            if (now != was) {
                e_val[j] = now;
                emit |= 1u << j;
                ++cnt;
            }
            was = now;
        }
    }
    // Output space is reserved with one atomicAdd per tile (chunks land in
    // completion order); dense_gather_kernel puts them in tile order afterwards.
    // A chained scan would serialise here: equally sized tiles finish as one
    // wavefront and each would walk back through the whole wave.
    int tile_out;
    int o = block_excl_sum<int>(cnt, s_scan, tile_out);
    if (threadIdx.x == 0) {
        unsigned long long b = tile_out ? atomicAdd(out_count, (unsigned long long)tile_out) : 0ull;
        s_obase = b;
        tile_base[tile] = (uint32_t)b;
        tile_cnt[tile] = (uint32_t)tile_out;
    }
    __syncthreads();
    unsigned long long wbase = s_obase + o;
#pragma unroll
    for (int j = 0; j < DENSE_PER; ++j)
        if ((emit >> j) & 1u) {
            tmp_pos[wbase] = (double)(a + p0 + j);
            tmp_val[wbase] = e_val[j];
            ++wbase;
        }
}

// one warp per tile: move the tile's chunk to its place in tile order
__global__ void dense_gather_kernel(const double* __restrict__ tmp_pos,
                                    const double* __restrict__ tmp_val,
                                    const uint32_t* __restrict__ tile_base,
                                    const uint32_t* __restrict__ tile_cnt,
                                    const uint32_t* __restrict__ tile_off, unsigned n_tiles,
                                    double* __restrict__ out_pos, double* __restrict__ out_val) {
    unsigned t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    unsigned c = tile_cnt[t], src = tile_base[t], dst = tile_off[t];
    for (unsigned i = lane_id(); i < c; i += 32) {
        out_pos[dst + i] = tmp_pos[src + i];
        out_val[dst + i] = tmp_val[src + i];
    }
}

__global__ void __launch_bounds__(MERGE_THREADS)
linear_tiles_kernel(const double* __restrict__ xu, const uint32_t* __restrict__ wpre,
                    WindowParams wp, const uint32_t* __restrict__ bounds,
                    const uint32_t* __restrict__ pick, unsigned n_tiles,
                    double* __restrict__ out_pos, double* __restrict__ out_val,
                    unsigned long long* __restrict__ status, unsigned* __restrict__ ticket,
                    unsigned long long* __restrict__ out_count, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long* s_stage = reinterpret_cast<unsigned long long*>(s_raw);   // view-major
    unsigned long long* s_key = s_stage + MERGE_CAP;                              // merged order
    uint32_t* s_mul = reinterpret_cast<uint32_t*>(s_key + MERGE_CAP);
    uint8_t* s_tag = reinterpret_cast<uint8_t*>(s_mul + MERGE_CAP);
    __shared__ uint32_t s_lo[MERGE_MAXQ], s_hi[MERGE_MAXQ], s_off[MERGE_MAXQ + 1];
    __shared__ int s_base[NC];
    __shared__ int s_warp[MERGE_THREADS / 32][NC];
    __shared__ int s_scan[33];
    __shared__ unsigned long long s_obase;
    // tiles are taken in launch order (blockIdx): a tile only waits on earlier ones
    const unsigned tile = blockIdx.x;
    const int nq = 2 * wp.n_windows;
    if ((int)threadIdx.x < nq) {
        s_lo[threadIdx.x] = bounds[(size_t)pick[tile] * nq + threadIdx.x];
        s_hi[threadIdx.x] = bounds[(size_t)pick[tile + 1] * nq + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned off = 0;
        for (int q = 0; q < nq; ++q) { s_off[q] = off; off += s_hi[q] - s_lo[q]; }
        s_off[nq] = off;
    }
    // counts in force at the tile's left edge, straight from the split indexes
    if (threadIdx.x >= 32 && (int)threadIdx.x < 32 + NC) {
        int k = threadIdx.x - 32;
        int v = 0;
        if (k < wp.n_windows) v = (int)wpre[s_lo[2 * k]] - (int)wpre[s_lo[2 * k + 1]];
        else if (k == 3) v = (int)wpre[s_lo[0]];
        s_base[k] = v;
    }
    __syncthreads();
    const unsigned m = s_off[nq];
    if (m > MERGE_CAP) {          // cannot happen (<= T+2 per view); fail loudly
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) atomicExch(err, GPC_ERR_INTERNAL);
            lookback_excl_warp(status, tile, 0ull);
        }
        return;
    }
    // stage every view's slice as keys
    for (unsigned e = threadIdx.x; e < m; e += MERGE_THREADS) {
        int q = 0;
        while (e >= s_off[q + 1]) ++q;
        unsigned i = s_lo[q] + (e - s_off[q]);
        s_stage[e] = f64_to_key(view_value(xu[i], q, wp));
    }
    __syncthreads();
    // merged rank = own offset + elements of the other views that sort before
    // (ties between views are broken by view id, so ranks are unique).  Every
    // thread ranks a contiguous chunk of one view: the first element costs a
    // binary search per other view, the following ones just advance pointers.
    const unsigned per = (m + MERGE_THREADS - 1) / MERGE_THREADS;
    const unsigned a0 = min(m, threadIdx.x * per), a1 = min(m, a0 + per);
    {
        unsigned ptr[MERGE_MAXQ];
        int q = -1;
        for (unsigned e = a0; e < a1; ++e) {
            const unsigned long long key = s_stage[e];
            const bool fresh = (q < 0) || e >= s_off[q + 1];
            if (fresh) {
                q = 0;
                while (e >= s_off[q + 1]) ++q;
            }
            unsigned rank = e - s_off[q];
#pragma unroll
            for (int o = 0; o < MERGE_MAXQ; ++o) {
                if (o >= nq || o == q) continue;
                const unsigned end = s_off[o + 1];
                unsigned p;
                if (fresh) {
                    unsigned lo = s_off[o], hi = end;
                    if (o < q) { while (lo < hi) { unsigned mid = (lo + hi) >> 1; if (s_stage[mid] <= key) lo = mid + 1; else hi = mid; } }
                    else       { while (lo < hi) { unsigned mid = (lo + hi) >> 1; if (s_stage[mid] <  key) lo = mid + 1; else hi = mid; } }
                    p = lo;
                } else {
                    p = ptr[o];
                    if (o < q) { while (p < end && s_stage[p] <= key) ++p; }
                    else       { while (p < end && s_stage[p] <  key) ++p; }
                }
                ptr[o] = p;
                rank += p - s_off[o];
            }
            unsigned i = s_lo[q] + (e - s_off[q]);
            s_key[rank] = key;
            s_tag[rank] = (uint8_t)q;
            s_mul[rank] = wpre[i + 1] - wpre[i];
        }
    }
    __syncthreads();
    // same contiguous chunks, now over the merged order: block scan of the counters
    int tot[NC] = {0, 0, 0, 0};
    for (unsigned r = a0; r < a1; ++r) {
        int q = s_tag[r], mul = (int)s_mul[r];
        tot[q >> 1] += (q & 1) ? -mul : mul;
        if (q == 0) tot[3] += mul;
    }
    // block exclusive scan of the four counters with one barrier pair
    int run[NC];
    {
        int incl[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            incl[k] = warp_incl_sum(tot[k]);
            if (lane_id() == 31) s_warp[warp_id()][k] = incl[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            int pre = s_base[k];
            for (unsigned w = 0; w < warp_id(); ++w) pre += s_warp[w][k];
            run[k] = pre + incl[k] - tot[k];
        }
    }
    // emit one breakpoint per group of equal keys whose lambda differs from the
    // lambda before the group (groups never span tiles: tiles are cut by value).
    // `was` is the lambda in force before the current group; inside a chunk it is
    // simply the previous group's lambda, only the chunk's first group has to
    // reconstruct it (its group may have started in an earlier thread's chunk).
    double e_pos[MERGE_PER_THREAD], e_val[MERGE_PER_THREAD];
    int cnt = 0;
    double was = 0.0;
    if (a0 < a1) {
        int before[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) before[k] = run[k];
        for (unsigned g = a0; g-- > 0;) {              // earlier members of my first group
            if (s_key[g] != s_key[a0]) break;
            int qq = s_tag[g], mm = (int)s_mul[g];
            before[qq >> 1] -= (qq & 1) ? -mm : mm;
            if (qq == 0) before[3] -= mm;
        }
        was = lambda_of(before, wp);
    }
    for (unsigned r = a0; r < a1; ++r) {
        int q = s_tag[r], mul = (int)s_mul[r];
        run[q >> 1] += (q & 1) ? -mul : mul;
        if (q == 0) run[3] += mul;
        bool last = (r + 1 >= m) || s_key[r + 1] != s_key[r];
        if (last) {
            double now = lambda_of(run, wp);
            if (now != was && cnt < MERGE_PER_THREAD) {
                e_pos[cnt] = key_to_f64(s_key[r]);
                e_val[cnt] = now;
                ++cnt;
            }
            was = now;
        }
    }
    int tile_out;
    int o = block_excl_sum<int>(cnt, s_scan, tile_out);
    if (threadIdx.x < 32) {
        unsigned long long b = lookback_excl_warp(status, tile, (unsigned long long)tile_out);
        if (threadIdx.x == 0) {
            s_obase = b;
            if (tile == n_tiles - 1) *out_count = b + tile_out;
        }
    }
    __syncthreads();
    unsigned long long wbase = s_obase + o;
    for (int j = 0; j < cnt; ++j) { out_pos[wbase + j] = e_pos[j]; out_val[wbase + j] = e_val[j]; }
}

// ---- back-projection ---------------------------------------------------------

// numpy's float64 floor_divide (npy_divmod): fmod based, then corrected.
__device__ __forceinline__ double np_floor_divide(double a, double b) {
    // b == 1 (every node whose linear span equals its length): fmod(a, 1) is the
    // fractional part, (a - mod) / 1 the integer part towards zero, minus one when
    // a is negative and not integral -- i.e. exactly floor(a); the result is -0.0
    // only for a == -0.0, and callers add an integer offset anyway
    if (b == 1.0) return floor(a);
    double mod = fmod(a, b);
    double div = __ddiv_rn(__dsub_rn(a, mod), b);
    if (mod != 0.0) {
        if ((b < 0) != (mod < 0)) div = __dsub_rn(div, 1.0);
    }
    double fl;
    if (div != 0.0) {
        fl = floor(div);
        if (__dsub_rn(div, fl) > 0.5) fl = __dadd_rn(fl, 1.0);
    } else {
        fl = copysign(0.0, __ddiv_rn(a, b));
    }
    return fl;
}

__device__ __forceinline__ unsigned upper_bound_f64(const double* a, unsigned n, double x) {
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ unsigned lower_bound_f64(const double* a, unsigned n, double x) {
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        unsigned mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void backproject_count_kernel(gpc_graph_t graph, const uint8_t* __restrict__ touched,
                                         const double* __restrict__ bp, unsigned n_bp,
                                         uint32_t* __restrict__ first, uint32_t* __restrict__ count) {
    unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= graph.n_nodes) return;
    if (touched && !touched[v]) { first[v] = 0; count[v] = 1; return; }
    unsigned j0 = upper_bound_f64(bp, n_bp, graph.lin_start[v]);
    unsigned j1 = lower_bound_f64(bp, n_bp, graph.lin_end[v]);
    first[v] = j0;
    count[v] = 1 + (j1 > j0 ? j1 - j0 : 0);
}

__global__ void backproject_emit_kernel(gpc_graph_t graph, const uint8_t* __restrict__ touched,
                                        const double* __restrict__ bp, const double* __restrict__ val,
                                        const uint32_t* __restrict__ first,
                                        const uint32_t* __restrict__ count,
                                        const uint32_t* __restrict__ offs, double min_value,
                                        double value_scale, uint32_t* __restrict__ out_pos,
                                        double* __restrict__ out_val) {
    unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= graph.n_nodes) return;
    uint4 rec = node_record(graph, v);
    double node_off = (double)rec.x;
    uint32_t o = offs[v];
    if (touched && !touched[v]) {
        out_pos[o] = rec.x;
        out_val[o] = __dmul_rn(min_value, value_scale);
        return;
    }
    double ls = graph.lin_start[v], le = graph.lin_end[v];
    double scale = __ddiv_rn(__dsub_rn(le, ls), (double)rec.w);
    unsigned j0 = first[v], c = count[v];
    // entry 0: the value in force at the node start, clamped to the node start
    double t0 = j0 > 0 ? bp[j0 - 1] : 0.0;
    double v0 = j0 > 0 ? val[j0 - 1] : 0.0;
    double m0 = __dadd_rn(np_floor_divide(__dsub_rn(t0, ls), scale), node_off);
    if (!(m0 > node_off)) m0 = node_off;      // max(node_off, m0)
    out_pos[o] = (uint32_t)(long long)m0;
    out_val[o] = __dmul_rn(v0, value_scale);
    for (unsigned k = 1; k < c; ++k) {
        double t = bp[j0 + k - 1];
        double m = __dadd_rn(np_floor_divide(__dsub_rn(t, ls), scale), node_off);
        out_pos[o + k] = (uint32_t)(long long)m;
        out_val[o + k] = __dmul_rn(val[j0 + k - 1], value_scale);
    }
}

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_control_linear_track(const gpc_graph_t* graph, const int32_t* start_node,
                                        const int32_t* start_off, const uint32_t* read_index,
                                        uint64_t n_reads, const int32_t* extensions,
                                        int32_t n_extensions, int32_t fragment_length,
                                        double min_value, double* out_pos, double* out_val,
                                        uint64_t* n_out, double* x_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (n_extensions < 1 || n_extensions > MAX_WINDOWS) {
        set_error("between 1 and %d extension windows are supported", MAX_WINDOWS);
        return GPC_ERR_ARG;
    }
    if (!graph->lin_start || !graph->lin_end) {
        set_error("graph has no linear map");
        return GPC_ERR_ARG;
    }
    if (n_reads == 0) return GPC_OK;
    WindowParams wp;
    wp.n_windows = n_extensions;
    wp.min_value = min_value;
    for (int w = 0; w < MAX_WINDOWS; ++w) { wp.half[w] = 0; wp.weight[w] = 0; }
    for (int w = 0; w < n_extensions; ++w) {
        int half = extensions[w] / 2;
        if (half <= 0) { set_error("extension window too small"); return GPC_ERR_ARG; }
        wp.half[w] = (double)half;
        // controlgenerator.py:31: diffs (+-1.0) /= (extension*2/fragment_length)
        double s = (double)(half * 2) / (double)fragment_length;
        wp.weight[w] = 1.0 / s;
    }
    const unsigned n = (unsigned)n_reads;
    const unsigned nq = 2 * n_extensions;
    Scratch k0, k1, small, total, xu, wpre;
    GPC_CUDA(small.alloc(64, stream));
    GPC_CUDA(cudaMemsetAsync(small.p, 0, 64, stream));
    int* err = small.as<int>();
    GPC_CUDA(total.alloc(8, stream));
    GPC_CUDA(xu.alloc((size_t)n * 8, stream));
    GPC_CUDA(wpre.alloc((size_t)(n + 1) * 4, stream));
    unsigned long long hU = 0, last_key = 0;
    int herr = 0, not_integer = 0;
    double x_last = 0.0;
    // every node's linear span is a whole multiple of its length: all starts are integers
    // below the linear length, known before anything is projected
    const bool int_keys = (graph->flags & GPC_GRAPH_INTEGER_LINEAR) && graph->n_nodes > 0 &&
                          !getenv("GPC_FORCE_MERGE_PATH") && !getenv("GPC_CONTROL_KEYS64");
    if (int_keys) {
        GPC_CUDA(k0.alloc((size_t)n * 4, stream));
        GPC_CUDA(k1.alloc((size_t)n * 4, stream));
        GPC_PROF_BYTES("project_keys", 16.0 * n);
        project_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(*graph, start_node, start_off, read_index,
                                                                n, nullptr, k0.as<uint32_t>(), x_out, err);
        GPC_LAUNCH_CHECK();
        bool in_tmp = false;
        int bits = 1;                 // a start lies below the linear length <= size of the graph
        while (bits < 30 && ((unsigned long long)graph->size_bp >> bits)) ++bits;
        GPC_TRY((radix_sort<uint32_t, false>(k0.as<uint32_t>(), k1.as<uint32_t>(), nullptr, nullptr, n,
                                             0, bits, false, &in_tmp, stream)));
        const uint32_t* keys = in_tmp ? k1.as<uint32_t>() : k0.as<uint32_t>();
        GPC_TRY(select_if("unique_starts", UniqueStartsOp32{keys, n, xu.as<double>(), wpre.as<uint32_t>()},
                          n, total.as<unsigned long long>(), stream));
        uint32_t last32 = 0;
        GPC_READ_BACK({total.p, &hU, (unsigned)(8)}, {err, &herr, (unsigned)(4)}, {err + 1, &not_integer, (unsigned)(4)}, {keys + (n - 1), &last32, (unsigned)(4)});
        x_last = (double)last32;
        if (!herr && not_integer) {
            set_error("linear map flagged all-integer, but a read start projects to a non-integer");
            return GPC_ERR_INTERNAL;
        }
    } else {
        GPC_CUDA(k0.alloc((size_t)n * 8, stream));
        GPC_CUDA(k1.alloc((size_t)n * 8, stream));
        GPC_PROF_BYTES("project_keys", 20.0 * n);
        project_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(*graph, start_node, start_off, read_index,
                                                                n, k0.as<unsigned long long>(), nullptr,
                                                                x_out, err);
        GPC_LAUNCH_CHECK();
        bool in_tmp = false;
        GPC_TRY((radix_sort<unsigned long long, false>(k0.as<unsigned long long>(),
                                                       k1.as<unsigned long long>(), nullptr, nullptr, n,
                                                       0, 64, true, &in_tmp, stream)));
        const unsigned long long* keys = in_tmp ? k1.as<unsigned long long>() : k0.as<unsigned long long>();
        // distinct starts; largest start = last sorted key; integrality of the starts
        GPC_TRY(select_if("unique_starts",
                          UniqueStartsOp{keys, n, xu.as<double>(), wpre.as<uint32_t>(), err + 1}, n,
                          total.as<unsigned long long>(), stream));
        GPC_READ_BACK({total.p, &hU, (unsigned)(8)}, {err, &herr, (unsigned)(4)}, {err + 1, &not_integer, (unsigned)(4)}, {keys + (n - 1), &last_key, (unsigned)(8)});
        x_last = key_to_f64(last_key);
    }
    if (herr) {
        set_error("Interval has node(s) not part of graph/pileup");
        return herr;
    }
    const unsigned U = (unsigned)hU;
    // integer coordinates (no node with a non-integer scale was hit): dense path -- when the
    // reads are dense enough.  It works per linear POSITION (7 ps each, measured), the tile
    // path below per EVENT (55 ps each): below ~0.12 events per position -- whole-genome
    // depth, 0.075 on the human workload -- the event-driven tiles are the faster ones.
    long long e_max = 0;
    for (int w = 0; w < n_extensions; ++w) e_max = (long long)wp.half[w] > e_max ? (long long)wp.half[w] : e_max;
    const long long origin = -e_max;
    const long long span = (long long)x_last + 2 * e_max + 1;
    double dense_min = 0.12;
    if (const char* e = getenv("GPC_LINEAR_DENSE_MIN")) dense_min = atof(e);
    const bool dense_enough = (double)nq * (double)U >= dense_min * (double)span;
    if (!not_integer && dense_enough && !getenv("GPC_FORCE_MERGE_PATH")) {
        const unsigned n_dense = (unsigned)((span + DENSE_W - 1) / DENSE_W);
        Scratch dsc, dbounds, tbase, tcnt, toff, tpos, tval, ttotal;
        GPC_CUDA(dsc.alloc(64, stream));
        GPC_CUDA(cudaMemsetAsync(dsc.p, 0, 64, stream));
        unsigned long long* dcount = dsc.as<unsigned long long>();
        const size_t cap = (size_t)nq * U;                      // every event could be a breakpoint
        GPC_CUDA(tbase.alloc((size_t)n_dense * 4, stream));
        GPC_CUDA(tcnt.alloc((size_t)n_dense * 4, stream));
        GPC_CUDA(toff.alloc((size_t)n_dense * 4, stream));
        GPC_CUDA(tpos.alloc(cap * 8, stream));
        GPC_CUDA(tval.alloc(cap * 8, stream));
        GPC_CUDA(ttotal.alloc(8, stream));
        GPC_CUDA(dbounds.alloc((size_t)(n_dense + 1) * nq * 4, stream));
        GPC_PROF("dense_bounds");
        dense_bounds_kernel<<<div_up((size_t)(n_dense + 1) * nq, 256), 256, 0, stream>>>(
            xu.as<double>(), U, wp, origin, n_dense, dbounds.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        GPC_PROF_BYTES("linear_dense", 12.0 * nq * U);
        linear_dense_kernel<<<n_dense, DENSE_THREADS, 0, stream>>>(
            xu.as<double>(), wpre.as<uint32_t>(), U, wp, origin, n_dense, dbounds.as<uint32_t>(),
            tpos.as<double>(), tval.as<double>(), tbase.as<uint32_t>(), tcnt.as<uint32_t>(), dcount);
        GPC_LAUNCH_CHECK();
        GPC_TRY(exclusive_scan<uint32_t>(tcnt.as<uint32_t>(), toff.as<uint32_t>(), n_dense,
                                         ttotal.as<unsigned long long>(), stream));
        GPC_PROF("dense_gather");
        dense_gather_kernel<<<div_up((size_t)n_dense * 32, 256), 256, 0, stream>>>(
            tpos.as<double>(), tval.as<double>(), tbase.as<uint32_t>(), tcnt.as<uint32_t>(),
            toff.as<uint32_t>(), n_dense, out_pos, out_val);
        GPC_LAUNCH_CHECK();
        unsigned long long hd = 0;
        GPC_READ_BACK({dcount, &hd, (unsigned)(8)});
        *n_out = hd;
        prof_add_bytes("linear_dense", 16.0 * (double)hd);
        return GPC_OK;
    }
    // splitters: every MERGE_T-th element of every view, sorted
    const unsigned per_view = div_up(U, MERGE_T);
    const unsigned ns = per_view * nq;
    Scratch sp0, sp1, bounds, sc;
    GPC_CUDA(sp0.alloc((size_t)ns * 8, stream));
    GPC_CUDA(sp1.alloc((size_t)ns * 8, stream));
    GPC_PROF("splitters");
    splitters_kernel<<<div_up(ns, 256), 256, 0, stream>>>(xu.as<double>(), U, wp, per_view,
                                                          sp0.as<unsigned long long>());
    GPC_LAUNCH_CHECK();
    bool sp_tmp = false;
    GPC_TRY((radix_sort<unsigned long long, false>(sp0.as<unsigned long long>(),
                                                   sp1.as<unsigned long long>(), nullptr, nullptr,
                                                   ns, 0, 64, true, &sp_tmp, stream)));
    const unsigned long long* sp = sp_tmp ? sp1.as<unsigned long long>() : sp0.as<unsigned long long>();
    const unsigned n_bounds = ns + 2;
    Scratch brank, pick;
    GPC_CUDA(bounds.alloc((size_t)n_bounds * nq * 4, stream));
    GPC_CUDA(brank.alloc((size_t)n_bounds * 8, stream));
    GPC_PROF("tile_bounds");
    tile_bounds_kernel<<<div_up((size_t)n_bounds * nq, 256), 256, 0, stream>>>(
        xu.as<double>(), U, wp, sp, ns, bounds.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("boundary_rank");
    boundary_rank_kernel<<<div_up(n_bounds, 256), 256, 0, stream>>>(bounds.as<uint32_t>(), n_bounds,
                                                                  nq, brank.as<unsigned long long>());
    GPC_LAUNCH_CHECK();
    const unsigned long long total_ev = (unsigned long long)nq * U;
    const unsigned long long stride = MERGE_CAP - (unsigned long long)nq * (MERGE_T + 2);
    const unsigned n_tiles = (unsigned)((total_ev + stride - 1) / stride);
    GPC_CUDA(pick.alloc((size_t)(n_tiles + 1) * 4, stream));
    GPC_PROF("pick_tiles");
    pick_tiles_kernel<<<div_up(n_tiles + 1, 256), 256, 0, stream>>>(
        brank.as<unsigned long long>(), n_bounds, n_tiles, stride, pick.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    size_t bytes = (size_t)n_tiles * 8 + 64;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    unsigned long long* count = sc.as<unsigned long long>();
    unsigned* ticket = reinterpret_cast<unsigned*>(count + 1);
    unsigned long long* status = count + 8;
    static bool smem_set = false;
    if (!smem_set) {
        GPC_CUDA(cudaFuncSetAttribute(linear_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)MERGE_SMEM));
        smem_set = true;
    }
    GPC_PROF_BYTES("linear_tiles", (8.0 + 4.0) * nq * U);
    linear_tiles_kernel<<<n_tiles, MERGE_THREADS, MERGE_SMEM, stream>>>(
        xu.as<double>(), wpre.as<uint32_t>(), wp, bounds.as<uint32_t>(), pick.as<uint32_t>(), n_tiles,
        out_pos, out_val, status, ticket, count, err);
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    GPC_READ_BACK({count, &h, (unsigned)(8)}, {err, &herr, (unsigned)(4)});
    if (herr) {
        set_error("linear track: merge tile overflow");
        return herr;
    }
    *n_out = h;
    prof_add_bytes("linear_tiles", 16.0 * (double)h);
    return GPC_OK;
}

extern "C" int gpc_control_backproject(const gpc_graph_t* graph, const uint8_t* touched,
                                       const double* bp, const double* val, uint64_t n_bp,
                                       double min_value, double value_scale, uint32_t* out_pos,
                                       double* out_val, uint64_t out_capacity, uint64_t* n_out,
                                       void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (!graph->lin_start || !graph->lin_end) {
        set_error("graph has no linear map");
        return GPC_ERR_ARG;
    }
    const unsigned n = graph->n_nodes;
    Scratch first, count, offs, total;
    GPC_CUDA(first.alloc((size_t)n * 4, stream));
    GPC_CUDA(count.alloc((size_t)n * 4, stream));
    GPC_CUDA(offs.alloc((size_t)n * 4, stream));
    GPC_CUDA(total.alloc(8, stream));
    GPC_PROF("backproject_count");
    backproject_count_kernel<<<div_up(n, 256), 256, 0, stream>>>(
        *graph, touched, bp, (unsigned)n_bp, first.as<uint32_t>(), count.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    GPC_TRY(exclusive_scan<uint32_t>(count.as<uint32_t>(), offs.as<uint32_t>(), n,
                                     total.as<unsigned long long>(), stream));
    unsigned long long h = 0;
    GPC_READ_BACK({total.p, &h, (unsigned)(8)});
    *n_out = h;
    if (h > out_capacity) return GPC_ERR_CAPACITY;
    GPC_PROF("backproject_emit");
    backproject_emit_kernel<<<div_up(n, 256), 256, 0, stream>>>(
        *graph, touched, bp, val, first.as<uint32_t>(), count.as<uint32_t>(),
        offs.as<uint32_t>(), min_value, value_scale, out_pos, out_val);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
