// Run-length tracks from ARBITRARY event lists on sm_100a: the generic form of T1/T2.
//
// The callpeaks path makes its tracks from packed +-1 events (csrc/sample.cu) or a
// difference array; a caller of the reference's container API can hand SparseDiffs any
// (index, diff) list -- float diffs, unsorted, repeated indices -- and these entries do
// what the reference's numpy statements do with it:
//   SparseDiffs.get_sparse_values     sparsediffs.py:137-141  (stable argsort, cumsum,
//                                     SparseValues._sanitize :13-20)
//   SparseDiffs.from_pileup           sparsediffs.py:180-190  (its event list is the input)
//   SparseDiffs.maximum               sparsediffs.py:143-159
//   SparseDiffs.apply_binary_func     sparsediffs.py:208-224  (the merge and the two
//                                     running sums; the callable itself is the caller's)
//   SparseDiffs.clip_min              sparsediffs.py:130-135  (maximum with a constant)
//
// Pipeline: 32-bit keys + source index -> LSD radix sort of pairs (stable, like the
// reference's mergesort) -> gather of the diffs in sorted order (two columns when two
// tracks are merged: a track contributes 0 where the entry is the other track's, as
// sparsediffs.py:149-153 does) -> float64 running sums by reduce-then-scan (tile sums, one
// block over the tile sums, independent tiles) -> compaction of the last entry of every
// index group -> optional elementwise combination -> compaction of entries that repeat
// their predecessor's value.
//
// Numerics: np.cumsum adds left to right; a parallel scan associates differently.  Sums
// of integer-valued (or any exactly representable, e.g. dyadic) steps are exact in both and
// therefore bit-identical; for other floats the two agree to rounding (a few ulp of the
// largest partial sum), and the reference's own sums already depend on the unspecified
// order its unstable argsort leaves equal indices in (sparsediffs.py:174,187).
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace gpc {

constexpr int TD_THREADS = 256;
constexpr int TD_ITEMS = 8;
constexpr int TD_TILE = TD_THREADS * TD_ITEMS;

__global__ void td_keys_kernel(const long long* __restrict__ idx_a, unsigned na,
                               const long long* __restrict__ idx_b, unsigned nb,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ src,
                               int* __restrict__ bad) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na + nb) return;
    long long v = i < na ? idx_a[i] : idx_b[i - na];
    if (v < 0 || v >= 0xffffffffll) {          // positions are uint32, 0xffffffff is a sentinel
        *bad = 1;
        v = 0;
    }
    keys[i] = (uint32_t)v;
    src[i] = i;
}

// diffs in sorted order.  One track (get_sparse_values, sparsediffs.py:138-139): every diff
// follows its index, diffs[args].  Two tracks (maximum / apply_binary_func, :149-153): the
// reference lays diffs_a out over the slots of track a IN THEIR GIVEN ORDER -- the k-th slot
// of track a in merged order takes diffs_a[k] -- which is the same thing for index-sorted
// lists (the stable sort keeps a track's entries in order) and is reproduced literally for
// the others: rank_a[i] = slots of track a before slot i.
__global__ void td_slot_flags_kernel(const uint32_t* __restrict__ src, unsigned n, unsigned na,
                                     uint32_t* __restrict__ is_a) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_a[i] = src[i] < na ? 1u : 0u;
}
__global__ void td_gather_kernel(const uint32_t* __restrict__ src, unsigned n,
                                 const double* __restrict__ diff_a, unsigned na,
                                 const double* __restrict__ diff_b, const uint32_t* __restrict__ rank_a,
                                 double* __restrict__ col_a, double* __restrict__ col_b) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned s = src[i];
    if (!col_b) {
        col_a[i] = diff_a[s];
        return;
    }
    const unsigned k = rank_a[i];
    col_a[i] = s < na ? diff_a[k] : 0.0;
    col_b[i] = s >= na ? diff_b[i - k] : 0.0;
}

// ---- float64 running sums: reduce, then scan --------------------------------------------
__global__ void __launch_bounds__(TD_THREADS)
td_tile_sums_kernel(const double* __restrict__ col_a, const double* __restrict__ col_b, unsigned n,
                    double* __restrict__ tsum /* [2 * tiles] */) {
    __shared__ double s_red[2][TD_THREADS / 32];
    const unsigned base = blockIdx.x * TD_TILE + threadIdx.x * TD_ITEMS;
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int j = 0; j < TD_ITEMS; ++j) {
        const unsigned i = base + j;
        if (i < n) {
            sa += col_a[i];
            if (col_b) sb += col_b[i];
        }
    }
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    if (lane_id() == 0) { s_red[0][warp_id()] = sa; s_red[1][warp_id()] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int w = 0; w < TD_THREADS / 32; ++w) { ta += s_red[0][w]; tb += s_red[1][w]; }
        tsum[2 * blockIdx.x] = ta;
        tsum[2 * blockIdx.x + 1] = tb;
    }
}

// one block: tile sums -> exclusive prefixes, in place, both columns
constexpr int TD_PREFIX_THREADS = 1024;
__global__ void __launch_bounds__(TD_PREFIX_THREADS)
td_tile_prefix_kernel(double* __restrict__ tsum, unsigned tiles) {
    __shared__ double s_scan[33];
    for (int col = 0; col < 2; ++col) {
        double run = 0.0;
        for (unsigned first = 0; first < tiles; first += TD_PREFIX_THREADS) {
            const unsigned i = first + threadIdx.x;
            const double v = i < tiles ? tsum[2 * i + col] : 0.0;
            double total;
            const double ex = block_excl_sum<double>(v, s_scan, total);
            if (i < tiles) tsum[2 * i + col] = run + ex;
            run += total;
        }
    }
}

// inclusive running sums in place; every tile starts from its prefix
__global__ void __launch_bounds__(TD_THREADS)
td_scan_kernel(double* __restrict__ col_a, double* __restrict__ col_b, unsigned n,
               const double* __restrict__ tprefix) {
    __shared__ double s_scan[33];
    const unsigned base = blockIdx.x * TD_TILE + threadIdx.x * TD_ITEMS;
    for (int col = 0; col < 2; ++col) {
        double* c = col == 0 ? col_a : col_b;
        if (!c) break;                                   // (uniform over the block)
        double v[TD_ITEMS], t = 0.0;
#pragma unroll
        for (int j = 0; j < TD_ITEMS; ++j) {
            v[j] = base + j < n ? c[base + j] : 0.0;
            t += v[j];
        }
        double total;
        const double ex = block_excl_sum<double>(t, s_scan, total);
        double run = tprefix[2 * blockIdx.x + col] + ex;
#pragma unroll
        for (int j = 0; j < TD_ITEMS; ++j) {
            run += v[j];
            if (base + j < n) c[base + j] = run;
        }
    }
}

// ---- compactions -----------------------------------------------------------------------
// last entry of every index group (SparseValues._sanitize first half / the unique_mask of
// sparsediffs.py:146,211)
struct GroupLastOp {
    const uint32_t* keys;
    const double* col_a;
    const double* col_b;
    unsigned n;
    uint32_t* out_pos;
    double* out_a;
    double* out_b;
    __device__ __forceinline__ bool keep(unsigned i) const { return i + 1 == n || keys[i + 1] != keys[i]; }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        out_pos[w] = keys[i];
        out_a[w] = col_a[i];
        if (out_b) out_b[w] = col_b[i];
    }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

// entries that repeat their predecessor's value are dropped (SparseValues._sanitize second
// half); drop_leading_zero: the first entry goes too when its value is 0 -- that is what
// SparseDiffs._sanitize (:161-165) leaves of ediff1d(values, to_begin=values[0])
struct ChangeOp {
    const uint32_t* in_pos;
    const double* in_v;
    uint32_t* out_pos;
    double* out_v;
    int drop_leading_zero;
    __device__ __forceinline__ bool keep(unsigned i) const {
        if (i == 0) return !(drop_leading_zero && in_v[0] == 0.0);
        return in_v[i] != in_v[i - 1];
    }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        out_pos[w] = in_pos[i];
        out_v[w] = in_v[i];
    }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

// a[i] = op(a[i], b[i]); NaN propagates as in numpy (np.max, np.min, +, -, *)
__global__ void td_combine_kernel(double* __restrict__ a, const double* __restrict__ b, unsigned n_max,
                                  const unsigned long long* __restrict__ n_dev, int op) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = (unsigned)min((unsigned long long)n_max, *n_dev);
    if (i >= n) return;
    const double x = a[i], y = b[i];
    double r;
    switch (op) {
        case GPC_TRACK_MAX: r = (x > y || x != x) ? x : y; break;
        case GPC_TRACK_MIN: r = (x < y || x != x) ? x : y; break;
        case GPC_TRACK_ADD: r = __dadd_rn(x, y); break;
        case GPC_TRACK_SUB: r = __dsub_rn(x, y); break;
        default: r = __dmul_rn(x, y); break;
    }
    a[i] = r;
}

// Union of the positions of one or two event lists with each track's running sum there
// (the last entry of every index group): pos / val_a / val_b need room for na + nb entries,
// *count_dev receives their number, *bad_dev is set when an index is out of range.
static int union_values(const int64_t* idx_a, const double* diff_a, uint64_t na, const int64_t* idx_b,
                        const double* diff_b, uint64_t nb, uint32_t* pos, double* val_a, double* val_b,
                        unsigned long long* count_dev, int* bad_dev, cudaStream_t stream) {
    const uint64_t n64 = na + nb;
    GPC_CUDA(cudaMemsetAsync(bad_dev, 0, sizeof(int), stream));
    if (n64 == 0) {
        GPC_CUDA(cudaMemsetAsync(count_dev, 0, 8, stream));
        return GPC_OK;
    }
    if (n64 >= (1ull << 30)) {
        set_error("track entries: %llu events, at most 2^30 - 1", (unsigned long long)n64);
        return GPC_ERR_ARG;
    }
    const unsigned n = (unsigned)n64;
    const bool two = val_b != nullptr;
    Scratch keys, keys_tmp, src, src_tmp, col_a, col_b, tsum;
    GPC_CUDA(keys.alloc((size_t)n * 4, stream));
    GPC_CUDA(keys_tmp.alloc((size_t)n * 4, stream));
    GPC_CUDA(src.alloc((size_t)n * 4, stream));
    GPC_CUDA(src_tmp.alloc((size_t)n * 4, stream));
    GPC_CUDA(col_a.alloc((size_t)n * 8, stream));
    if (two) GPC_CUDA(col_b.alloc((size_t)n * 8, stream));
    const unsigned tiles = div_up(n, TD_TILE);
    GPC_CUDA(tsum.alloc((size_t)tiles * 16, stream));

    GPC_PROF("td_keys");
    td_keys_kernel<<<div_up(n, 256), 256, 0, stream>>>(
        reinterpret_cast<const long long*>(idx_a), (unsigned)na, reinterpret_cast<const long long*>(idx_b),
        (unsigned)nb, keys.as<uint32_t>(), src.as<uint32_t>(), bad_dev);
    GPC_LAUNCH_CHECK();
    uint32_t* k = keys.as<uint32_t>();
    uint32_t* s = src.as<uint32_t>();
    if (n > 1) {
        bool in_tmp = false;
        GPC_TRY((radix_sort<uint32_t, true>(keys.as<uint32_t>(), keys_tmp.as<uint32_t>(), src.as<uint32_t>(),
                                            src_tmp.as<uint32_t>(), n, 0, 32, false, &in_tmp, stream)));
        if (in_tmp) { k = keys_tmp.as<uint32_t>(); s = src_tmp.as<uint32_t>(); }
    }
    double* ca = col_a.as<double>();
    double* cb = two ? col_b.as<double>() : nullptr;
    Scratch is_a, rank_a;
    if (two) {
        GPC_CUDA(is_a.alloc((size_t)n * 4, stream));
        GPC_CUDA(rank_a.alloc((size_t)n * 4, stream));
        GPC_PROF("td_slot_flags");
        td_slot_flags_kernel<<<div_up(n, 256), 256, 0, stream>>>(s, n, (unsigned)na, is_a.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        GPC_TRY(exclusive_scan<uint32_t>(is_a.as<uint32_t>(), rank_a.as<uint32_t>(), n, nullptr, stream));
    }
    GPC_PROF("td_gather");
    td_gather_kernel<<<div_up(n, 256), 256, 0, stream>>>(s, n, diff_a, (unsigned)na, diff_b,
                                                          two ? rank_a.as<uint32_t>() : nullptr, ca, cb);
    GPC_LAUNCH_CHECK();
    GPC_PROF("td_tile_sums");
    td_tile_sums_kernel<<<tiles, TD_THREADS, 0, stream>>>(ca, cb, n, tsum.as<double>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("td_tile_prefix");
    td_tile_prefix_kernel<<<1, TD_PREFIX_THREADS, 0, stream>>>(tsum.as<double>(), tiles);
    GPC_LAUNCH_CHECK();
    GPC_PROF("td_scan");
    td_scan_kernel<<<tiles, TD_THREADS, 0, stream>>>(ca, cb, n, tsum.as<double>());
    GPC_LAUNCH_CHECK();
    GroupLastOp op{k, ca, cb, n, pos, val_a, val_b};
    GPC_TRY(select_if("td_group_last", op, n, count_dev, stream));
    return GPC_OK;
}

static int finish_counts(unsigned long long* count_dev, int* bad_dev, uint64_t* n_out, cudaStream_t stream) {
    unsigned long long h_count = 0;
    int h_bad = 0;
    GPC_READ_BACK({count_dev, &h_count, 8u}, {bad_dev, &h_bad, 4u});
    *n_out = h_count;
    if (h_bad) {
        set_error("track entries: an index is negative or does not fit a uint32 position");
        return GPC_ERR_ARG;
    }
    return GPC_OK;
}

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_diffs_to_runs(const int64_t* indices, const double* diffs, uint64_t n, uint32_t* out_pos,
                                 double* out_val, uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch counts, pos1, val1;
    GPC_CUDA(counts.alloc(32, stream));
    unsigned long long* count1 = counts.as<unsigned long long>();
    unsigned long long* count2 = count1 + 1;
    int* bad = reinterpret_cast<int*>(count1 + 2);
    GPC_CUDA(pos1.alloc((size_t)n * 4, stream));
    GPC_CUDA(val1.alloc((size_t)n * 8, stream));
    GPC_TRY(union_values(indices, diffs, n, nullptr, nullptr, 0, pos1.as<uint32_t>(), val1.as<double>(), nullptr,
                         count1, bad, stream));
    ChangeOp op{pos1.as<uint32_t>(), val1.as<double>(), out_pos, out_val, 0};
    GPC_TRY(select_if("td_changes", op, n, count2, stream, count1));
    return finish_counts(count2, bad, n_out, stream);
}

extern "C" int gpc_diffs_merge(const int64_t* idx_a, const double* diff_a, uint64_t n_a, const int64_t* idx_b,
                               const double* diff_b, uint64_t n_b, uint32_t* out_pos, double* out_a,
                               double* out_b, uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_b) {
        set_error("gpc_diffs_merge: out_b is NULL");
        return GPC_ERR_ARG;
    }
    Scratch counts;
    GPC_CUDA(counts.alloc(32, stream));
    unsigned long long* count = counts.as<unsigned long long>();
    int* bad = reinterpret_cast<int*>(count + 2);
    GPC_TRY(union_values(idx_a, diff_a, n_a, idx_b, diff_b, n_b, out_pos, out_a, out_b, count, bad, stream));
    return finish_counts(count, bad, n_out, stream);
}

extern "C" int gpc_diffs_combine(const int64_t* idx_a, const double* diff_a, uint64_t n_a,
                                 const int64_t* idx_b, const double* diff_b, uint64_t n_b, int32_t op,
                                 int32_t drop_leading_zero, uint32_t* out_pos, double* out_val,
                                 uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (op < GPC_TRACK_MAX || op > GPC_TRACK_MUL) {
        set_error("gpc_diffs_combine: unknown operation %d", (int)op);
        return GPC_ERR_ARG;
    }
    const uint64_t n = n_a + n_b;
    Scratch counts, pos1, va, vb;
    GPC_CUDA(counts.alloc(32, stream));
    unsigned long long* count1 = counts.as<unsigned long long>();
    unsigned long long* count2 = count1 + 1;
    int* bad = reinterpret_cast<int*>(count1 + 2);
    GPC_CUDA(pos1.alloc((size_t)n * 4, stream));
    GPC_CUDA(va.alloc((size_t)n * 8, stream));
    GPC_CUDA(vb.alloc((size_t)n * 8, stream));
    GPC_TRY(union_values(idx_a, diff_a, n_a, idx_b, diff_b, n_b, pos1.as<uint32_t>(), va.as<double>(),
                         vb.as<double>(), count1, bad, stream));
    if (n > 0 && n < (1ull << 30)) {
        GPC_PROF("td_combine");
        td_combine_kernel<<<div_up(n, 256), 256, 0, stream>>>(va.as<double>(), vb.as<double>(), (unsigned)n,
                                                               count1, (int)op);
        GPC_LAUNCH_CHECK();
    }
    ChangeOp cop{pos1.as<uint32_t>(), va.as<double>(), out_pos, out_val, drop_leading_zero ? 1 : 0};
    GPC_TRY(select_if("td_changes", cop, n, count2, stream, count1));
    return finish_counts(count2, bad, n_out, stream);
}
