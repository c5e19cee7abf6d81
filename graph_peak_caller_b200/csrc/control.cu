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
// the reference's exact float64 operations.  Each extension window w turns a
// start x into a +1 at x-e_w and a -1 at x+e_w; all windows' events are sorted
// once (64-bit order-preserving keys) and one chained scan keeps an integer
// count per window.  The linear lambda is
//     max( min_value , max_w count_w * weight_w )
// (before the first event of window 0 the clip and window 0 do not apply: that
// is what clip_min on the first track followed by maximum() does).  Counts are
// integers, so lambda = count * weight is exact where the reference accumulates
// +-weight steps in float64; the two agree to ~1e-13 relative.
// Back-projection to graph positions is two binary searches per node plus a
// segmented copy, with numpy's float floor-division replicated bit for bit.
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
__global__ void project_events_kernel(gpc_graph_t graph, const int32_t* __restrict__ start_node,
                                      const int32_t* __restrict__ start_off,
                                      const uint32_t* __restrict__ read_index, unsigned n_reads,
                                      WindowParams wp, unsigned long long* __restrict__ keys,
                                      uint32_t* __restrict__ tags, double* __restrict__ x_out,
                                      int* __restrict__ err) {
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
    for (int w = 0; w < wp.n_windows; ++w) {
        size_t o = ((size_t)w * n_reads + i) * 2;
        keys[o] = f64_to_key(__dsub_rn(x, wp.half[w]));
        tags[o] = (uint32_t)(w << 1);
        keys[o + 1] = f64_to_key(__dadd_rn(x, wp.half[w]));
        tags[o + 1] = (uint32_t)(w << 1) | 1u;
    }
}

// ---- sorted window events -> linear lambda track -----------------------------
constexpr int LIN_THREADS = 256;
constexpr int LIN_ITEMS = 4;
constexpr int LIN_TILE = LIN_THREADS * LIN_ITEMS;
constexpr int NC = 4;            // count per window (3) + "window-0 starts seen"

struct LinAgg {
    int tot[NC];
    int seg[NC];
    int has_head;
    int pad[3];
};

struct SegN { int head; int c[NC]; };
__device__ __forceinline__ SegN segn_combine(const SegN& x, const SegN& y) {
    if (y.head) return y;
    SegN r;
    r.head = x.head;
#pragma unroll
    for (int k = 0; k < NC; ++k) r.c[k] = x.c[k] + y.c[k];
    return r;
}

__device__ __forceinline__ void tag_delta(uint32_t tag, int d[NC]) {
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = 0;
    int w = tag >> 1;
    bool is_end = tag & 1u;
    d[w] = is_end ? -1 : 1;
    if (w == 0 && !is_end) d[3] = 1;
}

__device__ __forceinline__ double lambda_of(const int c[NC], const WindowParams& wp) {
    double best = 0.0;
    bool clipped = c[3] > 0;
    for (int w = clipped ? 0 : 1; w < wp.n_windows; ++w)
        best = fmax(best, __dmul_rn((double)c[w], wp.weight[w]));
    if (clipped) best = fmax(best, wp.min_value);
    return best;
}

__global__ void __launch_bounds__(LIN_THREADS)
linear_track_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ tags,
                    unsigned n, WindowParams wp, double* __restrict__ out_pos,
                    double* __restrict__ out_val, LinAgg* __restrict__ aggs,
                    LinAgg* __restrict__ prefixes, unsigned* __restrict__ st1,
                    unsigned long long* __restrict__ st2, unsigned* __restrict__ ticket,
                    unsigned long long* __restrict__ out_count) {
    __shared__ unsigned s_slot;
    __shared__ int s_scan[33];
    __shared__ SegN s_seg[LIN_THREADS / 32];
    __shared__ LinAgg s_carry;
    __shared__ unsigned long long s_obase;
    const unsigned tile = claim_tile(ticket, &s_slot);
    const unsigned base = tile * LIN_TILE + threadIdx.x * LIN_ITEMS;

    unsigned long long key[LIN_ITEMS + 1];
    uint32_t tag[LIN_ITEMS];
#pragma unroll
    for (int j = 0; j <= LIN_ITEMS; ++j)
        key[j] = (base + j < n) ? keys[base + j] : ~0ull;
#pragma unroll
    for (int j = 0; j < LIN_ITEMS; ++j) tag[j] = (base + j < n) ? tags[base + j] : 0u;
    const unsigned long long prev_key = (base > 0 && base - 1 < n) ? keys[base - 1] : ~0ull;

    int tot[NC] = {0, 0, 0, 0};
    SegN seg;
    seg.head = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) seg.c[k] = 0;
    unsigned long long pk = prev_key;
#pragma unroll
    for (int j = 0; j < LIN_ITEMS; ++j) {
        if (base + j < n) {
            SegN e;
            tag_delta(tag[j], e.c);
            e.head = (base + j == 0 || key[j] != pk) ? 1 : 0;
            seg = segn_combine(seg, e);
#pragma unroll
            for (int k = 0; k < NC; ++k) tot[k] += e.c[k];
            pk = key[j];
        }
    }
    // exclusive prefix of the totals over the threads of the tile
    int ex[NC], tile_tot[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) ex[k] = block_excl_sum<int>(tot[k], s_scan, tile_tot[k]);
    // exclusive segmented prefix over the threads
    SegN incl = seg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        SegN o;
        o.head = __shfl_up_sync(FULL, incl.head, d);
#pragma unroll
        for (int k = 0; k < NC; ++k) o.c[k] = __shfl_up_sync(FULL, incl.c[k], d);
        if (lane_id() >= (unsigned)d) incl = segn_combine(o, incl);
    }
    if (lane_id() == 31) s_seg[warp_id()] = incl;
    __syncthreads();
    SegN carry;
    carry.head = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) carry.c[k] = 0;
    for (unsigned w = 0; w < warp_id(); ++w) carry = segn_combine(carry, s_seg[w]);
    {
        SegN up;
        up.head = __shfl_up_sync(FULL, incl.head, 1);
#pragma unroll
        for (int k = 0; k < NC; ++k) up.c[k] = __shfl_up_sync(FULL, incl.c[k], 1);
        if (lane_id() > 0) carry = segn_combine(carry, up);
    }
    if (threadIdx.x == LIN_THREADS - 1) {
        SegN all = segn_combine(carry, seg);
        LinAgg mine, in;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            mine.tot[k] = tile_tot[k];
            mine.seg[k] = all.c[k];
            in.tot[k] = 0;
            in.seg[k] = 0;
        }
        mine.has_head = all.head;
        in.has_head = 0;
        volatile unsigned* vs = st1;
        if (tile > 0) {
            aggs[tile] = mine;
            __threadfence();
            vs[tile] = 1;
            bool seg_closed = false;
            int t = (int)tile - 1;
            while (true) {
                unsigned s = vs[t];
                if (s == 0) continue;
                __threadfence();
                LinAgg a = (s == 2) ? prefixes[t] : aggs[t];
#pragma unroll
                for (int k = 0; k < NC; ++k) in.tot[k] += a.tot[k];
                if (!seg_closed) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) in.seg[k] += a.seg[k];
                    if (a.has_head) seg_closed = true;
                }
                if (s == 2) break;
                --t;
            }
        }
        LinAgg pre;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            pre.tot[k] = in.tot[k] + mine.tot[k];
            pre.seg[k] = mine.has_head ? mine.seg[k] : in.seg[k] + mine.seg[k];
        }
        pre.has_head = 1;
        prefixes[tile] = pre;
        __threadfence();
        vs[tile] = 2;
        s_carry = in;
    }
    __syncthreads();
    const LinAgg tin = s_carry;
    int run[NC];
    SegN sg;
    sg.head = carry.head;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        run[k] = tin.tot[k] + ex[k];
        sg.c[k] = carry.head ? carry.c[k] : carry.c[k] + tin.seg[k];
    }
    double e_pos[LIN_ITEMS], e_val[LIN_ITEMS];
    int cnt = 0;
    pk = prev_key;
#pragma unroll
    for (int j = 0; j < LIN_ITEMS; ++j) {
        if (base + j < n) {
            int d[NC];
            tag_delta(tag[j], d);
            bool head = (base + j == 0) || key[j] != pk;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (head) sg.c[k] = 0;
                sg.c[k] += d[k];
                run[k] += d[k];
            }
            bool last = (base + j + 1 >= n) || key[j + 1] != key[j];
            if (last) {
                int before[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) before[k] = run[k] - sg.c[k];
                double now = lambda_of(run, wp), was = lambda_of(before, wp);
                if (now != was) {
                    e_pos[cnt] = key_to_f64(key[j]);
                    e_val[cnt] = now;
                    ++cnt;
                }
            }
            pk = key[j];
        }
    }
    int tile_out;
    int o = block_excl_sum<int>(cnt, s_scan, tile_out);
    if (threadIdx.x == 0) {
        s_obase = lookback_excl(st2, tile, (unsigned long long)tile_out);
        if (tile == gridDim.x - 1) *out_count = s_obase + tile_out;
    }
    __syncthreads();
    unsigned long long wbase = s_obase + o;
#pragma unroll
    for (int j = 0; j < LIN_ITEMS; ++j)
        if (j < cnt) { out_pos[wbase + j] = e_pos[j]; out_val[wbase + j] = e_val[j]; }
}

// ---- back-projection ---------------------------------------------------------

// numpy's float64 floor_divide (npy_divmod): fmod based, then corrected.
__device__ __forceinline__ double np_floor_divide(double a, double b) {
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
    const size_t n_ev = (size_t)n_reads * 2 * n_extensions;
    Scratch k0, k1, t0, t1, small;
    GPC_CUDA(k0.alloc(n_ev * 8, stream));
    GPC_CUDA(k1.alloc(n_ev * 8, stream));
    GPC_CUDA(t0.alloc(n_ev * 4, stream));
    GPC_CUDA(t1.alloc(n_ev * 4, stream));
    GPC_CUDA(small.alloc(64, stream));
    GPC_CUDA(cudaMemsetAsync(small.p, 0, 64, stream));
    int* err = small.as<int>();
    GPC_PROF_BYTES("project_events", 12.0 * n_reads + 12.0 * n_ev);
    project_events_kernel<<<div_up(n_reads, 256), 256, 0, stream>>>(
        *graph, start_node, start_off, read_index, (unsigned)n_reads, wp,
        k0.as<unsigned long long>(), t0.as<uint32_t>(), x_out, err);
    GPC_LAUNCH_CHECK();
    bool in_tmp = false;
    GPC_TRY((radix_sort<unsigned long long, true>(k0.as<unsigned long long>(),
                                                  k1.as<unsigned long long>(), t0.as<uint32_t>(),
                                                  t1.as<uint32_t>(), n_ev, 0, 64, true, &in_tmp,
                                                  stream)));
    const unsigned long long* keys = in_tmp ? k1.as<unsigned long long>() : k0.as<unsigned long long>();
    const uint32_t* tags = in_tmp ? t1.as<uint32_t>() : t0.as<uint32_t>();
    unsigned tiles = div_up(n_ev, LIN_TILE);
    size_t bytes = (size_t)tiles * (2 * sizeof(LinAgg) + 4 + 8) + 64;
    Scratch sc;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    char* p = sc.as<char>();
    unsigned long long* count = reinterpret_cast<unsigned long long*>(p);
    unsigned* ticket = reinterpret_cast<unsigned*>(p + 16);
    LinAgg* aggs = reinterpret_cast<LinAgg*>(p + 64);
    LinAgg* prefixes = aggs + tiles;
    unsigned long long* st2 = reinterpret_cast<unsigned long long*>(prefixes + tiles);
    unsigned* st1 = reinterpret_cast<unsigned*>(st2 + tiles);
    GPC_PROF_BYTES("linear_track", 12.0 * n_ev);
    linear_track_kernel<<<tiles, LIN_THREADS, 0, stream>>>(keys, tags, (unsigned)n_ev, wp, out_pos,
                                                          out_val, aggs, prefixes, st1, st2,
                                                          ticket, count);
    GPC_LAUNCH_CHECK();
    unsigned long long h = 0;
    int herr = 0;
    GPC_CUDA(cudaMemcpyAsync(&h, count, 8, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    if (herr) {
        set_error("Interval has node(s) not part of graph/pileup");
        return herr;
    }
    *n_out = h;
    prof_add_bytes("linear_track", 16.0 * (double)h);
    return GPC_OK;
}

extern "C" int gpc_control_backproject(const gpc_graph_t* graph, const uint8_t* touched,
                                       const double* bp, const double* val, uint64_t n_bp,
                                       double min_value, double value_scale, uint32_t* out_pos,
                                       double* out_val, uint64_t out_capacity, uint64_t* n_out,
                                       void* stream_) {
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
    GPC_CUDA(cudaMemcpyAsync(&h, total.p, 8, cudaMemcpyDeviceToHost, stream));
    GPC_CUDA(cudaStreamSynchronize(stream));
    *n_out = h;
    if (h > out_capacity) return GPC_ERR_CAPACITY;
    GPC_PROF("backproject_emit");
    backproject_emit_kernel<<<div_up(n, 256), 256, 0, stream>>>(
        *graph, touched, bp, val, first.as<uint32_t>(), count.as<uint32_t>(),
        offs.as<uint32_t>(), min_value, value_scale, out_pos, out_val);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
