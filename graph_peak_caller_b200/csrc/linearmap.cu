// LinearMap.from_graph on sm_100a.
//
// Replaces  LinearMap.from_graph / find_starts / find_ends   control/linearmap.py:102-154
//
// The reference sweeps the nodes in topological order and pushes
//   max_dists[next] = max(max_dists[next], max_dists[node] + node_size(node))
// along every edge: a longest-path recurrence whose dependency chain is as long as the
// graph (1.5 M levels at chr22 scale), so a level-synchronous kernel is hopeless.  What
// a variation graph does offer is BARRIERS: a node i that no edge (u, v) jumps over
// (u < i < v).  Every path from a node before i to a node after i passes through i, so
// between two consecutive barriers the recurrence is a small local problem (a bubble,
// a handful of nodes), and across barriers it is the composition of functions
//   x -> max(x + a, b)           (a: longest way from barrier to barrier, b: longest way
//                                 from a source inside the segment, if there is one)
// which is associative:  (a1, b1) then (a2, b2)  =  (a1 + a2, max(b1 + a2, b2)).
//
//   1. cover[i] = number of edges jumping over node i: a +1 / -1 difference array filled
//      with one atomic pair per edge that is not (u, u + 1), then one prefix sum
//   2. barriers = nodes with cover 0, compacted (single-pass select); node 0 and node
//      n - 1 always are barriers
//   3. one thread per segment, both directions in one launch: the local recurrence from
//      the segment's barrier (rel = longest way from the barrier, loc = longest way from
//      a local source), pulled over predecessor lists (starts) or successor lists (ends),
//      and the segment's (a, b)
//   4. scan of the (a, b) pairs over the segments (block composites, one small serial
//      pass over the block composites, apply) -> the value at every barrier
//   5. one thread per segment: value = max(x + rel, loc) for its nodes, as float64
//      (ends: linear length - value, length = value at node 0 + its size)
//
// All arithmetic is int64 (node sizes are integers; the reference's float64 sums of
// integers below 2^53 are exact, so the result is bit-identical).  A graph whose barriers
// are rare (one huge segment) degenerates to one thread walking it: correct, slow.
#include "common.cuh"
#include "scan.cuh"

namespace gpc {

constexpr long long LM_NEG = -(1ll << 60);          // "no such path"; every real value is >= 0

__device__ __forceinline__ long long lm_add(long long x, long long y) {
    return (x < 0 || y < 0) ? LM_NEG : x + y;
}

__global__ void lm_cover_kernel(gpc_graph_t g, uint32_t* __restrict__ diff) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= g.n_nodes) return;
    const uint4 a = node_record(g, u), b = node_record(g, u + 1);
    for (uint32_t e = a.y; e < b.y; ++e) {
        const uint32_t v = g.succ[e];
        if (v > u + 1) {                       // covers u + 1 .. v - 1
            atomicAdd(diff + u + 1, 1u);
            atomicAdd(diff + v, 0xffffffffu);
        }
    }
}

struct BarrierOp {
    const uint32_t* diff;       // difference array
    const uint32_t* excl;       // its exclusive prefix sum
    uint32_t* out;
    __device__ __forceinline__ bool keep(unsigned i) const { return excl[i] + diff[i] == 0u; }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const { out[w] = i; }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

// Upstream neighbours of node w in the direction of travel: predecessors when the sweep
// ascends (starts), successors when it descends (ends).
template <bool REV, typename F>
__device__ __forceinline__ bool lm_for_upstream(const gpc_graph_t& g, uint32_t w, F&& fn) {
    const uint4 a = node_record(g, w), b = node_record(g, w + 1);
    const uint32_t e0 = REV ? a.y : a.z, e1 = REV ? b.y : b.z;
    const uint32_t* adj = REV ? g.succ : g.pred;
    for (uint32_t e = e0; e < e1; ++e) fn(adj[e]);
    return e1 > e0;
}

template <bool REV>
__device__ __forceinline__ void lm_pull(const gpc_graph_t& g, uint32_t w,
                                        const long long* rel, const long long* loc,
                                        long long& r, long long& l) {
    // (rel / loc are read back by the thread that wrote them: no __restrict__, no read-only path)
    r = LM_NEG;
    l = LM_NEG;
    const bool any = lm_for_upstream<REV>(g, w, [&](uint32_t v) {
        const long long len = (long long)node_record(g, v).w;
        r = max(r, lm_add(rel[v], len));
        l = max(l, lm_add(loc[v], len));
    });
    if (!any) l = 0;                             // a source: max_dists stays 0
}

// Segment k in sweep order: its barrier, the nodes up to the next barrier, and that
// barrier (n_seg - 1 is the last segment: no next barrier).
template <bool REV>
__device__ __forceinline__ void lm_segment(const uint32_t* __restrict__ barrier, uint32_t n_seg, uint32_t n,
                                           uint32_t k, long long& anchor, long long& next) {
    if (!REV) {
        anchor = barrier[k];
        next = k + 1 < n_seg ? (long long)barrier[k + 1] : (long long)n;
    } else {
        anchor = barrier[n_seg - 1 - k];
        next = k + 1 < n_seg ? (long long)barrier[n_seg - 2 - k] : -1ll;
    }
}

template <bool REV>
__device__ __forceinline__ void lm_local(const gpc_graph_t& g, const uint32_t* __restrict__ barrier,
                                         uint32_t n_seg, uint32_t k, long long* rel, long long* loc,
                                         long long* __restrict__ fa, long long* __restrict__ fb) {
    long long anchor, next;
    lm_segment<REV>(barrier, n_seg, g.n_nodes, k, anchor, next);
    rel[anchor] = 0;
    loc[anchor] = LM_NEG;
    const long long step = REV ? -1 : 1;
    for (long long w = anchor + step; w != next; w += step) {
        long long r, l;
        lm_pull<REV>(g, (uint32_t)w, rel, loc, r, l);
        rel[w] = r;
        loc[w] = l;
    }
    if (k + 1 < n_seg) {
        long long r, l;
        lm_pull<REV>(g, (uint32_t)next, rel, loc, r, l);
        fa[k] = r;
        fb[k] = l;
    }
}

// blockIdx.y = direction; rel / loc / fa / fb of the descending sweep follow those of the
// ascending one (offset n resp. n_seg)
__global__ void lm_local_kernel(gpc_graph_t g, const uint32_t* __restrict__ barrier, uint32_t n_seg,
                                long long* rel, long long* loc,
                                long long* __restrict__ fa, long long* __restrict__ fb) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    if (blockIdx.y == 0)
        lm_local<false>(g, barrier, n_seg, k, rel, loc, fa, fb);
    else
        lm_local<true>(g, barrier, n_seg, k, rel + g.n_nodes, loc + g.n_nodes, fa + n_seg, fb + n_seg);
}

// ---- scan of x -> max(x + a, b) over the segments ---------------------------------
struct LmFn { long long a, b; };
__device__ __forceinline__ LmFn lm_then(const LmFn& f, const LmFn& g) {       // f first, then g
    return LmFn{lm_add(f.a, g.a), max(lm_add(f.b, g.a), g.b)};
}
__device__ __forceinline__ LmFn lm_shfl_up(const LmFn& f, int d) {
    return LmFn{__shfl_up_sync(FULL, f.a, d), __shfl_up_sync(FULL, f.b, d)};
}

constexpr int LM_THREADS = 256;
constexpr int LM_ITEMS = 4;
constexpr int LM_TILE = LM_THREADS * LM_ITEMS;

// APPLY == false: block_fn[dir][block] = composite of the block's functions.
// APPLY == true : x[dir][k + 1] = (functions 0 .. k)(0), with the composite of all blocks
//                 before this one in block_fn (made exclusive by lm_block_scan_kernel).
template <bool APPLY>
__global__ void __launch_bounds__(LM_THREADS)
lm_scan_kernel(const long long* __restrict__ fa, const long long* __restrict__ fb, uint32_t n_seg,
               LmFn* __restrict__ block_fn, long long* __restrict__ x) {
    __shared__ LmFn s_warp[LM_THREADS / 32];
    const uint32_t n_fn = n_seg - 1;                      // functions per direction
    const uint32_t dir = blockIdx.y;
    fa += (size_t)dir * n_seg;
    fb += (size_t)dir * n_seg;
    const uint32_t base = blockIdx.x * LM_TILE + threadIdx.x * LM_ITEMS;
    LmFn f[LM_ITEMS];
    LmFn acc{0, LM_NEG};                                  // identity
#pragma unroll
    for (int j = 0; j < LM_ITEMS; ++j) {
        f[j] = base + j < n_fn ? LmFn{fa[base + j], fb[base + j]} : LmFn{0, LM_NEG};
        acc = lm_then(acc, f[j]);
    }
    // inclusive scan of the thread composites over the block
    LmFn incl = acc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const LmFn o = lm_shfl_up(incl, d);
        if (lane_id() >= (unsigned)d) incl = lm_then(o, incl);
    }
    if (lane_id() == 31) s_warp[warp_id()] = incl;
    __syncthreads();
    LmFn before{0, LM_NEG};                               // composite of the warps before mine
    for (unsigned w = 0; w < warp_id(); ++w) before = lm_then(before, s_warp[w]);
    if (!APPLY) {
        if (threadIdx.x == LM_THREADS - 1)
            block_fn[(size_t)dir * gridDim.x + blockIdx.x] = lm_then(before, incl);
        return;
    }
    // exclusive composite of this thread = blocks before, warps before, lanes before
    LmFn excl = lm_shfl_up(incl, 1);
    if (lane_id() == 0) excl = LmFn{0, LM_NEG};
    LmFn run = lm_then(lm_then(block_fn[(size_t)dir * gridDim.x + blockIdx.x], before), excl);
    long long* xd = x + (size_t)dir * n_seg;
    if (blockIdx.x == 0 && threadIdx.x == 0) xd[0] = 0;
#pragma unroll
    for (int j = 0; j < LM_ITEMS; ++j) {
        if (base + j < n_fn) {
            run = lm_then(run, f[j]);
            xd[base + j + 1] = max(run.a, run.b);         // applied to x0 = 0
        }
    }
}

// block composites -> exclusive (serial: a few thousand entries per direction)
__global__ void lm_block_scan_kernel(LmFn* __restrict__ block_fn, uint32_t n_blocks) {
    if (threadIdx.x != 0) return;
    LmFn* f = block_fn + (size_t)blockIdx.x * n_blocks;
    LmFn run{0, LM_NEG};
    for (uint32_t i = 0; i < n_blocks; ++i) {
        const LmFn t = f[i];
        f[i] = run;
        run = lm_then(run, t);
    }
}

template <bool REV>
__device__ __forceinline__ void lm_final(const gpc_graph_t& g, const uint32_t* __restrict__ barrier,
                                         uint32_t n_seg, uint32_t k, const long long* __restrict__ rel,
                                         const long long* __restrict__ loc, const long long* __restrict__ x,
                                         double* __restrict__ out) {
    long long anchor, next;
    lm_segment<REV>(barrier, n_seg, g.n_nodes, k, anchor, next);
    const long long xk = x[k];
    // ends: linear length = longest distance of node 0 (the last barrier of the descending
    // sweep) from the sink + its size (linearmap.py:153)
    const long long length = REV ? x[n_seg - 1] + (long long)node_record(g, 0).w : 0;
    const long long step = REV ? -1 : 1;
    for (long long w = anchor; w != next; w += step) {
        const long long v = max(lm_add(xk, rel[w]), loc[w]);
        out[w] = REV ? (double)(length - v) : (double)v;
    }
}

__global__ void lm_final_kernel(gpc_graph_t g, const uint32_t* __restrict__ barrier, uint32_t n_seg,
                                const long long* __restrict__ rel, const long long* __restrict__ loc,
                                const long long* __restrict__ x, double* __restrict__ starts,
                                double* __restrict__ ends) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_seg) return;
    if (blockIdx.y == 0)
        lm_final<false>(g, barrier, n_seg, k, rel, loc, x, starts);
    else
        lm_final<true>(g, barrier, n_seg, k, rel + g.n_nodes, loc + g.n_nodes, x + n_seg, ends);
}

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_linear_map_from_graph(const gpc_graph_t* graph, double* starts, double* ends,
                                         uint64_t* n_barriers, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_barriers) *n_barriers = 0;
    if (!graph || !starts || !ends) { set_error("gpc_linear_map_from_graph: null argument"); return GPC_ERR_ARG; }
    if (graph->flags & GPC_GRAPH_UNORDERED) {
        set_error("gpc_linear_map_from_graph: node ids must ascend along every edge");
        return GPC_ERR_ARG;
    }
    const uint32_t n = graph->n_nodes;
    if (n == 0) return GPC_OK;
    Scratch diff, excl, barrier, cnt;
    GPC_CUDA(diff.alloc((size_t)(n + 1) * 4, stream));
    GPC_CUDA(excl.alloc((size_t)(n + 1) * 4, stream));
    GPC_CUDA(barrier.alloc((size_t)n * 4, stream));
    GPC_CUDA(cnt.alloc(8, stream));
    GPC_CUDA(cudaMemsetAsync(diff.p, 0, (size_t)(n + 1) * 4, stream));
    GPC_PROF("lm_cover");
    lm_cover_kernel<<<div_up(n, 256), 256, 0, stream>>>(*graph, diff.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    GPC_TRY(exclusive_scan<uint32_t>(diff.as<uint32_t>(), excl.as<uint32_t>(), (size_t)n + 1, nullptr, stream));
    GPC_TRY(select_if("lm_barriers", BarrierOp{diff.as<uint32_t>(), excl.as<uint32_t>(), barrier.as<uint32_t>()},
                      n, cnt.as<unsigned long long>(), stream));
    unsigned long long n_seg64 = 0;
    GPC_READ_BACK({cnt.p, &n_seg64, 8u});
    const uint32_t n_seg = (uint32_t)n_seg64;
    if (n_barriers) *n_barriers = n_seg64;
    if (n_seg == 0) { set_error("gpc_linear_map_from_graph: no barrier node (internal)"); return GPC_ERR_INTERNAL; }
    diff.release();
    excl.release();
    Scratch rel, loc, fa, fb, x, bf;
    GPC_CUDA(rel.alloc((size_t)n * 16, stream));
    GPC_CUDA(loc.alloc((size_t)n * 16, stream));
    GPC_CUDA(fa.alloc((size_t)n_seg * 16, stream));
    GPC_CUDA(fb.alloc((size_t)n_seg * 16, stream));
    GPC_CUDA(x.alloc((size_t)n_seg * 16, stream));
    const unsigned seg_blocks = div_up(n_seg, 128);
    GPC_PROF("lm_local");
    lm_local_kernel<<<dim3(seg_blocks, 2), 128, 0, stream>>>(*graph, barrier.as<uint32_t>(), n_seg,
                                                            rel.as<long long>(), loc.as<long long>(),
                                                            fa.as<long long>(), fb.as<long long>());
    GPC_LAUNCH_CHECK();
    const unsigned scan_blocks = n_seg > 1 ? div_up(n_seg - 1, LM_TILE) : 1;
    GPC_CUDA(bf.alloc((size_t)scan_blocks * 2 * sizeof(LmFn), stream));
    GPC_PROF("lm_scan_reduce");
    lm_scan_kernel<false><<<dim3(scan_blocks, 2), LM_THREADS, 0, stream>>>(
        fa.as<long long>(), fb.as<long long>(), n_seg, bf.as<LmFn>(), x.as<long long>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("lm_block_scan");
    lm_block_scan_kernel<<<2, 32, 0, stream>>>(bf.as<LmFn>(), scan_blocks);
    GPC_LAUNCH_CHECK();
    GPC_PROF("lm_scan_apply");
    lm_scan_kernel<true><<<dim3(scan_blocks, 2), LM_THREADS, 0, stream>>>(
        fa.as<long long>(), fb.as<long long>(), n_seg, bf.as<LmFn>(), x.as<long long>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("lm_final");
    lm_final_kernel<<<dim3(seg_blocks, 2), 128, 0, stream>>>(*graph, barrier.as<uint32_t>(), n_seg,
                                                            rel.as<long long>(), loc.as<long long>(),
                                                            x.as<long long>(), starts, ends);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
