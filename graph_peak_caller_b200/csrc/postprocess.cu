// Threshold consumers on sm_100a: hole cleaning, max paths, peak scoring.
//
// Replaces the reference's
//   HolesCleaner                       postprocess/holecleaner.py:9-109
//   SegmentAnalyzer / SegmentSplitter  postprocess/segmentanalyzer.py:5-134
//   StubsFilter, LineGraph, DividedLinegraph.filter_small,
//   PosDividedLineGraph.max_paths      postprocess/graphs.py:14-355
//   SubGraphAnalyzer                   postprocess/subgraphanalyzer.py:4-39
//   SparseMaxPaths                     postprocess/maxpaths.py:11-191
//   CallPeaksFromQvalues.__get_max_paths tail, DensePileup.get_interval_values
//                                      callpeaks.py:186-212, mindense.py:25-59
//
// Holes and peaks are cut into per-node pieces (one thread per piece); pieces
// are the vertices of a line graph whose edges follow the graph's edges.  The
// reference runs scipy all-pairs shortest paths per connected component; here
// components come from a lock-free union-find, the "> 36 vertices" and "no
// entry vertex" rules are per-component reductions, and the distance test
// (shortest entry->sink path through a piece  >  read length) is a bounded
// number of relaxation rounds over the few vertices of small components.
// Max paths replay scipy's Bellman-Ford sweep order per component (one thread
// per component) so that ties break exactly as in the reference.
#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace gpc {

constexpr int PIECE_TAIL = 0, PIECE_WHOLE = 1, PIECE_HEAD = 2, PIECE_INTERNAL = 3;
constexpr int INF_I = 1 << 30;

__device__ __forceinline__ uint32_t node_off_at(const gpc_graph_t& g, uint32_t v) {
    return __ldg(g.node_rec + 4ull * v);
}
// largest v with node_off[v] <= pos   (np.digitize(pos, node_off) - 1)
__device__ __forceinline__ uint32_t node_containing(const gpc_graph_t& g, uint32_t pos) {
    uint32_t lo = 0, hi = g.n_nodes + 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (node_off_at(g, mid) <= pos) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}
// largest v with node_off[v] < pos    (np.digitize(pos, node_off, True) - 1)
__device__ __forceinline__ uint32_t node_ending(const gpc_graph_t& g, uint32_t pos) {
    uint32_t lo = 0, hi = g.n_nodes + 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (node_off_at(g, mid) < pos) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* a, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- segments (holes or peaks) -> per-node pieces ----------------------------

// seg s = [pos[first + 2s], pos[first + 2s + 1])  (end of the last one may be
// `tail_end` when the track ends inside a run)
struct Segments {
    const uint32_t* pos;
    uint32_t first;
    uint32_t n_seg;
    uint32_t n_pos;
    uint32_t tail_end;
};
__device__ __forceinline__ void segment_bounds(const Segments& sg, uint32_t s, uint32_t& a,
                                               uint32_t& b) {
    uint32_t i = sg.first + 2 * s;
    a = sg.pos[i];
    b = (i + 1 < sg.n_pos) ? sg.pos[i + 1] : sg.tail_end;
}

__global__ void segment_nodes_kernel(gpc_graph_t g, Segments sg, uint32_t* __restrict__ v0,
                                     uint32_t* __restrict__ v1, uint32_t* __restrict__ n_pieces,
                                     uint32_t* __restrict__ is_multi,
                                     uint32_t* __restrict__ n_between) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sg.n_seg) return;
    uint32_t a, b;
    segment_bounds(sg, s, a, b);
    uint32_t x = node_containing(g, a), y = node_ending(g, b);
    v0[s] = x;
    v1[s] = y;
    n_pieces[s] = y - x + 1;
    if (is_multi) {
        is_multi[s] = y != x;
        n_between[s] = y > x + 1 ? y - x - 1 : 0;
    }
}

// Hole pieces: any order is fine (one thread per piece, contiguous per hole).
__global__ void hole_pieces_kernel(gpc_graph_t g, Segments sg, const uint32_t* __restrict__ v0,
                                   const uint32_t* __restrict__ offs, uint32_t n_pieces,
                                   uint32_t* __restrict__ p_node, uint32_t* __restrict__ p_start,
                                   uint32_t* __restrict__ p_end, uint8_t* __restrict__ p_type) {
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= n_pieces) return;
    uint32_t s = upper_bound_u32(offs, sg.n_seg, pid) - 1;
    uint32_t a, b;
    segment_bounds(sg, s, a, b);
    uint32_t v = v0[s] + (pid - offs[s]);
    uint32_t ns = node_off_at(g, v), ne = node_off_at(g, v + 1);
    uint32_t ps = max(a, ns), pe = min(b, ne);
    bool at_start = ps == ns, at_end = pe == ne;
    p_node[pid] = v;
    p_start[pid] = ps;
    p_end[pid] = pe;
    p_type[pid] = at_start ? (at_end ? PIECE_WHOLE : PIECE_HEAD)
                           : (at_end ? PIECE_TAIL : PIECE_INTERNAL);
}

// Peak pieces in SegmentSplitter order (segmentanalyzer.py:106-134): first the
// first piece of every segment, then the last piece of every multi-node
// segment, then the nodes in between, all in segment order.
__global__ void peak_pieces_kernel(gpc_graph_t g, Segments sg, const uint32_t* __restrict__ v0,
                                   const uint32_t* __restrict__ v1,
                                   const uint32_t* __restrict__ multi_rank,
                                   const uint32_t* __restrict__ between_offs, uint32_t n_multi,
                                   uint32_t* __restrict__ p_node, uint32_t* __restrict__ p_start,
                                   uint32_t* __restrict__ p_end, uint8_t* __restrict__ p_type) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sg.n_seg) return;
    uint32_t a, b;
    segment_bounds(sg, s, a, b);
    auto put = [&](uint32_t pid, uint32_t v) {
        uint32_t ns = node_off_at(g, v), ne = node_off_at(g, v + 1);
        uint32_t ps = max(a, ns), pe = min(b, ne);
        bool at_start = ps == ns, at_end = pe == ne;
        p_node[pid] = v;
        p_start[pid] = ps;
        p_end[pid] = pe;
        p_type[pid] = at_start ? (at_end ? PIECE_WHOLE : PIECE_HEAD)
                               : (at_end ? PIECE_TAIL : PIECE_INTERNAL);
    };
    put(s, v0[s]);
    if (v1[s] != v0[s]) {
        put(sg.n_seg + multi_rank[s], v1[s]);
        uint32_t base = sg.n_seg + n_multi + between_offs[s];
        for (uint32_t v = v0[s] + 1; v < v1[s]; ++v) put(base + (v - v0[s] - 1), v);
    }
}

// ---- line graph over the pieces ----------------------------------------------

struct PieceGraph {
    uint32_t n_pieces;
    const uint32_t* p_node;
    const uint8_t* p_type;
    const uint8_t* p_vertex;     // 1: piece is a vertex of the line graph
    int32_t* in_vertex;          // node -> piece id of its whole/head vertex, or -1
    int32_t* out_vertex;         // node -> piece id of its tail/whole vertex, or -1
    uint8_t* node_into;          // node carries a whole/head piece (vertex or not)
    uint8_t* node_outof;         // node carries a tail/whole piece (vertex or not)
};

// Holes only: drop whole pieces on untouched nodes (they are always kept,
// holecleaner.py:61-70) and pieces on graph sources / sinks (StubsFilter,
// graphs.py:60-75); register the rest in the node maps.
__global__ void hole_vertices_kernel(gpc_graph_t g, const uint8_t* __restrict__ touched,
                                     uint32_t n_pieces, const uint32_t* __restrict__ p_node,
                                     const uint8_t* __restrict__ p_type,
                                     uint8_t* __restrict__ p_vertex, uint8_t* __restrict__ p_keep,
                                     int32_t* __restrict__ in_vertex,
                                     int32_t* __restrict__ out_vertex,
                                     uint8_t* __restrict__ node_into,
                                     uint8_t* __restrict__ node_outof) {
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= n_pieces) return;
    int t = p_type[pid];
    uint32_t v = p_node[pid];
    p_vertex[pid] = 0;
    p_keep[pid] = 0;
    if (t == PIECE_INTERNAL) return;
    if (t == PIECE_WHOLE && touched && !touched[v]) { p_keep[pid] = 1; return; }
    uint4 a = node_record(g, v), b = node_record(g, v + 1);
    bool has_succ = b.y > a.y, has_pred = b.z > a.z;
    if (t != PIECE_HEAD) node_outof[v] = 1;
    if (t != PIECE_TAIL) node_into[v] = 1;
    bool vertex = (t == PIECE_TAIL) ? has_succ : (t == PIECE_HEAD) ? has_pred : (has_succ && has_pred);
    if (!vertex) { p_keep[pid] = 1; return; }
    p_vertex[pid] = 1;
    if (t != PIECE_HEAD) out_vertex[v] = (int32_t)pid;
    if (t != PIECE_TAIL) in_vertex[v] = (int32_t)pid;
}

// Peaks: every non-internal piece is a vertex (PosStubFilter, graphs.py:83-106).
// Vertex ids follow the reference: tails, wholes, heads, each in piece order.
__global__ void peak_vertices_kernel(uint32_t n_pieces, const uint32_t* __restrict__ p_node,
                                     const uint8_t* __restrict__ p_type,
                                     const uint32_t* __restrict__ rank_tail,
                                     const uint32_t* __restrict__ rank_whole,
                                     const uint32_t* __restrict__ rank_head, uint32_t n_tail,
                                     uint32_t n_whole, int32_t* __restrict__ vertex_of_piece,
                                     uint32_t* __restrict__ piece_of_vertex,
                                     int32_t* __restrict__ in_vertex,
                                     int32_t* __restrict__ out_vertex) {
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= n_pieces) return;
    int t = p_type[pid];
    if (t == PIECE_INTERNAL) { vertex_of_piece[pid] = -1; return; }
    uint32_t vid = t == PIECE_TAIL ? rank_tail[pid]
                 : t == PIECE_WHOLE ? n_tail + rank_whole[pid]
                                    : n_tail + n_whole + rank_head[pid];
    vertex_of_piece[pid] = (int32_t)vid;
    piece_of_vertex[vid] = pid;
    uint32_t v = p_node[pid];
    if (t != PIECE_HEAD) out_vertex[v] = (int32_t)vid;
    if (t != PIECE_TAIL) in_vertex[v] = (int32_t)vid;
}

__global__ void class_flags_kernel(uint32_t n, const uint8_t* __restrict__ p_type,
                                   uint32_t* __restrict__ f_tail, uint32_t* __restrict__ f_whole,
                                   uint32_t* __restrict__ f_head, uint32_t* __restrict__ f_internal) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = p_type[i];
    f_tail[i] = t == PIECE_TAIL;
    f_whole[i] = t == PIECE_WHOLE;
    f_head[i] = t == PIECE_HEAD;
    f_internal[i] = t == PIECE_INTERNAL;
}

// union-find -------------------------------------------------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    while (true) {
        uint32_t p = parent[x];
        if (p == x) return x;
        uint32_t gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving (benign race)
        x = p;
    }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }     // hook larger root under smaller
        uint32_t old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
    }
}

__global__ void iota_kernel(uint32_t* a, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// Vertex ids: holes use piece ids, peaks use reference vertex ids; `node_of`
// gives the graph node of a vertex, `is_source` says whether it has out edges
// to other pieces (tails and wholes).
struct VertexView {
    uint32_t n;
    const uint32_t* node_of;     // per vertex
    const uint8_t* type_of;      // per vertex
    const uint8_t* alive;        // per vertex (may be NULL: all alive)
    const int32_t* in_vertex;    // per node
    const int32_t* out_vertex;   // per node
};

__global__ void uf_edges_kernel(gpc_graph_t g, VertexView vw, uint32_t* __restrict__ parent) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= vw.n) return;
    if (vw.alive && !vw.alive[u]) return;
    if (vw.type_of[u] == PIECE_HEAD) return;
    uint32_t a = vw.node_of[u];
    uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
    for (uint32_t e = r0.y; e < r1.y; ++e) {
        int32_t w = vw.in_vertex[g.succ[e]];
        if (w >= 0) uf_union(parent, u, (uint32_t)w);
    }
}

// Two-level variant for large vertex sets: edges between vertices of the same
// 1024-vertex block (vertex ids follow graph positions, so that is almost every
// edge) are united in shared memory without touching global memory; each block
// then writes its vertices' parents (the block-local minimum) with plain stores,
// and the few edges that cross a block edge go to a list that a second, tiny
// kernel unites with the global lock-free algorithm.  Replaces one global
// find/CAS chain per edge (0.26 ms for 2.9 M hole pieces) -- and also the iota.
constexpr int UFB_THREADS = 1024;

__device__ __forceinline__ uint32_t ufs_find(volatile uint32_t* parent, uint32_t x) {
    while (true) {
        uint32_t p = parent[x];
        if (p == x) return x;
        uint32_t gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
    }
}

__global__ void __launch_bounds__(UFB_THREADS)
uf_block_edges_kernel(gpc_graph_t g, VertexView vw, uint32_t* __restrict__ parent,
                      uint2* __restrict__ cross, unsigned long long cross_cap,
                      unsigned long long* __restrict__ n_cross) {
    __shared__ uint32_t s_parent[UFB_THREADS];
    const uint32_t base = blockIdx.x * UFB_THREADS;
    const uint32_t u = base + threadIdx.x;
    s_parent[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const bool live = u < vw.n && !(vw.alive && !vw.alive[u]) && vw.type_of[u] != PIECE_HEAD;
    if (live) {
        uint32_t a = vw.node_of[u];
        uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
        for (uint32_t e = r0.y; e < r1.y; ++e) {
            int32_t w = vw.in_vertex[g.succ[e]];
            if (w < 0) continue;
            if ((uint32_t)w >= base && (uint32_t)w < base + UFB_THREADS) {
                uint32_t x = threadIdx.x, y = (uint32_t)w - base;
                while (true) {                       // lock-free union, smaller index wins
                    x = ufs_find(s_parent, x);
                    y = ufs_find(s_parent, y);
                    if (x == y) break;
                    if (x < y) { uint32_t t = x; x = y; y = t; }
                    if (atomicCAS(&s_parent[x], x, y) == x) break;
                }
            } else {
                unsigned long long at = atomicAdd(n_cross, 1ull);
                if (at < cross_cap) cross[at] = make_uint2(u, (uint32_t)w);
            }
        }
    }
    __syncthreads();
    if (u < vw.n) parent[u] = base + ufs_find(s_parent, threadIdx.x);
}

__global__ void uf_cross_edges_kernel(const uint2* __restrict__ cross,
                                      const unsigned long long* __restrict__ n_cross,
                                      uint32_t* __restrict__ parent) {
    const unsigned long long n = *n_cross;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        uf_union(parent, cross[i].x, cross[i].y);
}

// entry / exit flags of hole vertices (graphs.py:47-58) and component stats
// np.searchsorted(node_off, last position): number of node offsets < pos.  Successors
// with a larger 1-based index are ignored (graphs.py:55) -> 0-based bound  s < last_node.
__global__ void last_node_kernel(gpc_graph_t g, const uint32_t* __restrict__ last_pos, int trailing,
                                 uint32_t* __restrict__ out) {
    if (!trailing) { *out = g.n_nodes + 1; return; }
    *out = node_ending(g, *last_pos) + 1;
}

__global__ void hole_flags_kernel(gpc_graph_t g, const uint8_t* __restrict__ touched,
                                  const uint32_t* __restrict__ last_node_ptr, VertexView vw,
                                  const uint8_t* __restrict__ node_into,
                                  const uint8_t* __restrict__ node_outof,
                                  uint32_t* __restrict__ parent, uint8_t* __restrict__ is_entry,
                                  uint8_t* __restrict__ is_exit, uint32_t* __restrict__ comp_size,
                                  uint32_t* __restrict__ comp_entry) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= vw.n) return;
    is_entry[u] = 0;
    is_exit[u] = 0;
    if (!vw.alive[u]) return;
    const uint32_t last_node = *last_node_ptr;
    int t = vw.type_of[u];
    uint32_t a = vw.node_of[u];
    uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
    bool entry = t == PIECE_TAIL;
    if (!entry) {
        for (uint32_t e = r0.z; e < r1.z; ++e) {
            uint32_t p = g.pred[e];
            if (!node_outof[p] && (!touched || touched[p])) { entry = true; break; }
        }
    }
    bool ex = t == PIECE_HEAD;
    if (!ex) {
        for (uint32_t e = r0.y; e < r1.y; ++e) {
            uint32_t s = g.succ[e];
            if (!node_into[s] && s < last_node && (!touched || touched[s])) { ex = true; break; }
        }
    }
    is_entry[u] = entry;
    is_exit[u] = ex;
    uint32_t root = uf_find(parent, u);
    parent[u] = root;     // flatten (root's own entry stays root)
    atomicAdd(&comp_size[root], 1u);
    if (entry) comp_entry[root] = 1;
}

__global__ void hole_active_kernel(VertexView vw, const uint32_t* __restrict__ parent,
                                   const uint32_t* __restrict__ comp_size,
                                   const uint32_t* __restrict__ comp_entry,
                                   const uint8_t* __restrict__ is_entry,
                                   uint8_t* __restrict__ p_keep, uint8_t* __restrict__ active,
                                   int32_t* __restrict__ to, int32_t* __restrict__ fr) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= vw.n) return;
    active[u] = 0;
    if (!vw.alive[u]) return;
    uint32_t root = parent[u];
    // root of u after flattening may itself have been re-parented later: chase
    while (parent[root] != root) root = parent[root];
    if (comp_size[root] > 36 || !comp_entry[root]) { p_keep[u] = 1; return; }   // graphs.py:254-261
    active[u] = 1;
    to[u] = is_entry[u] ? 0 : INF_I;
    fr[u] = INF_I;
}

// One relaxation round over the whole active list (large active sets: whole-genome
// chromosomes have 10^4-10^5 active pieces, far too many for the single CTA below).
// changed[round] is raised when a value improved; a round whose predecessor raised
// nothing returns at once, so the 37 launches cost a few microseconds each after the
// fixed point (usually three or four rounds).
__global__ void hole_relax_kernel(gpc_graph_t g, VertexView vw, const uint32_t* __restrict__ list,
                                  uint32_t n_list, const uint32_t* __restrict__ p_start,
                                  const uint32_t* __restrict__ p_end,
                                  const uint8_t* __restrict__ is_exit, int32_t* __restrict__ to,
                                  int32_t* __restrict__ fr, int round, int* __restrict__ changed) {
    if (round > 0 && !changed[round - 1]) return;           // (uniform over the grid)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) i = n_list - 1;                        // the tail repeats the last piece
    uint32_t u = list[i];
    int t = vw.type_of[u];
    uint32_t a = vw.node_of[u];
    uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
    int size = (int)(p_end[u] - p_start[u]);
    volatile int32_t* vto = to;
    volatile int32_t* vfr = fr;
    bool any = false;
    if (t != PIECE_TAIL) {                      // pull `to` over incoming edges
        const int old = vto[u];
        int best = old;
        for (uint32_t e = r0.z; e < r1.z; ++e) {
            int32_t w = vw.out_vertex[g.pred[e]];
            if (w >= 0) {
                int d = vto[w];
                if (d < INF_I) best = min(best, d + (int)(p_end[w] - p_start[w]));
            }
        }
        if (best != old) { vto[u] = best; any = true; }
    }
    const int old = vfr[u];
    int best = old;
    if (is_exit[u]) best = min(best, size);
    if (t != PIECE_HEAD) {
        for (uint32_t e = r0.y; e < r1.y; ++e) {
            int32_t w = vw.in_vertex[g.succ[e]];
            if (w >= 0) {
                int d = vfr[w];
                if (d < INF_I) best = min(best, d + size);
            }
        }
    }
    if (best != old) { vfr[u] = best; any = true; }
    if (__syncthreads_or(any ? 1 : 0) && threadIdx.x == 0) changed[round] = 1;
}

// All rounds in one launch for the usual small active set: one CTA, a block
// barrier between rounds (in-place relaxation is monotone, so a round that
// sees some fresh values only converges faster).
constexpr int RELAX_ROUNDS = 37;
__global__ void __launch_bounds__(1024)
hole_relax_fused_kernel(gpc_graph_t g, VertexView vw, const uint32_t* __restrict__ list,
                        uint32_t n_list, const uint32_t* __restrict__ p_start,
                        const uint32_t* __restrict__ p_end, const uint8_t* __restrict__ is_exit,
                        int32_t* __restrict__ to, int32_t* __restrict__ fr) {
    for (int round = 0; round < RELAX_ROUNDS; ++round) {
        int changed = 0;
        for (uint32_t i = threadIdx.x; i < n_list; i += blockDim.x) {
            uint32_t u = list[i];
            int t = vw.type_of[u];
            uint32_t a = vw.node_of[u];
            uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
            int size = (int)(p_end[u] - p_start[u]);
            volatile int32_t* vto = to;
            volatile int32_t* vfr = fr;
            if (t != PIECE_TAIL) {
                const int old = vto[u];
                int best = old;
                for (uint32_t e = r0.z; e < r1.z; ++e) {
                    int32_t w = vw.out_vertex[g.pred[e]];
                    if (w >= 0) {
                        int d = vto[w];
                        if (d < INF_I) best = min(best, d + (int)(p_end[w] - p_start[w]));
                    }
                }
                if (best != old) { vto[u] = best; changed = 1; }
            }
            const int old = vfr[u];
            int best = old;
            if (is_exit[u]) best = min(best, size);
            if (t != PIECE_HEAD) {
                for (uint32_t e = r0.y; e < r1.y; ++e) {
                    int32_t w = vw.in_vertex[g.succ[e]];
                    if (w >= 0) {
                        int d = vfr[w];
                        if (d < INF_I) best = min(best, d + size);
                    }
                }
            }
            if (best != old) { vfr[u] = best; changed = 1; }
        }
        // fixed point: a whole round without an improvement (small holes need a
        // handful of rounds, the bound of 37 is the reference's component size limit)
        if (!__syncthreads_or(changed)) break;
    }
}

__global__ void hole_decide_kernel(const uint32_t* __restrict__ list, uint32_t n_list,
                                   const int32_t* __restrict__ to, const int32_t* __restrict__ fr,
                                   int max_size, uint8_t* __restrict__ p_keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    uint32_t u = list[i];
    long long d = (long long)to[u] + (long long)fr[u];
    p_keep[u] = (to[u] >= INF_I || fr[u] >= INF_I || d > max_size) ? 1 : 0;
}

__global__ void hole_internal_keep_kernel(uint32_t n, const uint8_t* __restrict__ p_type,
                                          const uint32_t* __restrict__ p_start,
                                          const uint32_t* __restrict__ p_end, int max_size,
                                          uint8_t* __restrict__ p_keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && p_type[i] == PIECE_INTERNAL)
        p_keep[i] = (int)(p_end[i] - p_start[i]) > max_size;
}

// kept pieces -> boundary events  (position << 1 | is_start); slot 0 is the
// synthetic "track is True from 0" event, the optional last slot the trailing
// False run of the input (holecleaner.py:36-48)
__global__ void hole_events_kernel(uint32_t n, const uint8_t* __restrict__ p_keep,
                                   const uint32_t* __restrict__ offs,
                                   const uint32_t* __restrict__ p_start,
                                   const uint32_t* __restrict__ p_end, uint32_t n_kept,
                                   int trailing, uint32_t trailing_pos,
                                   uint32_t* __restrict__ events) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        events[0] = 0u;                                       // (0, value 1)
        if (trailing) events[1 + 2 * n_kept] = (trailing_pos << 1) | 1u;
    }
    if (i < n && p_keep[i]) {
        uint32_t o = 1 + 2 * offs[i];
        events[o] = (p_start[i] << 1) | 1u;
        events[o + 1] = (p_end[i] << 1);
    }
}

__global__ void bool_runs_flag_kernel(const uint32_t* __restrict__ ev, uint32_t n,
                                      uint8_t* __restrict__ flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = ev[i];
    bool last = (i + 1 >= n) || (ev[i + 1] >> 1) != (k >> 1);
    bool keep = false;
    if (last) {
        // first element of my position group
        uint32_t gidx = i;
        while (gidx > 0 && (ev[gidx - 1] >> 1) == (k >> 1)) --gidx;
        if (gidx == 0) keep = true;
        else keep = (ev[gidx - 1] & 1u) != (k & 1u);   // value = !is_start
    }
    flag[i] = keep;
}

__global__ void bool_runs_emit_kernel(const uint32_t* __restrict__ ev, uint32_t n,
                                      const uint8_t* __restrict__ flag,
                                      const uint32_t* __restrict__ offs,
                                      uint32_t* __restrict__ out_pos, uint8_t* __restrict__ out_val) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) {
        out_pos[offs[i]] = ev[i] >> 1;
        out_val[offs[i]] = (ev[i] & 1u) ? 0 : 1;
    }
}

__global__ void compact_u32_kernel(const uint8_t* __restrict__ flag, const uint32_t* __restrict__ offs,
                                   uint32_t n, uint32_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[offs[i]] = i;
}

// ---- max paths -----------------------------------------------------------------

// score of a piece = sum of the score track over [start, end)  (maxpaths.py:175-191)
template <typename ValT>
__global__ void piece_scores_kernel(uint32_t n, const uint32_t* __restrict__ p_start,
                                    const uint32_t* __restrict__ p_end,
                                    const uint32_t* __restrict__ s_pos,
                                    const ValT* __restrict__ s_val, uint32_t n_runs,
                                    uint32_t track_size, double* __restrict__ score) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = p_start[i], b = p_end[i];
    uint32_t j = upper_bound_u32(s_pos, n_runs, a);
    double sum = 0.0;
    if (j > 0) --j; else { score[i] = 0.0; return; }
    uint32_t cur = a;
    while (cur < b && j < n_runs) {
        uint32_t nxt = (j + 1 < n_runs) ? s_pos[j + 1] : track_size;
        uint32_t stop = min(b, nxt);
        if (stop > cur) sum += (double)(stop - cur) * (double)s_val[j];
        cur = stop;
        ++j;
    }
    score[i] = sum;
}

__global__ void peak_flags_kernel(gpc_graph_t g, VertexView vw, uint8_t* __restrict__ is_entry,
                                  uint8_t* __restrict__ is_exit) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= vw.n) return;
    int t = vw.type_of[u];
    uint32_t a = vw.node_of[u];
    uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
    bool entry = true;
    if (t != PIECE_TAIL)
        for (uint32_t e = r0.z; e < r1.z; ++e)
            if (vw.out_vertex[g.pred[e]] >= 0) { entry = false; break; }
    bool ex = true;
    if (t != PIECE_HEAD)
        for (uint32_t e = r0.y; e < r1.y; ++e)
            if (vw.in_vertex[g.succ[e]] >= 0) { ex = false; break; }
    is_entry[u] = entry;
    is_exit[u] = ex;     // heads: always connected to the sink
}

__global__ void comp_keys_kernel(uint32_t n, uint32_t* __restrict__ parent,
                                 unsigned long long* __restrict__ keys) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    uint32_t root = uf_find(parent, u);
    keys[u] = ((unsigned long long)root << 32) | u;
}

__global__ void comp_heads_kernel(const unsigned long long* __restrict__ keys, uint32_t n,
                                  uint8_t* __restrict__ head) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0) || (keys[i] >> 32) != (keys[i - 1] >> 32);
}

__global__ void comp_ptr_kernel(const uint8_t* __restrict__ head, const uint32_t* __restrict__ offs,
                                uint32_t n, uint32_t n_comp, uint32_t* __restrict__ comp_ptr,
                                const unsigned long long* __restrict__ keys,
                                uint32_t* __restrict__ local_of_vertex,
                                uint32_t* __restrict__ comp_of_vertex) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) comp_ptr[n_comp] = n;
    if (i >= n) return;
    if (head[i]) comp_ptr[offs[i]] = i;
    uint32_t c = offs[i] + (head[i] ? 0 : -1);     // exclusive scan: heads before i
    uint32_t u = (uint32_t)(keys[i] & 0xffffffffu);
    comp_of_vertex[u] = c;
    local_of_vertex[u] = i;                         // position in member order
}

// Successor lists of the component graph in member order (member i = i-th
// entry of the sorted (component, vertex) keys): targets as member indices,
// ascending (= the CSR order of the reference's sliced matrix).  Built by one
// thread per vertex so that the sequential replay below touches two arrays per
// vertex instead of chasing node records, edges and vertex maps.
__device__ __forceinline__ uint32_t member_vertex(const unsigned long long* keys, uint32_t i) {
    return (uint32_t)(keys[i] & 0xffffffffu);
}

__global__ void adj_count_kernel(gpc_graph_t g, VertexView vw, uint32_t V,
                                 const unsigned long long* __restrict__ member_keys,
                                 uint32_t* __restrict__ deg /* [V + 1] */) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V) return;
    uint32_t cnt = 0;
    if (i < V) {
        uint32_t u = member_vertex(member_keys, i);
        if (vw.type_of[u] != PIECE_HEAD) {
            uint32_t a = vw.node_of[u];
            uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
            for (uint32_t e = r0.y; e < r1.y; ++e) cnt += vw.in_vertex[g.succ[e]] >= 0;
        }
    }
    deg[i] = cnt;
}

__global__ void adj_fill_kernel(gpc_graph_t g, VertexView vw, uint32_t V,
                                const unsigned long long* __restrict__ member_keys,
                                const uint32_t* __restrict__ local_of_vertex,
                                const uint32_t* __restrict__ adj_ptr,
                                const int32_t* __restrict__ weight,
                                const uint8_t* __restrict__ is_entry,
                                const uint8_t* __restrict__ is_exit, uint32_t* __restrict__ adj,
                                int32_t* __restrict__ m_weight, uint8_t* __restrict__ m_flags,
                                uint8_t* __restrict__ comp_backward) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    uint32_t u = member_vertex(member_keys, i);
    m_weight[i] = weight[u];
    m_flags[i] = (uint8_t)((is_entry[u] ? 1 : 0) | (is_exit[u] ? 2 : 0));
    if (vw.type_of[u] == PIECE_HEAD) return;
    uint32_t a = vw.node_of[u];
    uint4 r0 = node_record(g, a), r1 = node_record(g, a + 1);
    uint32_t w = adj_ptr[i];
    long long last = -1;
    while (true) {                           // selection by increasing vertex id
        long long best = -1;
        for (uint32_t e = r0.y; e < r1.y; ++e) {
            int32_t t = vw.in_vertex[g.succ[e]];
            if (t >= 0 && (long long)t > last && (best < 0 || (long long)t < best)) best = t;
        }
        if (best < 0) break;
        uint32_t m = local_of_vertex[best];
        if (m <= i) comp_backward[(uint32_t)(member_keys[i] >> 32)] = 1;
        adj[w++] = m;
        last = best;
    }
}

struct PathOut {
    uint32_t* path;        // [n_vertices] path of comp c at comp_ptr[c] (forward order, vertex ids)
    uint32_t* path_len;    // [n_comp]
    uint8_t* info;         // [n_comp] bit0 is_diff (two bindings), bit1 ambiguous
};

// PosDividedLineGraph.max_paths for ONE component, run by one thread on whatever memory
// the view points at: scipy's Bellman-Ford (csgraph.shortest_path(method="BF"), a sweep
// over the vertices in index order until nothing changes, first strict improvement wins)
// restated so that ONE pass in topological order gives the same distances AND the same
// predecessors.
//
// The reference numbers the vertices [pieces at node starts, whole nodes, pieces at node
// ends], so its sweep order is not topological: a path runs end piece -> whole nodes ->
// start piece and needs a new sweep after every step back in index.  The first version
// here replayed the sweeps literally: 6.8 sweeps per component and relaxation on the
// dense-variation workload (ncu-free measurement: 5 700 cycles per vertex).  What the
// sweeps compute is order-free except for the predecessor of a vertex with several
// tight in-edges, and that one has a closed form.  Say vertex j reaches its final
// distance in sweep s_j by a relaxation from p_j.  j passes a final value on in the sweep
// it is next visited in: s_j if p_j < j (later in the same sweep), else s_j + 1.  Among the
// in-neighbours of t whose edge is tight, the one that sets dist[t] for the last time is
// the first to arrive, i.e. the smallest (that sweep, j) -- all later ones find equality,
// not improvement.  So, with the in-neighbours of a vertex finished before the vertex (node
// ids ascend along every edge: the order `topo` -- members sorted by node -- is
// topological), one push pass keeping the lexicographic minimum of (distance, sweep, j)
// per vertex reproduces scipy's result exactly; distances to the sink are values only and
// take one reverse pass.  Writes the path as local vertex indexes, BACKWARDS, into v.path;
// returns its length, info = bit 0 is_diff (two bindings), bit 1 ambiguous.
struct MpView {
    const uint32_t* adj_ptr;   // [nv + 1] indexes into adj
    const uint32_t* adj;       // targets; local index = adj[e] - v_base
    uint32_t v_base;
    const int32_t* weight;     // [nv]
    const uint8_t* flags;      // [nv] bit 0 entry, bit 1 exit
    const uint32_t* topo;      // [nv] local index = topo[i] - v_base, in topological order
    long long* dist;           // [nv] working arrays
    long long* dsink;
    int32_t* pred;
    int32_t* sweep;
    uint32_t* path;            // [nv]
};

__device__ __forceinline__ int mp_solve(const MpView& v, const int nv, int& info_out) {
    const long long INF = (1ll << 60);
    long long* const dist = v.dist;
    long long* const dsink = v.dsink;
    int32_t* const pred = v.pred;
    int32_t* const sweep = v.sweep;
    // Edge iteration of local vertex j: piece targets, sink (local index nv) last.
    auto for_edges = [&](int j, auto&& fn) {
        const uint32_t e0 = v.adj_ptr[j], e1 = v.adj_ptr[j + 1];
        const bool to_sink = (v.flags[j] & 2) != 0;
        for (uint32_t e = e0; e < e1; ++e) fn((int)(v.adj[e] - v.v_base));
        if (to_sink) fn(nv);
    };
    // distance of every vertex to the sink (value only): reverse topological order
    for (int i = nv - 1; i >= 0; --i) {
        const int j = (int)(v.topo[i] - v.v_base);
        const long long w = -(long long)v.weight[j];
        long long best = INF;
        for_edges(j, [&](int tgt) {
            const long long d = tgt == nv ? 0 : dsink[tgt];
            if (d < INF && w + d < best) best = w + d;
        });
        dsink[j] = best;
    }
    // start = first entry vertex with minimal distance to the sink (np.argmin)
    int first = -1;
    for (int j = 0; j < nv; ++j)
        if ((v.flags[j] & 1) && (first < 0 || dsink[j] < dsink[first])) first = j;
    if (first < 0) first = 0;
    // shortest paths from `first` with scipy's predecessors (see above)
    long long dist_sink = INF;
    int pred_sink = -9999, sweep_sink = 0;
    for (int j = 0; j < nv; ++j) { dist[j] = INF; pred[j] = -9999; sweep[j] = 0; }
    dist[first] = 0;
    for (int i = 0; i < nv; ++i) {
        const int j = (int)(v.topo[i] - v.v_base);
        const long long d1 = dist[j];
        if (d1 >= INF) continue;
        const long long c = d1 - (long long)v.weight[j];
        const int sj = sweep[j] + (pred[j] > j ? 1 : 0);      // the sweep j passes its final value on in
        for_edges(j, [&](int tgt) {
            if (tgt == nv) {
                if (c < dist_sink || (c == dist_sink && (sj < sweep_sink || (sj == sweep_sink && j < pred_sink)))) {
                    dist_sink = c; sweep_sink = sj; pred_sink = j;
                }
            } else {
                const long long d2 = dist[tgt];
                if (c < d2 || (c == d2 && (sj < sweep[tgt] || (sj == sweep[tgt] && j < pred[tgt])))) {
                    dist[tgt] = c; sweep[tgt] = sj; pred[tgt] = j;
                }
            }
        });
    }
    // _backtrace (graphs.py:273-283): stops at local vertex 0 as well
    int len = 0;
    int cur = pred_sink;
    while (cur > 0) {
        v.path[len++] = (uint32_t)cur;       // local ids, backwards
        cur = pred[cur];
    }
    if (len == 0 || (int)v.path[len - 1] != first) v.path[len++] = (uint32_t)first;
    // SubGraphAnalyzer (subgraphanalyzer.py:11-39) on the forward path
    long long score = 0;                          // dists[sink, sink] == 0 takes part in the min
    long long wsum = 0;
    for (int j = 0; j < nv; ++j) {
        if (dsink[j] < score) score = dsink[j];
        wsum += (long long)v.weight[j];
    }
    int info = 0;
    if (len > 1) {
        if (-wsum != score) info |= 1;
        // "is the target on the path" as a flag per vertex (the predecessor array has done
        // its job): a scan of the path per edge made peaks of a few hundred pieces -- dense
        // variation -- quadratic
        for (int j = 0; j < nv; ++j) pred[j] = 0;
        for (int z = 0; z < len; ++z) pred[(int)v.path[z]] = 1;
        bool amb = false;
        for (int q = len - 1; q >= 1 && !amb; --q) {      // path[:-1] in forward order
            int node = (int)v.path[q];
            long long reach = dist[node];
            long long w = -(long long)v.weight[node];
            for_edges(node, [&](int tgt) {
                if (amb) return;
                if (tgt != nv && pred[tgt]) return;
                long long tail = tgt == nv ? 0 : dsink[tgt];
                if (tail < INF && reach + w + tail == score) amb = true;
            });
        }
        if (amb) info |= 2;
    }
    info_out = info;
    return len;
}

// One WARP per component.  The pass itself is one dependent access after the other and
// runs on lane 0; what the other lanes are for is the memory around it: they stage the
// component's adjacency, weights, flags and topological order in shared memory (coalesced),
// the working arrays live there too (~30 cycles per dependent access instead of an L2
// round trip), and they write the path out in forward order.  Components come in size
// classes -- a list per class, filled by mp_classify_kernel, one launch per class -- so
// that the shared memory of a block fits the components it holds and every warp of a
// block has work:
//   class 1: <= 64 vertices, 2: <= 256, 3: <= 1024 (edges <= 3 per vertex each),
//   class 0: everything larger, working arrays in global memory.
// (One THREAD per component, eight to a block, was the first version: the threads of a
// warp diverge on everything, so a warp took the time of its eight components one after
// the other.)
constexpr int MPW_WARPS = 4;
constexpr int MP_BYTES_PER_VERTEX = 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 12 + 1;   // dist, dsink, pred, sweep, path,
                                                                            // weight, adj_ptr, topo, adj (3x), flags
__host__ __device__ constexpr int mp_class_cap(int cls) { return cls == 1 ? 64 : (cls == 2 ? 256 : 1024); }
__host__ __device__ constexpr size_t mp_warp_smem(int capv) { return (size_t)capv * MP_BYTES_PER_VERTEX + 32; }
__device__ __forceinline__ int mp_class(uint32_t nv, uint32_t ne) {
    for (int cls = 1; cls <= 3; ++cls)
        if (nv <= (uint32_t)mp_class_cap(cls) && ne <= 3u * mp_class_cap(cls)) return cls;
    return 0;
}

// class_count[cls] components in class_list[cls * n_comp ...] (order inside a class is
// irrelevant: every component writes its own outputs)
__global__ void mp_classify_kernel(uint32_t n_comp, const uint32_t* __restrict__ comp_ptr,
                                   const uint32_t* __restrict__ adj_ptr, uint32_t* __restrict__ class_count,
                                   uint32_t* __restrict__ class_list) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_comp) return;
    const uint32_t lo = comp_ptr[c], hi = comp_ptr[c + 1];
    const int cls = mp_class(hi - lo, adj_ptr[hi] - adj_ptr[lo]);
    class_list[(size_t)cls * n_comp + atomicAdd(class_count + cls, 1u)] = c;
}

// topological keys of the members: (component << node_bits) | node of the vertex
__global__ void mp_topo_keys_kernel(uint32_t V, const unsigned long long* __restrict__ member_keys,
                                    const uint32_t* __restrict__ comp_of_vertex,
                                    const uint32_t* __restrict__ node_of, int node_bits,
                                    unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const uint32_t u = member_vertex(member_keys, i);
    keys[i] = ((unsigned long long)comp_of_vertex[u] << node_bits) | node_of[u];
    vals[i] = i;
}

template <int CLS>
__global__ void __launch_bounds__(MPW_WARPS * 32)
max_paths_kernel(uint32_t n_comp, const uint32_t* __restrict__ class_count,
                 const uint32_t* __restrict__ class_list, const uint32_t* __restrict__ comp_ptr,
                 const unsigned long long* __restrict__ member_keys,
                 const uint32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                 const int32_t* __restrict__ m_weight, const uint8_t* __restrict__ m_flags,
                 const uint32_t* __restrict__ topo_member,
                 long long* g_dist, long long* g_dsink, int32_t* g_pred, int32_t* g_sweep, PathOut out) {
    extern __shared__ __align__(16) unsigned char s_mp[];
    const unsigned lane = lane_id();
    const uint32_t slot = blockIdx.x * MPW_WARPS + warp_id();
    if (slot >= class_count[CLS]) return;
    const uint32_t c = class_list[(size_t)CLS * n_comp + slot];
    const uint32_t lo = comp_ptr[c], hi = comp_ptr[c + 1];
    const int nv = (int)(hi - lo);           // vertices without the sink
    if (nv == 1) {                           // graphs.py:342-345
        if (lane == 0) {
            out.path[lo] = member_vertex(member_keys, lo);
            out.path_len[c] = 1;
            out.info[c] = 0;
        }
        return;
    }
    MpView v;
    if (CLS != 0) {
        constexpr int CAP = mp_class_cap(CLS == 0 ? 1 : CLS);
        const uint32_t e_lo = adj_ptr[lo], ne = adj_ptr[hi] - e_lo;
        unsigned char* base = s_mp + (size_t)warp_id() * mp_warp_smem(CAP);
        long long* s_dist = reinterpret_cast<long long*>(base);
        long long* s_dsink = s_dist + CAP;
        int32_t* s_pred = reinterpret_cast<int32_t*>(s_dsink + CAP);
        int32_t* s_sweep = s_pred + CAP;
        uint32_t* s_path = reinterpret_cast<uint32_t*>(s_sweep + CAP);
        int32_t* s_w = reinterpret_cast<int32_t*>(s_path + CAP);
        uint32_t* s_topo = reinterpret_cast<uint32_t*>(s_w + CAP);
        uint32_t* s_adjptr = s_topo + CAP;                                 // CAP + 1 used
        uint32_t* s_adj = s_adjptr + CAP + 4;
        uint8_t* s_flags = reinterpret_cast<uint8_t*>(s_adj + 3 * CAP);
        for (int j = lane; j < nv; j += 32) {
            s_w[j] = m_weight[lo + j];
            s_flags[j] = m_flags[lo + j];
            s_topo[j] = topo_member[lo + j] - lo;
        }
        for (int j = lane; j <= nv; j += 32) s_adjptr[j] = adj_ptr[lo + j] - e_lo;
        for (uint32_t e = lane; e < ne; e += 32) s_adj[e] = adj[e_lo + e] - lo;
        v = MpView{s_adjptr, s_adj, 0u, s_w, s_flags, s_topo, s_dist, s_dsink, s_pred, s_sweep, s_path};
    } else {
        v = MpView{adj_ptr + lo, adj, lo, m_weight + lo, m_flags + lo, topo_member + lo, g_dist + lo,
                   g_dsink + lo, g_pred + lo, g_sweep + lo, out.path + lo};
    }
    __syncwarp();
    int len = 0, info = 0;
    if (lane == 0) len = mp_solve(v, nv, info);
    __syncwarp();
    len = __shfl_sync(FULL, len, 0);
    // forward order, vertex ids
    if (CLS != 0) {
        for (int a = lane; a < len; a += 32)
            out.path[lo + a] = member_vertex(member_keys, lo + v.path[len - 1 - a]);
    } else {
        for (int a = lane; a < len / 2; a += 32) {
            const uint32_t t = v.path[a];
            v.path[a] = v.path[len - 1 - a];
            v.path[len - 1 - a] = t;
        }
        __syncwarp();
        for (int a = lane; a < len; a += 32) v.path[a] = member_vertex(member_keys, lo + v.path[a]);
    }
    if (lane == 0) {
        out.path_len[c] = (uint32_t)len;
        out.info[c] = (uint8_t)info;
    }
}

// peaks: component paths first (component order), then internal pieces
// (piece order).  One thread per peak fills the fixed-size fields; node lists
// go to path_nodes at path_ptr.
__global__ void peak_lengths_kernel(uint32_t n_comp, uint32_t n_internal,
                                    const uint32_t* __restrict__ path_len,
                                    uint32_t* __restrict__ n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_comp) n_nodes[i] = path_len[i];
    else if (i < n_comp + n_internal) n_nodes[i] = 1;
}

// One warp per peak: lanes take the path nodes round-robin, each scans the q
// runs under its piece; warp reductions give length and score.
__global__ void peaks_kernel(gpc_graph_t g, uint32_t n_comp, uint32_t n_internal,
                             const uint32_t* __restrict__ comp_ptr, const uint32_t* __restrict__ path,
                             const uint32_t* __restrict__ path_len, const uint8_t* __restrict__ info,
                             const uint32_t* __restrict__ piece_of_vertex,
                             const uint32_t* __restrict__ internal_piece,
                             const uint32_t* __restrict__ p_node, const uint32_t* __restrict__ p_start,
                             const uint32_t* __restrict__ p_end, const uint32_t* __restrict__ q_pos,
                             const double* __restrict__ q_val, uint32_t n_q, int internal_id_shift,
                             const uint32_t* __restrict__ node_ptr, int32_t* __restrict__ out_nodes,
                             int32_t* __restrict__ out_start, int32_t* __restrict__ out_end,
                             int32_t* __restrict__ out_length, double* __restrict__ out_score,
                             uint8_t* __restrict__ out_info) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = lane_id();
    if (i >= n_comp + n_internal) return;
    const uint32_t o = node_ptr[i];
    const uint32_t len = i < n_comp ? path_len[i] : 1;
    double best = -INFINITY;
    long long bp = 0;
    uint32_t first_start = 0, last_end = 0;
    for (uint32_t k = lane; k < len; k += 32) {
        uint32_t pid = i < n_comp ? piece_of_vertex[path[comp_ptr[i] + k]]
                                  : internal_piece[i - n_comp];
        uint32_t v = p_node[pid];
        uint32_t ns = node_off_at(g, v), ne = node_off_at(g, v + 1);
        // the reference takes the whole node for inner path nodes and the piece
        // borders only at the two ends (maxpaths.py:157-163, mindense.py:41-52)
        uint32_t a = (k == 0) ? p_start[pid] : ns;
        uint32_t b = (k == len - 1) ? p_end[pid] : ne;
        if (k == 0) first_start = a - ns;
        if (k == len - 1) last_end = b - ns;
        int id = (int)v + g.min_node;
        if (i >= n_comp) id = (int)v + 1 + internal_id_shift;     // maxpaths.py:34
        out_nodes[o + k] = id;
        bp += (long long)b - (long long)a;
        if (b > a) {
            uint32_t j = upper_bound_u32(q_pos, n_q, a);
            j = j > 0 ? j - 1 : 0;
            while (j < n_q && q_pos[j] < b) { best = fmax(best, q_val[j]); ++j; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        best = fmax(best, __shfl_xor_sync(FULL, best, d));
        bp += __shfl_xor_sync(FULL, bp, d);
        first_start += __shfl_xor_sync(FULL, first_start, d);   // only one lane is non-zero
        last_end += __shfl_xor_sync(FULL, last_end, d);
    }
    if (lane == 0) {
        out_start[i] = (int32_t)first_start;
        out_end[i] = (int32_t)last_end;
        out_length[i] = (int32_t)bp;
        out_score[i] = bp == 0 ? 0.0 : best;
        out_info[i] = i < n_comp ? info[i] : 0;
    }
}

__global__ void score_keys_kernel(const double* __restrict__ score, uint32_t n,
                                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ~f64_to_key(score[i]); idx[i] = i; }
}

__global__ void gather_internal_kernel(uint32_t n, const uint8_t* __restrict__ p_type,
                                       const uint32_t* __restrict__ rank,
                                       uint32_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && p_type[i] == PIECE_INTERNAL) out[rank[i]] = i;
}
__global__ void vertex_fields_kernel(uint32_t n_vertices, const uint32_t* __restrict__ piece_of_vertex,
                                     const uint32_t* __restrict__ p_node,
                                     const uint8_t* __restrict__ p_type,
                                     const double* __restrict__ score, uint32_t* __restrict__ v_node,
                                     uint8_t* __restrict__ v_type, int32_t* __restrict__ weight) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_vertices) return;
    uint32_t pid = piece_of_vertex[u];
    v_node[u] = p_node[pid];
    v_type[u] = p_type[pid];
    weight[u] = (int32_t)score[pid];     // astype("int") then int32 (graphs.py:150-151, 183)
}

static int scan_flags(const uint8_t* flag, uint32_t* offs, size_t n, unsigned long long* total_host,
                      cudaStream_t stream) {
    Scratch t;
    GPC_CUDA(t.alloc(8, stream));
    GPC_TRY(exclusive_scan<uint8_t>(flag, offs, n, t.as<unsigned long long>(), stream));
    GPC_READ_BACK({t.p, total_host, (unsigned)(8)});
    return GPC_OK;
}
static int scan_counts(const uint32_t* cnt, uint32_t* offs, size_t n, unsigned long long* total_host,
                       cudaStream_t stream) {
    Scratch t;
    GPC_CUDA(t.alloc(8, stream));
    GPC_TRY(exclusive_scan<uint32_t>(cnt, offs, n, t.as<unsigned long long>(), stream));
    GPC_READ_BACK({t.p, total_host, (unsigned)(8)});
    return GPC_OK;
}

// parent[] <- union-find forest of the vertex graph (callers resolve roots with
// uf_find).  Large vertex sets take the two-level kernels.
static int uf_components(const gpc_graph_t* graph, const VertexView& vw, uint32_t* parent,
                         cudaStream_t stream) {
    const uint32_t V = vw.n;
    if (V < 4 * UFB_THREADS) {
        GPC_PROF("iota");
        iota_kernel<<<div_up(V, 256), 256, 0, stream>>>(parent, V);
        GPC_LAUNCH_CHECK();
        GPC_PROF("uf_edges");
        uf_edges_kernel<<<div_up(V, 256), 256, 0, stream>>>(*graph, vw, parent);
        GPC_LAUNCH_CHECK();
        return GPC_OK;
    }
    Scratch cross, counter;
    const unsigned long long cap = (unsigned long long)graph->n_edges + 16;
    GPC_CUDA(cross.alloc(cap * sizeof(uint2), stream));
    GPC_CUDA(counter.alloc(8, stream));
    GPC_CUDA(cudaMemsetAsync(counter.p, 0, 8, stream));
    GPC_PROF("uf_edges");
    uf_block_edges_kernel<<<div_up(V, UFB_THREADS), UFB_THREADS, 0, stream>>>(
        *graph, vw, parent, cross.as<uint2>(), cap, counter.as<unsigned long long>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("uf_cross_edges");
    uf_cross_edges_kernel<<<sm_count(), 256, 0, stream>>>(cross.as<uint2>(),
                                                         counter.as<unsigned long long>(), parent);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

}  // namespace gpc

using namespace gpc;

#define GRID(n) div_up((n), 256), 256, 0, stream

extern "C" int gpc_clean_holes(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                               uint64_t n_runs, int32_t max_size, const uint8_t* touched,
                               uint32_t* out_pos, uint8_t* out_val, uint64_t out_capacity,
                               uint64_t* n_out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_out = 0;
    if (n_runs == 0) { set_error("empty track"); return GPC_ERR_ARG; }
    const uint32_t n = graph->n_nodes;
    uint8_t v_first = 0, v_last = 0;
    uint32_t p_last = 0;
    GPC_READ_BACK({val, &v_first, (unsigned)(1)}, {val + n_runs - 1, &v_last, (unsigned)(1)}, {pos + n_runs - 1, &p_last, (unsigned)(4)});
    const uint32_t lo = v_first ? 1 : 0;
    uint32_t hi = (uint32_t)n_runs;
    const int trailing = v_last == 0;
    if (trailing) hi -= 1;
    const uint32_t n_holes = hi > lo ? (hi - lo) / 2 : 0;
    Segments sg{pos, lo, n_holes, hi, 0};
    Scratch v0, v1, cnt, offs;
    unsigned long long n_pieces = 0;
    if (n_holes) {
        GPC_CUDA(v0.alloc((size_t)n_holes * 4, stream));
        GPC_CUDA(v1.alloc((size_t)n_holes * 4, stream));
        GPC_CUDA(cnt.alloc((size_t)n_holes * 4, stream));
        GPC_CUDA(offs.alloc((size_t)n_holes * 4, stream));
        GPC_PROF("segment_nodes");
        segment_nodes_kernel<<<GRID(n_holes)>>>(*graph, sg, v0.as<uint32_t>(), v1.as<uint32_t>(),
                                                cnt.as<uint32_t>(), nullptr, nullptr);
        GPC_LAUNCH_CHECK();
        GPC_TRY(scan_counts(cnt.as<uint32_t>(), offs.as<uint32_t>(), n_holes, &n_pieces, stream));
    }
    Scratch last_node;
    GPC_CUDA(last_node.alloc(4, stream));
    GPC_PROF("last_node");
    last_node_kernel<<<1, 1, 0, stream>>>(*graph, pos + n_runs - 1, trailing, last_node.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    const uint32_t P = (uint32_t)n_pieces;
    Scratch p_node, p_start, p_end, p_type, p_vertex, p_keep, in_v, out_v, n_into, n_outof;
    GPC_CUDA(p_node.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_start.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_end.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_type.alloc(P, stream));
    GPC_CUDA(p_vertex.alloc(P, stream));
    GPC_CUDA(p_keep.alloc(P, stream));
    unsigned long long n_kept = 0;
    Scratch keep_offs;
    GPC_CUDA(keep_offs.alloc((size_t)P * 4, stream));
    if (P) {
        GPC_CUDA(in_v.alloc((size_t)n * 4, stream));
        GPC_CUDA(out_v.alloc((size_t)n * 4, stream));
        GPC_CUDA(n_into.alloc(n, stream));
        GPC_CUDA(n_outof.alloc(n, stream));
        GPC_CUDA(cudaMemsetAsync(in_v.p, 0xff, (size_t)n * 4, stream));
        GPC_CUDA(cudaMemsetAsync(out_v.p, 0xff, (size_t)n * 4, stream));
        GPC_CUDA(cudaMemsetAsync(n_into.p, 0, n, stream));
        GPC_CUDA(cudaMemsetAsync(n_outof.p, 0, n, stream));
        GPC_PROF("hole_pieces");
        hole_pieces_kernel<<<GRID(P)>>>(*graph, sg, v0.as<uint32_t>(), offs.as<uint32_t>(), P,
                                        p_node.as<uint32_t>(), p_start.as<uint32_t>(),
                                        p_end.as<uint32_t>(), p_type.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        GPC_PROF("hole_vertices");
        hole_vertices_kernel<<<GRID(P)>>>(*graph, touched, P, p_node.as<uint32_t>(),
                                          p_type.as<uint8_t>(), p_vertex.as<uint8_t>(),
                                          p_keep.as<uint8_t>(), in_v.as<int32_t>(),
                                          out_v.as<int32_t>(), n_into.as<uint8_t>(),
                                          n_outof.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        GPC_PROF("hole_internal_keep");
        hole_internal_keep_kernel<<<GRID(P)>>>(P, p_type.as<uint8_t>(), p_start.as<uint32_t>(),
                                               p_end.as<uint32_t>(), max_size, p_keep.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        VertexView vw{P, p_node.as<uint32_t>(), p_type.as<uint8_t>(), p_vertex.as<uint8_t>(),
                      in_v.as<int32_t>(), out_v.as<int32_t>()};
        Scratch parent, is_entry, is_exit, csize, centry, active, to, fr, aoffs, alist;
        GPC_CUDA(parent.alloc((size_t)P * 4, stream));
        GPC_CUDA(is_entry.alloc(P, stream));
        GPC_CUDA(is_exit.alloc(P, stream));
        GPC_CUDA(csize.alloc((size_t)P * 4, stream));
        GPC_CUDA(centry.alloc((size_t)P * 4, stream));
        GPC_CUDA(active.alloc(P, stream));
        GPC_CUDA(to.alloc((size_t)P * 4, stream));
        GPC_CUDA(fr.alloc((size_t)P * 4, stream));
        GPC_CUDA(aoffs.alloc((size_t)P * 4, stream));
        GPC_CUDA(cudaMemsetAsync(csize.p, 0, (size_t)P * 4, stream));
        GPC_CUDA(cudaMemsetAsync(centry.p, 0, (size_t)P * 4, stream));
        GPC_TRY(uf_components(graph, vw, parent.as<uint32_t>(), stream));
        GPC_PROF("hole_flags");
        hole_flags_kernel<<<GRID(P)>>>(*graph, touched, last_node.as<uint32_t>(), vw, n_into.as<uint8_t>(),
                                       n_outof.as<uint8_t>(), parent.as<uint32_t>(),
                                       is_entry.as<uint8_t>(), is_exit.as<uint8_t>(),
                                       csize.as<uint32_t>(), centry.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        GPC_PROF("hole_active");
        hole_active_kernel<<<GRID(P)>>>(vw, parent.as<uint32_t>(), csize.as<uint32_t>(),
                                        centry.as<uint32_t>(), is_entry.as<uint8_t>(),
                                        p_keep.as<uint8_t>(), active.as<uint8_t>(),
                                        to.as<int32_t>(), fr.as<int32_t>());
        GPC_LAUNCH_CHECK();
        unsigned long long n_active = 0;
        GPC_TRY(scan_flags(active.as<uint8_t>(), aoffs.as<uint32_t>(), P, &n_active, stream));
        if (n_active) {
            GPC_CUDA(alist.alloc((size_t)n_active * 4, stream));
            GPC_PROF("compact_u32");
            compact_u32_kernel<<<GRID(P)>>>(active.as<uint8_t>(), aoffs.as<uint32_t>(), P,
                                            alist.as<uint32_t>());
            GPC_LAUNCH_CHECK();
            if (n_active <= 4096) {
                GPC_PROF("hole_relax_fused");
                hole_relax_fused_kernel<<<1, 1024, 0, stream>>>(
                    *graph, vw, alist.as<uint32_t>(), (uint32_t)n_active, p_start.as<uint32_t>(),
                    p_end.as<uint32_t>(), is_exit.as<uint8_t>(), to.as<int32_t>(), fr.as<int32_t>());
                GPC_LAUNCH_CHECK();
            } else {
                // rounds in batches of six, then one look at the change flags: the fixed point
                // usually comes after three or four rounds, and 31 launches that return at once
                // still cost 4-5 us each (24 chromosomes: 4 ms of a whole-genome pass)
                Scratch chg;
                GPC_CUDA(chg.alloc(RELAX_ROUNDS * sizeof(int), stream));
                GPC_CUDA(cudaMemsetAsync(chg.p, 0, RELAX_ROUNDS * sizeof(int), stream));
                for (int round = 0; round < RELAX_ROUNDS;) {
                    const int stop = min(RELAX_ROUNDS, round + 6);
                    for (; round < stop; ++round) {
                        GPC_PROF("hole_relax");
                        hole_relax_kernel<<<GRID(n_active)>>>(*graph, vw, alist.as<uint32_t>(),
                                                              (uint32_t)n_active, p_start.as<uint32_t>(),
                                                              p_end.as<uint32_t>(), is_exit.as<uint8_t>(),
                                                              to.as<int32_t>(), fr.as<int32_t>(), round,
                                                              chg.as<int>());
                        GPC_LAUNCH_CHECK();
                    }
                    int still = 0;
                    GPC_READ_BACK({chg.as<int>() + (round - 1), &still, (unsigned)(4)});
                    if (!still) break;
                }
            }
            GPC_PROF("hole_decide");
            hole_decide_kernel<<<GRID(n_active)>>>(alist.as<uint32_t>(), (uint32_t)n_active,
                                                   to.as<int32_t>(), fr.as<int32_t>(), max_size,
                                                   p_keep.as<uint8_t>());
            GPC_LAUNCH_CHECK();
        }
        GPC_TRY(scan_flags(p_keep.as<uint8_t>(), keep_offs.as<uint32_t>(), P, &n_kept, stream));
    }
    // kept pieces -> sorted boundary events -> canonical boolean track
    const size_t n_ev = 1 + 2 * (size_t)n_kept + (trailing ? 1 : 0);
    Scratch ev0, flag, eoffs;
    GPC_CUDA(ev0.alloc(n_ev * 4, stream));
    GPC_CUDA(flag.alloc(n_ev, stream));
    GPC_CUDA(eoffs.alloc(n_ev * 4, stream));
    GPC_PROF("hole_events");
    hole_events_kernel<<<GRID(P ? P : 1)>>>(P, p_keep.as<uint8_t>(), keep_offs.as<uint32_t>(),
                                            p_start.as<uint32_t>(), p_end.as<uint32_t>(),
                                            (uint32_t)n_kept, trailing, p_last, ev0.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    // No sort: pieces are numbered hole by hole and node by node, i.e. by position,
    // the compaction keeps that order, a piece's end precedes the next piece's start
    // (equal positions: end first, as the canonicalisation wants), the synthetic
    // event sits at position 0 and the trailing one after every kept piece.
    const uint32_t* ev = ev0.as<uint32_t>();
    GPC_PROF("bool_runs_flag");
    bool_runs_flag_kernel<<<GRID(n_ev)>>>(ev, (uint32_t)n_ev, flag.as<uint8_t>());
    GPC_LAUNCH_CHECK();
    unsigned long long total = 0;
    GPC_TRY(scan_flags(flag.as<uint8_t>(), eoffs.as<uint32_t>(), n_ev, &total, stream));
    // cleaning only fills holes, so the result never has more runs than the input (+2)
    if (total > out_capacity) { *n_out = total; return GPC_ERR_CAPACITY; }
    GPC_PROF("bool_runs_emit");
    bool_runs_emit_kernel<<<GRID(n_ev)>>>(ev, (uint32_t)n_ev, flag.as<uint8_t>(),
                                          eoffs.as<uint32_t>(), out_pos, out_val);
    GPC_LAUNCH_CHECK();
    *n_out = total;
    return GPC_OK;
}

// Sub-graphs (gpc_peak_subgraphs): the line-graph components as they stand after the
// adjacency lists are made -- members in (component, vertex) order, successor lists as
// columns local to the component.
struct SubgraphOut {
    uint32_t* comp_ptr;
    uint64_t comp_capacity;
    int32_t* member_node;
    int32_t* member_weight;
    uint8_t* member_exit;
    uint32_t* row_ptr;
    uint64_t member_capacity;
    uint32_t* col;
    uint64_t col_capacity;
    uint64_t* n_comp;
    uint64_t* n_members;
    uint64_t* n_cols;
};

__global__ void subgraph_emit_kernel(uint32_t V, int32_t min_node,
                                     const unsigned long long* __restrict__ member_keys,
                                     const uint32_t* __restrict__ v_node,
                                     const uint32_t* __restrict__ comp_of_vertex,
                                     const uint32_t* __restrict__ comp_ptr,
                                     const uint32_t* __restrict__ adj_ptr,
                                     const uint32_t* __restrict__ adj,
                                     const int32_t* __restrict__ m_weight,
                                     const uint8_t* __restrict__ m_flags,
                                     int32_t* __restrict__ out_node, int32_t* __restrict__ out_weight,
                                     uint8_t* __restrict__ out_exit, uint32_t* __restrict__ out_row_ptr,
                                     uint32_t* __restrict__ out_col) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V) return;
    const uint32_t lo = adj_ptr[i];
    out_row_ptr[i] = lo;
    if (i == V) return;
    const uint32_t u = member_vertex(member_keys, i);
    out_node[i] = (int32_t)v_node[u] + min_node;
    out_weight[i] = m_weight[i];
    out_exit[i] = (m_flags[i] >> 1) & 1;
    const uint32_t first = comp_ptr[comp_of_vertex[u]], hi = adj_ptr[i + 1];
    for (uint32_t e = lo; e < hi; ++e) out_col[e] = adj[e] - first;
}

static int max_paths_impl(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                          uint64_t n_runs, const uint32_t* score_pos, const double* score_val,
                          const int32_t* score_val_i32, uint64_t n_score,
                          const uint32_t* q_pos, const double* q_val,
                          uint64_t n_q, int32_t internal_id_shift, int32_t* peak_start,
                          int32_t* peak_end, int32_t* peak_length, double* peak_score,
                          uint8_t* peak_info, uint32_t* peak_node_ptr, uint64_t peak_capacity,
                          int32_t* peak_nodes, uint64_t node_capacity, uint32_t* order,
                          uint64_t* n_peaks, uint64_t* n_nodes_total, const SubgraphOut* sub,
                          cudaStream_t stream) {
    *n_peaks = 0;
    *n_nodes_total = 0;
    if (n_runs == 0) { set_error("empty track"); return GPC_ERR_ARG; }
    const uint32_t n = graph->n_nodes;
    uint8_t v_first = 0, v_last = 0;
    GPC_READ_BACK({val, &v_first, (unsigned)(1)}, {val + n_runs - 1, &v_last, (unsigned)(1)});
    // get_segments (maxpaths.py:43-50)
    const uint32_t first = v_first ? 0 : 1;
    const uint32_t n_idx = (uint32_t)n_runs - first + (v_last ? 1 : 0);
    const uint32_t S = n_idx / 2;
    if (S == 0) return GPC_OK;
    Segments sg{pos, first, S, (uint32_t)n_runs, graph->size_bp};
    Scratch v0, v1, cnt, multi, between, multi_rank, between_offs;
    GPC_CUDA(v0.alloc((size_t)S * 4, stream));
    GPC_CUDA(v1.alloc((size_t)S * 4, stream));
    GPC_CUDA(cnt.alloc((size_t)S * 4, stream));
    GPC_CUDA(multi.alloc((size_t)S * 4, stream));
    GPC_CUDA(between.alloc((size_t)S * 4, stream));
    GPC_CUDA(multi_rank.alloc((size_t)S * 4, stream));
    GPC_CUDA(between_offs.alloc((size_t)S * 4, stream));
    GPC_PROF("segment_nodes");
    segment_nodes_kernel<<<GRID(S)>>>(*graph, sg, v0.as<uint32_t>(), v1.as<uint32_t>(),
                                      cnt.as<uint32_t>(), multi.as<uint32_t>(),
                                      between.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    unsigned long long n_multi = 0, n_between = 0;
    {
        const uint32_t* in[2] = {multi.as<uint32_t>(), between.as<uint32_t>()};
        uint32_t* out[2] = {multi_rank.as<uint32_t>(), between_offs.as<uint32_t>()};
        unsigned long long tot[2];
        GPC_TRY(multi_scan(2, in, out, S, tot, stream));
        n_multi = tot[0];
        n_between = tot[1];
    }
    const uint32_t P = S + (uint32_t)n_multi + (uint32_t)n_between;
    Scratch p_node, p_start, p_end, p_type, f_t, f_w, f_h, f_i, r_t, r_w, r_h, r_i;
    GPC_CUDA(p_node.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_start.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_end.alloc((size_t)P * 4, stream));
    GPC_CUDA(p_type.alloc(P, stream));
    GPC_PROF("peak_pieces");
    peak_pieces_kernel<<<GRID(S)>>>(*graph, sg, v0.as<uint32_t>(), v1.as<uint32_t>(),
                                    multi_rank.as<uint32_t>(), between_offs.as<uint32_t>(),
                                    (uint32_t)n_multi, p_node.as<uint32_t>(), p_start.as<uint32_t>(),
                                    p_end.as<uint32_t>(), p_type.as<uint8_t>());
    GPC_LAUNCH_CHECK();
    for (Scratch* s : {&f_t, &f_w, &f_h, &f_i, &r_t, &r_w, &r_h, &r_i})
        GPC_CUDA(s->alloc((size_t)P * 4, stream));
    GPC_PROF("class_flags");
    class_flags_kernel<<<GRID(P)>>>(P, p_type.as<uint8_t>(), f_t.as<uint32_t>(), f_w.as<uint32_t>(),
                                    f_h.as<uint32_t>(), f_i.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    unsigned long long n_t = 0, n_w = 0, n_h = 0, n_i = 0;
    {
        const uint32_t* in[4] = {f_t.as<uint32_t>(), f_w.as<uint32_t>(), f_h.as<uint32_t>(),
                                 f_i.as<uint32_t>()};
        uint32_t* out[4] = {r_t.as<uint32_t>(), r_w.as<uint32_t>(), r_h.as<uint32_t>(),
                            r_i.as<uint32_t>()};
        unsigned long long tot[4];
        GPC_TRY(multi_scan(4, in, out, P, tot, stream));
        n_t = tot[0]; n_w = tot[1]; n_h = tot[2]; n_i = tot[3];
    }
    const uint32_t V = (uint32_t)(n_t + n_w + n_h);
    // internal pieces in piece order
    Scratch internal_piece;
    GPC_CUDA(internal_piece.alloc((size_t)(n_i ? n_i : 1) * 4, stream));
    // vertex tables
    Scratch vertex_of_piece, piece_of_vertex, in_v, out_v, v_node, v_type;
    GPC_CUDA(vertex_of_piece.alloc((size_t)P * 4, stream));
    GPC_CUDA(piece_of_vertex.alloc((size_t)(V ? V : 1) * 4, stream));
    GPC_CUDA(in_v.alloc((size_t)n * 4, stream));
    GPC_CUDA(out_v.alloc((size_t)n * 4, stream));
    GPC_CUDA(cudaMemsetAsync(in_v.p, 0xff, (size_t)n * 4, stream));
    GPC_CUDA(cudaMemsetAsync(out_v.p, 0xff, (size_t)n * 4, stream));
    GPC_PROF("peak_vertices");
    peak_vertices_kernel<<<GRID(P)>>>(P, p_node.as<uint32_t>(), p_type.as<uint8_t>(),
                                      r_t.as<uint32_t>(), r_w.as<uint32_t>(), r_h.as<uint32_t>(),
                                      (uint32_t)n_t, (uint32_t)n_w, vertex_of_piece.as<int32_t>(),
                                      piece_of_vertex.as<uint32_t>(), in_v.as<int32_t>(),
                                      out_v.as<int32_t>());
    GPC_LAUNCH_CHECK();
    GPC_PROF("gather_internal");
    gather_internal_kernel<<<GRID(P)>>>(P, p_type.as<uint8_t>(), r_i.as<uint32_t>(),
                                        internal_piece.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    // piece scores -> int32 weights per vertex (graphs.py:150-151, 183)
    Scratch score, weight;
    GPC_CUDA(score.alloc((size_t)P * 8, stream));
    GPC_PROF("piece_scores");
    if (score_val_i32)
        piece_scores_kernel<int32_t><<<GRID(P)>>>(P, p_start.as<uint32_t>(), p_end.as<uint32_t>(),
                                                  score_pos, score_val_i32, (uint32_t)n_score,
                                                  graph->size_bp, score.as<double>());
    else
        piece_scores_kernel<double><<<GRID(P)>>>(P, p_start.as<uint32_t>(), p_end.as<uint32_t>(),
                                                 score_pos, score_val, (uint32_t)n_score,
                                                 graph->size_bp, score.as<double>());
    GPC_LAUNCH_CHECK();
    Scratch parent, is_entry, is_exit, keys0, keys1, head, hoffs, comp_ptr, local_of, comp_of;
    Scratch dist, dsink, pred, path, path_len, info;
    Scratch adj_deg, adj_ptr, adj, m_weight, m_flags, comp_backward;
    unsigned long long n_comp = 0;
    if (V) {
        GPC_CUDA(v_node.alloc((size_t)V * 4, stream));
        GPC_CUDA(v_type.alloc(V, stream));
        GPC_CUDA(weight.alloc((size_t)V * 4, stream));
        GPC_PROF("vertex_fields");
        vertex_fields_kernel<<<GRID(V)>>>(V, piece_of_vertex.as<uint32_t>(), p_node.as<uint32_t>(),
                                          p_type.as<uint8_t>(), score.as<double>(),
                                          v_node.as<uint32_t>(), v_type.as<uint8_t>(),
                                          weight.as<int32_t>());
        GPC_LAUNCH_CHECK();
        VertexView vw{V, v_node.as<uint32_t>(), v_type.as<uint8_t>(), nullptr, in_v.as<int32_t>(),
                      out_v.as<int32_t>()};
        GPC_CUDA(parent.alloc((size_t)V * 4, stream));
        GPC_CUDA(is_entry.alloc(V, stream));
        GPC_CUDA(is_exit.alloc(V, stream));
        GPC_CUDA(keys0.alloc((size_t)V * 8, stream));
        GPC_CUDA(keys1.alloc((size_t)V * 8, stream));
        GPC_CUDA(head.alloc(V, stream));
        GPC_CUDA(hoffs.alloc((size_t)V * 4, stream));
        GPC_CUDA(local_of.alloc((size_t)V * 4, stream));
        GPC_CUDA(comp_of.alloc((size_t)V * 4, stream));
        GPC_TRY(uf_components(graph, vw, parent.as<uint32_t>(), stream));
        GPC_PROF("peak_flags");
        peak_flags_kernel<<<GRID(V)>>>(*graph, vw, is_entry.as<uint8_t>(), is_exit.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        GPC_PROF("comp_keys");
        comp_keys_kernel<<<GRID(V)>>>(V, parent.as<uint32_t>(), keys0.as<unsigned long long>());
        GPC_LAUNCH_CHECK();
        bool in_tmp = false;
        // keys are (root << 32 | vertex) written in vertex order: a stable sort on the
        // root bits alone gives (root, vertex) order
        int root_bits = 1;
        while ((1ull << root_bits) < (unsigned long long)V) ++root_bits;
        GPC_TRY((radix_sort<unsigned long long, false>(keys0.as<unsigned long long>(),
                                                       keys1.as<unsigned long long>(), nullptr,
                                                       nullptr, V, 32, 32 + root_bits, false,
                                                       &in_tmp, stream)));
        const unsigned long long* mk = in_tmp ? keys1.as<unsigned long long>() : keys0.as<unsigned long long>();
        GPC_PROF("comp_heads");
        comp_heads_kernel<<<GRID(V)>>>(mk, V, head.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        GPC_TRY(scan_flags(head.as<uint8_t>(), hoffs.as<uint32_t>(), V, &n_comp, stream));
        GPC_CUDA(comp_ptr.alloc((size_t)(n_comp + 1) * 4, stream));
        GPC_PROF("comp_ptr");
        comp_ptr_kernel<<<GRID(V)>>>(head.as<uint8_t>(), hoffs.as<uint32_t>(), V, (uint32_t)n_comp,
                                     comp_ptr.as<uint32_t>(), mk, local_of.as<uint32_t>(),
                                     comp_of.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        GPC_CUDA(dist.alloc((size_t)V * 8, stream));
        GPC_CUDA(dsink.alloc((size_t)V * 8, stream));
        GPC_CUDA(pred.alloc((size_t)V * 4, stream));
        GPC_CUDA(path.alloc((size_t)V * 4, stream));
        GPC_CUDA(path_len.alloc((size_t)n_comp * 4, stream));
        GPC_CUDA(info.alloc((size_t)n_comp, stream));
        PathOut po{path.as<uint32_t>(), path_len.as<uint32_t>(), info.as<uint8_t>()};
        // successor lists in member order (at most one list entry per graph edge)
        GPC_CUDA(adj_deg.alloc((size_t)(V + 1) * 4, stream));
        GPC_CUDA(adj_ptr.alloc((size_t)(V + 1) * 4, stream));
        GPC_CUDA(adj.alloc((size_t)graph->n_edges * 4 + 16, stream));
        GPC_CUDA(m_weight.alloc((size_t)V * 4, stream));
        GPC_CUDA(m_flags.alloc(V, stream));
        GPC_CUDA(comp_backward.alloc(V, stream));         // indexed by component label (< V)
        GPC_CUDA(cudaMemsetAsync(comp_backward.p, 0, V, stream));
        GPC_PROF("adj_count");
        adj_count_kernel<<<GRID(V + 1)>>>(*graph, vw, V, mk, adj_deg.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        GPC_TRY(exclusive_scan<uint32_t>(adj_deg.as<uint32_t>(), adj_ptr.as<uint32_t>(), (size_t)V + 1,
                                         nullptr, stream));
        GPC_PROF("adj_fill");
        adj_fill_kernel<<<GRID(V)>>>(*graph, vw, V, mk, local_of.as<uint32_t>(),
                                     adj_ptr.as<uint32_t>(), weight.as<int32_t>(),
                                     is_entry.as<uint8_t>(), is_exit.as<uint8_t>(),
                                     adj.as<uint32_t>(), m_weight.as<int32_t>(),
                                     m_flags.as<uint8_t>(), comp_backward.as<uint8_t>());
        GPC_LAUNCH_CHECK();
        if (sub) {
            uint32_t n_adj = 0;
            GPC_READ_BACK({adj_ptr.as<uint32_t>() + V, &n_adj, 4u});
            *sub->n_comp = n_comp;
            *sub->n_members = V;
            *sub->n_cols = n_adj;
            if (n_comp + 1 > sub->comp_capacity || (uint64_t)V + 1 > sub->member_capacity ||
                n_adj > sub->col_capacity)
                return GPC_ERR_CAPACITY;
            GPC_CUDA(cudaMemcpyAsync(sub->comp_ptr, comp_ptr.p, (size_t)(n_comp + 1) * 4,
                                     cudaMemcpyDeviceToDevice, stream));
            GPC_PROF("subgraph_emit");
            subgraph_emit_kernel<<<GRID((size_t)V + 1)>>>(
                V, graph->min_node, mk, v_node.as<uint32_t>(), comp_of.as<uint32_t>(),
                comp_ptr.as<uint32_t>(), adj_ptr.as<uint32_t>(), adj.as<uint32_t>(),
                m_weight.as<int32_t>(), m_flags.as<uint8_t>(), sub->member_node, sub->member_weight,
                sub->member_exit, sub->row_ptr, sub->col);
            GPC_LAUNCH_CHECK();
            return GPC_OK;
        }
        // members of every component in topological order (sorted by node: ids ascend along
        // every edge), as positions in the member arrays
        Scratch tk0, tk1, tv0, tv1, sweep, ccount, clist;
        GPC_CUDA(tk0.alloc((size_t)V * 8, stream));
        GPC_CUDA(tk1.alloc((size_t)V * 8, stream));
        GPC_CUDA(tv0.alloc((size_t)V * 4, stream));
        GPC_CUDA(tv1.alloc((size_t)V * 4, stream));
        GPC_CUDA(sweep.alloc((size_t)V * 4, stream));
        int node_bits = 1, comp_bits = 1;
        while ((1ull << node_bits) < (unsigned long long)graph->n_nodes) ++node_bits;
        while ((1ull << comp_bits) < n_comp) ++comp_bits;
        GPC_PROF("mp_topo_keys");
        mp_topo_keys_kernel<<<GRID(V)>>>((uint32_t)V, mk, comp_of.as<uint32_t>(), v_node.as<uint32_t>(),
                                         node_bits, tk0.as<unsigned long long>(), tv0.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        bool topo_tmp = false;
        GPC_TRY((radix_sort<unsigned long long, true>(tk0.as<unsigned long long>(), tk1.as<unsigned long long>(),
                                                      tv0.as<uint32_t>(), tv1.as<uint32_t>(), V, 0,
                                                      node_bits + comp_bits, false, &topo_tmp, stream)));
        const uint32_t* topo_member = topo_tmp ? tv1.as<uint32_t>() : tv0.as<uint32_t>();
        GPC_CUDA(ccount.alloc(16, stream));
        GPC_CUDA(cudaMemsetAsync(ccount.p, 0, 16, stream));
        GPC_CUDA(clist.alloc((size_t)n_comp * 4 * 4, stream));
        GPC_PROF("mp_classify");
        mp_classify_kernel<<<GRID(n_comp)>>>((uint32_t)n_comp, comp_ptr.as<uint32_t>(), adj_ptr.as<uint32_t>(),
                                             ccount.as<uint32_t>(), clist.as<uint32_t>());
        GPC_LAUNCH_CHECK();
        // one launch per size class (see max_paths_kernel), each with room for every component
        // (the class sizes stay on the device); the larger classes need the opt-in to more than
        // 48 KB of dynamic shared memory per block
        const unsigned mp_blocks = div_up(n_comp, MPW_WARPS);
#define GPC_MP_LAUNCH(CLS, NAME)                                                                       \
        do {                                                                                           \
            const size_t smem__ = (CLS) == 0 ? 0 : MPW_WARPS * mp_warp_smem(mp_class_cap(CLS));        \
            if (smem__ > 48 * 1024)                                                                    \
                GPC_CUDA(cudaFuncSetAttribute(max_paths_kernel<CLS>,                                   \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem__)); \
            GPC_PROF(NAME);                                                                            \
            max_paths_kernel<CLS><<<mp_blocks, MPW_WARPS * 32, smem__, stream>>>(                      \
                (uint32_t)n_comp, ccount.as<uint32_t>(), clist.as<uint32_t>(), comp_ptr.as<uint32_t>(), \
                mk, adj_ptr.as<uint32_t>(), adj.as<uint32_t>(), m_weight.as<int32_t>(),                \
                m_flags.as<uint8_t>(), topo_member, dist.as<long long>(), dsink.as<long long>(),       \
                pred.as<int32_t>(), sweep.as<int32_t>(), po);                                          \
            GPC_LAUNCH_CHECK();                                                                        \
        } while (0)
        GPC_MP_LAUNCH(1, "max_paths");
        GPC_MP_LAUNCH(2, "max_paths_256");
        GPC_MP_LAUNCH(3, "max_paths_1024");
        GPC_MP_LAUNCH(0, "max_paths_large");
#undef GPC_MP_LAUNCH
    }
    const uint32_t n_pk = (uint32_t)n_comp + (uint32_t)n_i;
    *n_peaks = n_pk;
    if (sub) return GPC_OK;             // (no line-graph component at all)
    if (n_pk == 0) return GPC_OK;
    if (n_pk > peak_capacity) return GPC_ERR_CAPACITY;
    Scratch nn;
    GPC_CUDA(nn.alloc((size_t)n_pk * 4, stream));
    GPC_PROF("peak_lengths");
    peak_lengths_kernel<<<GRID(n_pk)>>>((uint32_t)n_comp, (uint32_t)n_i, path_len.as<uint32_t>(),
                                        nn.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    unsigned long long total_nodes = 0;
    GPC_TRY(scan_counts(nn.as<uint32_t>(), peak_node_ptr, n_pk, &total_nodes, stream));
    *n_nodes_total = total_nodes;
    if (total_nodes > node_capacity) return GPC_ERR_CAPACITY;
    uint32_t tn = (uint32_t)total_nodes;
    GPC_CUDA(cudaMemcpyAsync(peak_node_ptr + n_pk, &tn, 4, cudaMemcpyHostToDevice, stream));
    GPC_PROF("peaks");
    peaks_kernel<<<GRID((size_t)n_pk * 32)>>>(*graph, (uint32_t)n_comp, (uint32_t)n_i, comp_ptr.as<uint32_t>(),
                                 path.as<uint32_t>(), path_len.as<uint32_t>(), info.as<uint8_t>(),
                                 piece_of_vertex.as<uint32_t>(), internal_piece.as<uint32_t>(),
                                 p_node.as<uint32_t>(), p_start.as<uint32_t>(), p_end.as<uint32_t>(),
                                 q_pos, q_val, (uint32_t)n_q, internal_id_shift, peak_node_ptr,
                                 peak_nodes, peak_start, peak_end, peak_length, peak_score,
                                 peak_info);
    GPC_LAUNCH_CHECK();
    // stable order by score, descending (callpeaks.py:204-205)
    Scratch sk0, sk1, si0, si1;
    GPC_CUDA(sk0.alloc((size_t)n_pk * 8, stream));
    GPC_CUDA(sk1.alloc((size_t)n_pk * 8, stream));
    GPC_CUDA(si0.alloc((size_t)n_pk * 4, stream));
    GPC_CUDA(si1.alloc((size_t)n_pk * 4, stream));
    GPC_PROF("score_keys");
    score_keys_kernel<<<GRID(n_pk)>>>(peak_score, n_pk, sk0.as<unsigned long long>(), si0.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    bool in_tmp = false;
    GPC_TRY((radix_sort<unsigned long long, true>(sk0.as<unsigned long long>(),
                                                  sk1.as<unsigned long long>(), si0.as<uint32_t>(),
                                                  si1.as<uint32_t>(), n_pk, 0, 64, true, &in_tmp,
                                                  stream)));
    GPC_CUDA(cudaMemcpyAsync(order, in_tmp ? si1.p : si0.p, (size_t)n_pk * 4,
                             cudaMemcpyDeviceToDevice, stream));
    return GPC_OK;
}

extern "C" int gpc_max_paths(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                             uint64_t n_runs, const uint32_t* score_pos, const double* score_val,
                             const int32_t* score_val_i32, uint64_t n_score,
                             const uint32_t* q_pos, const double* q_val,
                             uint64_t n_q, int32_t internal_id_shift, int32_t* peak_start,
                             int32_t* peak_end, int32_t* peak_length, double* peak_score,
                             uint8_t* peak_info, uint32_t* peak_node_ptr, uint64_t peak_capacity,
                             int32_t* peak_nodes, uint64_t node_capacity, uint32_t* order,
                             uint64_t* n_peaks, uint64_t* n_nodes_total, void* stream_) {
    GPC_STAGE();
    return max_paths_impl(graph, pos, val, n_runs, score_pos, score_val, score_val_i32, n_score, q_pos,
                          q_val, n_q, internal_id_shift, peak_start, peak_end, peak_length, peak_score,
                          peak_info, peak_node_ptr, peak_capacity, peak_nodes, node_capacity, order,
                          n_peaks, n_nodes_total, nullptr, (cudaStream_t)stream_);
}

extern "C" int gpc_peak_subgraphs(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                                  uint64_t n_runs, const uint32_t* score_pos, const double* score_val,
                                  const int32_t* score_val_i32, uint64_t n_score, uint32_t* comp_ptr,
                                  uint64_t comp_capacity, int32_t* member_node, int32_t* member_weight,
                                  uint8_t* member_exit, uint32_t* row_ptr, uint64_t member_capacity,
                                  uint32_t* col, uint64_t col_capacity, uint64_t* n_comp,
                                  uint64_t* n_members, uint64_t* n_cols, void* stream_) {
    GPC_STAGE();
    *n_comp = *n_members = *n_cols = 0;
    SubgraphOut sub{comp_ptr, comp_capacity, member_node, member_weight, member_exit, row_ptr,
                   member_capacity, col, col_capacity, n_comp, n_members, n_cols};
    uint64_t n_peaks = 0, n_nodes_total = 0;
    return max_paths_impl(graph, pos, val, n_runs, score_pos, score_val, score_val_i32, n_score, nullptr,
                          nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0,
                          nullptr, &n_peaks, &n_nodes_total, &sub, (cudaStream_t)stream_);
}
