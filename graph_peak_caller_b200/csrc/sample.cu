// Sample (ChIP) pileup on sm_100a.
//
// Replaces, for the whole read set at once, the reference's
//   ReadsAdderWDirect.add_reads        sample/sparsegraphpileup.py:37-119
//   SparseExtender.run_linear          sample/sparsegraphpileup.py:151-184
//   ReverseSparseExtender              sample/sparsegraphpileup.py:190-211
//   SparseDiffs.from_pileup + get_sparse_values   sparsediffs.py:137-141,180-190
//   UniqueIntervals                    intervals.py:17-33
//
// The reference sweeps the nodes in topological order with a
// {read id -> remaining} dict per node.  Here every read is walked on its own:
// its body pieces, then a small frontier {node -> max remaining} popped in id
// (= topological) order; contiguous pieces are merged before they are emitted
// as packed +1/-1 events.  One event stream feeds both the fragment pileup and
// the direct (reads only) pileup; one radix sort and one chained scan turn it
// into the two canonical run-length tracks.
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

#include <mutex>

namespace gpc {

// event key = position << 3 | tag << 1 | (delta is -1)
constexpr unsigned TAG_BOTH = 0, TAG_DIRECT = 1, TAG_FRAG = 2;
constexpr int EVENT_POS_SHIFT = 3;

constexpr int WALK_THREADS = 128;
constexpr int WALK_STAGE = 1536;     // staged events per block (6 KB)
constexpr int FRONTIER_CAP = 32;

// Where the +1 / -1 steps of the walk go.  DENSE == false: packed event keys,
// staged per block in shared memory, appended to one global list that is then
// sorted (sparse read sets: whole-genome depth).  DENSE == true: straight into
// a difference array over the graph's positions, one 64-bit word per position
// holding both pileups (high word fragment, low word direct; the borrow of a
// negative low word is undone when unpacking) -- one fire-and-forget 64-bit
// reduction per step, no sort (read sets with more than ~0.1 events per base).
template <bool DENSE>
struct Emitter {
    uint32_t* s_buf;
    unsigned* s_count;
    uint32_t* g_events;
    unsigned long long* g_count;
    unsigned long long g_cap;
    unsigned long long* dense;
    uint32_t G;
    bool fwd;
    unsigned n_emitted;

    // +1 (open) or -1 (close) at oriented position p
    __device__ __forceinline__ void emit(uint32_t p, bool close, unsigned tag) {
        uint32_t pos = fwd ? p : (G - p);
        bool minus = fwd ? close : !close;
        if (DENSE) {
            const long long d = minus ? -1ll : 1ll;
            const long long add = tag == TAG_BOTH ? d * ((1ll << 32) + 1ll)
                                                  : (tag == TAG_DIRECT ? d : d * (1ll << 32));
            atomicAdd(dense + pos, (unsigned long long)add);
            ++n_emitted;
            return;
        }
        uint32_t key = (pos << EVENT_POS_SHIFT) | (tag << 1) | (minus ? 1u : 0u);
        unsigned slot = atomicAdd(s_count, 1u);
        if (slot < WALK_STAGE) {
            s_buf[slot] = key;
        } else {   // staging full: append straight to the global stream
            unsigned long long at = atomicAdd(g_count, 1ull);
            if (at < g_cap) g_events[at] = key;
        }
    }
};

// Everything the walk needs to know about one node, in its direction of travel.
struct WalkNode {
    uint32_t off, len;
    uint32_t n_adj;        // successors (forward) / predecessors (reverse)
    uint32_t adj0, adj1;   // the first two of them
};
__device__ __forceinline__ WalkNode load_walk_node(const gpc_graph_t& g, uint32_t v, bool fwd) {
    WalkNode w;
    if (g.walk_rec) {
        const uint4* p = reinterpret_cast<const uint4*>(g.walk_rec) + 2ull * v;
        uint4 a = __ldg(p), b = __ldg(p + 1);
        w.off = a.x;
        w.len = a.y;
        w.n_adj = fwd ? (a.z & 0xffffu) : (a.z >> 16);
        w.adj0 = fwd ? a.w : b.y;
        w.adj1 = fwd ? b.x : b.z;
    } else {
        uint4 a = node_record(g, v), b = node_record(g, v + 1);
        w.off = a.x;
        w.len = a.w;
        w.n_adj = fwd ? (b.y - a.y) : (b.z - a.z);
        const uint32_t* adj = fwd ? g.succ + a.y : g.pred + a.z;
        w.adj0 = w.n_adj > 0 ? adj[0] : 0;
        w.adj1 = w.n_adj > 1 ? adj[1] : 0;
    }
    return w;
}

// The extension frontier: pending (node key, bases still to cover) pairs, popped
// in increasing key order (= the reference's sweep over ascending node ids),
// one entry per node holding the largest remainder seen for it.  The first
// four entries live in registers, kept sorted -- the usual frontier of a bubble
// graph is 1-3 nodes, and the linear scans over a local-memory array were 45 %
// of the kernel's instructions; more entries spill into a local array.
struct FrontierSpill {                      // (a separate object: a struct holding a
    uint32_t key[FRONTIER_CAP], rem[FRONTIER_CAP];   //  dynamically indexed array lives in local memory as a whole)
};
// More than 4 + FRONTIER_CAP pending nodes (deeply nested bubbles with long branches): the
// spill list moves, once, into a chunk of a global pool that the launch provides
// (SPILL_CHUNK entries; the pool is handed out by an atomic counter).  Only a read that
// outgrows that too, or finds the pool empty, fails with GPC_ERR_FRONTIER_OVERFLOW.
constexpr int SPILL_CHUNK = 4096;
struct SpillPool {
    uint32_t* base;                 // n_chunks * 2 * SPILL_CHUNK words (keys, then remainders)
    unsigned* next;                 // chunks handed out so far
    unsigned n_chunks;
};

// Out of line on purpose (and pointers in, pointer out: nothing of the frontier has its
// address taken, so its state stays in registers): inlined at the three push sites the grow
// path cost the walk 10-17 % of its speed although it never ran.
__device__ __noinline__ uint32_t* spill_grow(const uint32_t* keys, const uint32_t* rems, int n,
                                             SpillPool pool) {
    if (!pool.base) return nullptr;
    const unsigned c = atomicAdd(pool.next, 1u);
    if (c >= pool.n_chunks) return nullptr;
    uint32_t* nk = pool.base + (size_t)c * 2 * SPILL_CHUNK;
    uint32_t* nr = nk + SPILL_CHUNK;
    for (int f = 0; f < n; ++f) { nk[f] = keys[f]; nr[f] = rems[f]; }
    return nk;
}

template <bool BIG>
struct Frontier {
    uint32_t k0, k1, k2, k3, r0, r1, r2, r3;
    int n;                                  // entries in registers (sorted by key)
    int xn;                                 // spilled entries (unsorted)
    int state;                              // bit 0: overflow, bit 1: spill list sits in the pool
    uint32_t* xkey;
    uint32_t* xrem;

    __device__ __forceinline__ void init(FrontierSpill& sp) {
        n = 0; xn = 0; state = 0;
        k0 = k1 = k2 = k3 = 0; r0 = r1 = r2 = r3 = 0;
        xkey = sp.key; xrem = sp.rem;
    }
    __device__ __forceinline__ bool empty() const { return n == 0 && xn == 0; }

    __device__ __forceinline__ void push(uint32_t key, uint32_t rem, const SpillPool& pool) {
        // a node that sits in the spill list must be found there before any of the
        // register paths may insert it (rare: only after more than four were pending)
        if (xn > 0) {
            for (int f = 0; f < xn; ++f)
                if (xkey[f] == key) { xrem[f] = max(xrem[f], rem); return; }
        }
        // the usual cases (a bubble graph keeps one or two nodes pending) first,
        // each with the few moves it needs
        if (n == 0) { k0 = key; r0 = rem; n = 1; return; }
        if (k0 == key) { r0 = max(r0, rem); return; }
        if (n == 1) {
            if (key < k0) { k1 = k0; r1 = r0; k0 = key; r0 = rem; }
            else { k1 = key; r1 = rem; }
            n = 2;
            return;
        }
        if (k1 == key) { r1 = max(r1, rem); return; }
        if (n == 2) {
            if (key < k0) { k2 = k1; r2 = r1; k1 = k0; r1 = r0; k0 = key; r0 = rem; }
            else if (key < k1) { k2 = k1; r2 = r1; k1 = key; r1 = rem; }
            else { k2 = key; r2 = rem; }
            n = 3;
            return;
        }
        if (k2 == key) { r2 = max(r2, rem); return; }
        if (n > 3 && k3 == key) { r3 = max(r3, rem); return; }
        if (n == 3) {
            k3 = key; r3 = rem; n = 4;
            uint32_t t;
            if (k3 < k2) { t = k2; k2 = k3; k3 = t; t = r2; r2 = r3; r3 = t; }
            if (k2 < k1) { t = k1; k1 = k2; k2 = t; t = r1; r1 = r2; r2 = t; }
            if (k1 < k0) { t = k0; k0 = k1; k1 = t; t = r0; r0 = r1; r1 = t; }
        } else if (!BIG) {
            // the fast kernel gives up here; the host then runs the launch again with BIG
            if (xn < FRONTIER_CAP) { xkey[xn] = key; xrem[xn] = rem; ++xn; }
            else state |= 1;
        } else {
            if (xn >= ((state & 2) ? SPILL_CHUNK : FRONTIER_CAP)) {
                uint32_t* nk = (state & 2) ? nullptr : spill_grow(xkey, xrem, xn, pool);
                if (!nk) { state |= 1; return; }
                xkey = nk; xrem = nk + SPILL_CHUNK; state |= 2;
            }
            xkey[xn] = key; xrem[xn] = rem; ++xn;
        }
    }

    // smallest key (must not be empty)
    __device__ __forceinline__ void pop(uint32_t& key, uint32_t& rem) {
        int best = -1;
        if (xn > 0) {
            best = 0;
            for (int f = 1; f < xn; ++f) if (xkey[f] < xkey[best]) best = f;
            if (n > 0 && k0 < xkey[best]) best = -1;
        }
        if (best < 0) {
            key = k0; rem = r0;
            k0 = k1; r0 = r1; k1 = k2; r1 = r2; k2 = k3; r2 = r3;
            --n;
        } else {
            key = xkey[best]; rem = xrem[best];
            --xn;
            xkey[best] = xkey[xn]; xrem[best] = xrem[xn];
        }
    }
};

template <bool DENSE, bool BIG>
__global__ void __launch_bounds__(WALK_THREADS, 10)
walk_reads_kernel(gpc_graph_t graph, gpc_reads_t reads, const uint32_t* __restrict__ read_index,
                  unsigned n_reads, int extension, uint32_t* __restrict__ events,
                  unsigned long long event_cap, unsigned long long* __restrict__ event_count,
                  unsigned long long* __restrict__ dense,
                  uint8_t* __restrict__ touched, int* __restrict__ err, SpillPool pool) {
    __shared__ uint32_t s_buf[DENSE ? 1 : WALK_STAGE];
    __shared__ unsigned s_count;
    __shared__ unsigned long long s_gbase;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();

    const unsigned i = blockIdx.x * WALK_THREADS + threadIdx.x;
    unsigned n_emitted = 0;
    if (i < n_reads) {
        const unsigned r = read_index ? read_index[i] : i;
        int so, eo;
        long long p0, p1;
        bool fwd, ok;
        if (reads.hdr) {
            // one 16-byte record per read {path offset, path nodes | bad << 31, start
            // offset, end offset}: a walk in graph order gathers one sector per read
            // instead of six (the start / end node are the ends of the path)
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(reads.hdr) + r);
            p0 = (long long)h.x;
            p1 = p0 + (long long)(h.y & 0x7fffffffu);
            so = (int)h.z;
            eo = (int)h.w;
            ok = !(h.y >> 31) && p1 > p0 && so >= 0 && eo >= 0;
            fwd = ok ? reads.path[p0] > 0 : true;
        } else {
            const int sn = reads.start_node[r], en = reads.end_node[r];
            so = reads.start_off[r];
            eo = reads.end_off[r];
            p0 = reads.path_ptr32 ? (long long)reads.path_ptr32[r] : reads.path_ptr[r];
            p1 = reads.path_ptr32 ? (long long)reads.path_ptr32[r + 1] : reads.path_ptr[r + 1];
            fwd = sn > 0;
            ok = (p1 > p0) && ((en > 0) == fwd) && so >= 0 && eo >= 0;
        }
        const uint32_t n = graph.n_nodes, G = graph.size_bp;
        const int mn = graph.min_node;
        if (ok) {
            for (long long q = p0; q < p1; ++q) {
                int id = reads.path[q];
                long long v = (long long)(id < 0 ? -id : id) - mn;
                if ((id > 0) != fwd || v < 0 || v >= (long long)n) ok = false;
            }
        }
        if (!ok) {
            atomicExch(err, GPC_ERR_INVALID_INTERVAL);
        } else {
            Emitter<DENSE> em{s_buf, &s_count, events, event_count, event_cap, dense, G, fwd, 0u};
            // ---- read body --------------------------------------------------
            uint32_t ce = 0;          // oriented end of the currently open interval
            uint32_t last_len = 0, last_start = 0;
            uint32_t last_v = 0;
            WalkNode cur{0, 0, 0, 0, 0};      // the node whose neighbours are pushed next
            const int k = (int)(p1 - p0);
            for (int j = 0; j < k; ++j) {
                int id = reads.path[p0 + j];
                uint32_t v = (uint32_t)((id < 0 ? -id : id) - mn);
                WalkNode wn = load_walk_node(graph, v, fwd);
                uint32_t len = wn.len;
                uint32_t ns = fwd ? wn.off : (G - wn.off - len);
                uint32_t ps = (j == 0) ? ns + (uint32_t)so : ns;
                touched[v] = 1;
                if (j == 0) {
                    em.emit(ps, false, TAG_BOTH);
                } else if (ps != ce) {
                    em.emit(ce, true, TAG_BOTH);
                    em.emit(ps, false, TAG_BOTH);
                }
                ce = ns + len;
                last_len = len; last_start = ns; last_v = v;
                cur = wn;
            }
            // ---- read end: direct pileup closes here, fragment goes on ------
            em.emit(last_start + (uint32_t)eo, true, TAG_DIRECT);
            uint32_t total = (uint32_t)eo + (uint32_t)extension;
            ce = last_start + min(total, last_len);
            // ---- extension frontier ----------------------------------------
            FrontierSpill spill;
            Frontier<BIG> fr;
            fr.init(spill);
            uint32_t cur_v = last_v;
            uint32_t carry = total > last_len ? total - last_len : 0;
            while (true) {
                if (carry) {
                    // neighbours of cur_v: inline in its walk record, CSR beyond two
                    if (cur.n_adj > 2) {
                        uint4 a = node_record(graph, cur_v);
                        const uint32_t* adj = (fwd ? graph.succ + a.y : graph.pred + a.z);
                        for (uint32_t e = 0; e < cur.n_adj; ++e) {
                            uint32_t wv = adj[e];
                            fr.push(fwd ? wv : ~wv, carry, pool);
                        }
                    } else {
                        if (cur.n_adj > 0) fr.push(fwd ? cur.adj0 : ~cur.adj0, carry, pool);
                        if (cur.n_adj > 1) fr.push(fwd ? cur.adj1 : ~cur.adj1, carry, pool);
                    }
                }
                if (fr.empty()) break;
                uint32_t key, rem;
                fr.pop(key, rem);
                cur_v = fwd ? key : ~key;
                cur = load_walk_node(graph, cur_v, fwd);
                uint32_t len = cur.len;
                uint32_t ns = fwd ? cur.off : (G - cur.off - len);
                touched[cur_v] = 1;
                if (ns != ce) {
                    em.emit(ce, true, TAG_FRAG);
                    em.emit(ns, false, TAG_FRAG);
                }
                ce = ns + min(rem, len);
                carry = rem > len ? rem - len : 0;
            }
            const bool overflow = (fr.state & 1) != 0;
            em.emit(ce, true, TAG_FRAG);
            if (overflow) atomicExch(err, GPC_ERR_FRONTIER_OVERFLOW);
            n_emitted = em.n_emitted;
        }
    }
    if (DENSE) {        // only the number of steps is reported (one add per block)
        n_emitted = warp_sum(n_emitted);
        if (lane_id() == 0 && n_emitted) atomicAdd(&s_count, n_emitted);
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(event_count, (unsigned long long)s_count);
        return;
    }
    __syncthreads();
    const unsigned staged = min(s_count, (unsigned)WALK_STAGE);
    if (threadIdx.x == 0) s_gbase = atomicAdd(event_count, (unsigned long long)staged);
    __syncthreads();
    const unsigned long long gb = s_gbase;
    for (unsigned t = threadIdx.x; t < staged; t += WALK_THREADS)
        if (gb + t < event_cap) events[gb + t] = s_buf[t];
}


// ---- graphs whose ids do not ascend along every edge (cycles) ---------------------
//
// The id-ordered frontier above needs a topological numbering.  Without one a read is
// extended the way the reference's legacy path does it (legacy/extender.py:273-284 over
// obg's GraphTraverser.extend_from_block): a worklist with visited[node] = the largest
// number of bases still to cover on arrival; a node is covered [0, min(remaining, size))
// and passes remaining - size on to its successors while that is positive.  What one
// read covers is a SET of positions (legacy/areas.py BinaryContinousAreas: full nodes,
// prefixes, suffixes): pieces of the read body and of the extension that overlap on a
// node -- the extension wrapping around a cycle into the read -- count once.  Known
// answers: the reference's tests/legacy/test_reversal.py:49-70.
constexpr int CYC_CAP = 48;            // intervals / visited nodes / stack entries per read

struct CycIntervals {
    uint32_t node[CYC_CAP], a[CYC_CAP], b[CYC_CAP];
    uint8_t body[CYC_CAP];
    int n;
    bool overflow;
    __device__ __forceinline__ void add(uint32_t v, uint32_t lo, uint32_t hi, bool is_body) {
        if (hi <= lo) return;
        if (n >= CYC_CAP) { overflow = true; return; }
        node[n] = v; a[n] = lo; b[n] = hi; body[n] = is_body ? 1 : 0;
        ++n;
    }
};

// emits the union of the intervals [first, last) of `iv` (sorted by node, then start) that
// pass `only_body`, as +1 / -1 steps with `tag`
template <bool DENSE>
__device__ __forceinline__ void cyc_emit_union(const gpc_graph_t& graph, bool fwd, const CycIntervals& iv,
                                               bool only_body, unsigned tag, Emitter<DENSE>& em) {
    const uint32_t G = graph.size_bp;
    int i = 0;
    while (i < iv.n) {
        if (only_body && !iv.body[i]) { ++i; continue; }
        const uint32_t v = iv.node[i];
        uint32_t lo = iv.a[i], hi = iv.b[i];
        int j = i + 1;
        for (; j < iv.n && iv.node[j] == v; ++j) {
            if (only_body && !iv.body[j]) continue;
            if (iv.a[j] <= hi) { hi = max(hi, iv.b[j]); continue; }
            const WalkNode w = load_walk_node(graph, v, fwd);
            const uint32_t ns = fwd ? w.off : (G - w.off - w.len);
            em.emit(ns + lo, false, tag);
            em.emit(ns + hi, true, tag);
            lo = iv.a[j]; hi = iv.b[j];
        }
        const WalkNode w = load_walk_node(graph, v, fwd);
        const uint32_t ns = fwd ? w.off : (G - w.off - w.len);
        em.emit(ns + lo, false, tag);
        em.emit(ns + hi, true, tag);
        i = j;
    }
}

template <bool DENSE>
__global__ void __launch_bounds__(WALK_THREADS)
walk_reads_cyclic_kernel(gpc_graph_t graph, gpc_reads_t reads, const uint32_t* __restrict__ read_index,
                         unsigned n_reads, int extension, uint32_t* __restrict__ events,
                         unsigned long long event_cap, unsigned long long* __restrict__ event_count,
                         unsigned long long* __restrict__ dense,
                         uint8_t* __restrict__ touched, int* __restrict__ err) {
    __shared__ uint32_t s_buf[DENSE ? 1 : WALK_STAGE];
    __shared__ unsigned s_count;
    __shared__ unsigned long long s_gbase;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const unsigned i = blockIdx.x * WALK_THREADS + threadIdx.x;
    unsigned n_emitted = 0;
    if (i < n_reads) {
        const unsigned r = read_index ? read_index[i] : i;
        int so, eo;
        long long p0, p1;
        bool fwd, ok;
        if (reads.hdr) {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(reads.hdr) + r);
            p0 = (long long)h.x;
            p1 = p0 + (long long)(h.y & 0x7fffffffu);
            so = (int)h.z;
            eo = (int)h.w;
            ok = !(h.y >> 31) && p1 > p0 && so >= 0 && eo >= 0;
            fwd = ok ? reads.path[p0] > 0 : true;
        } else {
            const int sn = reads.start_node[r], en = reads.end_node[r];
            so = reads.start_off[r];
            eo = reads.end_off[r];
            p0 = reads.path_ptr32 ? (long long)reads.path_ptr32[r] : reads.path_ptr[r];
            p1 = reads.path_ptr32 ? (long long)reads.path_ptr32[r + 1] : reads.path_ptr[r + 1];
            fwd = sn > 0;
            ok = (p1 > p0) && ((en > 0) == fwd) && so >= 0 && eo >= 0;
        }
        const uint32_t n = graph.n_nodes;
        const int mn = graph.min_node;
        if (ok) {
            for (long long q = p0; q < p1; ++q) {
                int id = reads.path[q];
                long long v = (long long)(id < 0 ? -id : id) - mn;
                if ((id > 0) != fwd || v < 0 || v >= (long long)n) ok = false;
            }
        }
        if (!ok) {
            atomicExch(err, GPC_ERR_INVALID_INTERVAL);
        } else {
            CycIntervals iv;
            iv.n = 0;
            iv.overflow = false;
            // ---- read body: [start offset, size) on the first node, whole nodes between,
            // [0, end offset) on the last (one node: [start offset, end offset)) ---------
            const int k = (int)(p1 - p0);
            uint32_t last_v = 0, last_len = 0;
            for (int j = 0; j < k; ++j) {
                int id = reads.path[p0 + j];
                uint32_t v = (uint32_t)((id < 0 ? -id : id) - mn);
                const WalkNode w = load_walk_node(graph, v, fwd);
                touched[v] = 1;
                const uint32_t lo = j == 0 ? (uint32_t)so : 0u;
                const uint32_t hi = j == k - 1 ? min((uint32_t)eo, w.len) : w.len;
                iv.add(v, lo, hi, true);
                last_v = v;
                last_len = w.len;
            }
            // ---- extension: the rest of the last node, then the worklist -----------------
            const uint32_t total = (uint32_t)eo + (uint32_t)extension;
            iv.add(last_v, min((uint32_t)eo, last_len), min(total, last_len), false);
            uint32_t vis_node[CYC_CAP], vis_rem[CYC_CAP], st_node[CYC_CAP], st_rem[CYC_CAP];
            int n_vis = 0, n_st = 0;
            bool overflow = false;
            auto push_adjacent = [&](uint32_t v, uint32_t rem) {
                const WalkNode w = load_walk_node(graph, v, fwd);
                const uint4 a = node_record(graph, v);
                const uint32_t* adj = fwd ? graph.succ + a.y : graph.pred + a.z;
                for (uint32_t e = 0; e < w.n_adj; ++e) {
                    const uint32_t t = w.n_adj > 2 ? adj[e] : (e == 0 ? w.adj0 : w.adj1);
                    if (n_st >= CYC_CAP) { overflow = true; return; }
                    st_node[n_st] = t; st_rem[n_st] = rem; ++n_st;
                }
            };
            if (total > last_len) push_adjacent(last_v, total - last_len);
            while (n_st > 0) {
                --n_st;
                const uint32_t v = st_node[n_st], rem = st_rem[n_st];
                int at = -1;
                for (int f = 0; f < n_vis; ++f) if (vis_node[f] == v) { at = f; break; }
                if (at >= 0 && vis_rem[at] >= rem) continue;          // been here with more to spend
                if (at < 0) {
                    if (n_vis >= CYC_CAP) { overflow = true; break; }
                    at = n_vis++;
                    vis_node[at] = v;
                }
                vis_rem[at] = rem;
                const WalkNode w = load_walk_node(graph, v, fwd);
                if (rem > w.len) push_adjacent(v, rem - w.len);
                if (overflow) break;
            }
            for (int f = 0; f < n_vis; ++f) {
                const WalkNode w = load_walk_node(graph, vis_node[f], fwd);
                touched[vis_node[f]] = 1;
                iv.add(vis_node[f], 0u, min(vis_rem[f], w.len), false);
            }
            if (overflow || iv.overflow) {
                atomicExch(err, GPC_ERR_FRONTIER_OVERFLOW);
            } else {
                // sort by (node, start): a few entries, insertion sort
                for (int x = 1; x < iv.n; ++x) {
                    const uint32_t nv = iv.node[x], na = iv.a[x], nb = iv.b[x];
                    const uint8_t nbody = iv.body[x];
                    int y = x - 1;
                    while (y >= 0 && (iv.node[y] > nv || (iv.node[y] == nv && iv.a[y] > na))) {
                        iv.node[y + 1] = iv.node[y]; iv.a[y + 1] = iv.a[y]; iv.b[y + 1] = iv.b[y];
                        iv.body[y + 1] = iv.body[y];
                        --y;
                    }
                    iv.node[y + 1] = nv; iv.a[y + 1] = na; iv.b[y + 1] = nb; iv.body[y + 1] = nbody;
                }
                Emitter<DENSE> em{s_buf, &s_count, events, event_count, event_cap, dense, graph.size_bp,
                                  fwd, 0u};
                cyc_emit_union<DENSE>(graph, fwd, iv, false, TAG_FRAG, em);
                cyc_emit_union<DENSE>(graph, fwd, iv, true, TAG_DIRECT, em);
                n_emitted = em.n_emitted;
            }
        }
    }
    if (DENSE) {
        n_emitted = warp_sum(n_emitted);
        if (lane_id() == 0 && n_emitted) atomicAdd(&s_count, n_emitted);
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(event_count, (unsigned long long)s_count);
        return;
    }
    __syncthreads();
    const unsigned staged = min(s_count, (unsigned)WALK_STAGE);
    if (threadIdx.x == 0) s_gbase = atomicAdd(event_count, (unsigned long long)staged);
    __syncthreads();
    const unsigned long long gb = s_gbase;
    for (unsigned t = threadIdx.x; t < staged; t += WALK_THREADS)
        if (gb + t < event_cap) events[gb + t] = s_buf[t];
}

// ---- events -> two canonical run-length tracks ------------------------------
//
// Input: event keys sorted ascending.  For each track (0 fragment, 1 direct):
// delta of an event = +-1 if its tag belongs to the track, else 0.  An entry
// (position, running sum) is written after the last event of every position
// whose net delta is non-zero, plus always one entry at position 0 -- the
// canonical SparseValues of the reference (sparsediffs.py:13-20).

constexpr int RUN_THREADS = 256;
constexpr int RUN_ITEMS = 8;
constexpr int RUN_TILE = RUN_THREADS * RUN_ITEMS;

struct RunAgg {              // tile aggregate for the first chained scan
    int tot[2];              // sum of deltas
    int seg[2];              // sum of deltas after the last position change in the tile
    int has_head;            // tile contains a position change
    int pad[3];
};

struct Seg { int head; int a, b; };
__device__ __forceinline__ Seg seg_combine(const Seg& x, const Seg& y) {
    if (y.head) return y;
    return Seg{x.head, x.a + y.a, x.b + y.b};
}

__device__ __forceinline__ void event_delta(uint32_t key, int& dfrag, int& ddir) {
    int s = (key & 1u) ? -1 : 1;
    unsigned tag = (key >> 1) & 3u;
    dfrag = (tag != TAG_DIRECT) ? s : 0;
    ddir = (tag != TAG_FRAG) ? s : 0;
}

__global__ void __launch_bounds__(RUN_THREADS)
events_to_runs_kernel(const uint32_t* __restrict__ keys, unsigned n,
                      uint32_t* __restrict__ frag_idx, int32_t* __restrict__ frag_val,
                      uint32_t* __restrict__ dir_idx, int32_t* __restrict__ dir_val,
                      unsigned long long* __restrict__ st_tot /* zeroed */,
                      unsigned long long* __restrict__ st_seg /* zeroed */,
                      unsigned long long* __restrict__ st2 /* zeroed */,
                      unsigned long long* __restrict__ st3 /* zeroed */,
                      unsigned* __restrict__ ticket, unsigned long long* __restrict__ out_counts,
                      unsigned long long cap) {
    __shared__ unsigned s_slot;
    __shared__ long long s_scan64[33];
    __shared__ Seg s_seg[RUN_THREADS / 32];
    __shared__ RunAgg s_carry;
    __shared__ Seg s_all;
    __shared__ unsigned long long s_obase[2];
    const unsigned tile = claim_tile(ticket, &s_slot);
    const unsigned base = tile * RUN_TILE + threadIdx.x * RUN_ITEMS;

    uint32_t key[RUN_ITEMS + 1];
    bool valid[RUN_ITEMS];
#pragma unroll
    for (int j = 0; j <= RUN_ITEMS; ++j)
        key[j] = (base + j < n) ? keys[base + j] : 0xffffffffu;
    uint32_t prev_key = (base > 0 && base - 1 < n) ? keys[base - 1] : 0xffffffffu;

    // thread-local pass: totals and the segmented (per position) partial sums
    int tf = 0, td = 0;
    Seg seg{0, 0, 0};
    uint32_t ppos = prev_key >> EVENT_POS_SHIFT;
#pragma unroll
    for (int j = 0; j < RUN_ITEMS; ++j) {
        valid[j] = base + j < n;
        if (valid[j]) {
            int df, dd;
            event_delta(key[j], df, dd);
            uint32_t pos = key[j] >> EVENT_POS_SHIFT;
            bool head = (base + j == 0) || pos != ppos;
            seg = seg_combine(seg, Seg{head ? 1 : 0, df, dd});
            tf += df; td += dd;
            ppos = pos;
        }
    }
    // block scan of totals (both tracks in one 64-bit scan: high word fragment,
    // low word direct; the borrow of a negative low word is undone when unpacking)
    int tile_f, tile_d, ex_f, ex_d;
    {
        long long tile_pk;
        long long ex_pk = block_excl_sum<long long>(((long long)tf << 32) + (long long)td, s_scan64,
                                                    tile_pk);
        ex_d = (int)(unsigned)(ex_pk & 0xffffffffll);
        ex_f = (int)((ex_pk - (long long)ex_d) >> 32);
        tile_d = (int)(unsigned)(tile_pk & 0xffffffffll);
        tile_f = (int)((tile_pk - (long long)tile_d) >> 32);
    }
    // block exclusive scan of the segmented operator
    Seg incl = seg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Seg o;
        o.head = __shfl_up_sync(FULL, incl.head, d);
        o.a = __shfl_up_sync(FULL, incl.a, d);
        o.b = __shfl_up_sync(FULL, incl.b, d);
        if (lane_id() >= (unsigned)d) incl = seg_combine(o, incl);
    }
    if (lane_id() == 31) s_seg[warp_id()] = incl;
    __syncthreads();
    Seg carry{0, 0, 0};      // everything before this thread inside the tile
    for (unsigned w = 0; w < warp_id(); ++w) carry = seg_combine(carry, s_seg[w]);
    {
        Seg up;
        up.head = __shfl_up_sync(FULL, incl.head, 1);
        up.a = __shfl_up_sync(FULL, incl.a, 1);
        up.b = __shfl_up_sync(FULL, incl.b, 1);
        if (lane_id() > 0) carry = seg_combine(carry, up);
    }
    // tile aggregates: totals go through a packed single-word chained scan; the
    // open-group sums only need the tiles back to the nearest one that contains
    // a position change, so they are published once and never prefixed
    if (threadIdx.x == RUN_THREADS - 1) {
        Seg all = seg_combine(carry, seg);
        s_all = all;
    }
    __syncthreads();
    // the two chains are independent: warp 0 walks back the open-group sums while
    // warp 1 resolves the totals
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        const Seg all = s_all;
        if (lane == 0)
            lb_store(&st_seg[tile], (1ull << 62) | ((unsigned long long)(all.head ? 1 : 0) << 61) |
                                        ((unsigned long long)((unsigned)all.a & 0x3fffffffu) << 30) |
                                        (unsigned long long)((unsigned)all.b & 0x3fffffffu));
        int in_a = 0, in_b = 0;
        int newest = (int)tile - 1;
        while (newest >= 0) {
            int t = newest - (int)lane;
            unsigned long long w = (t >= 0) ? lb_load(&st_seg[t]) : ((1ull << 62) | (1ull << 61));
            unsigned ready = __ballot_sync(FULL, (w >> 62) != 0);
            unsigned heads = __ballot_sync(FULL, (w >> 61) & 1ull);
            int first_head = heads ? (__ffs(heads) - 1) : 32;
            unsigned need = (first_head >= 31) ? FULL : ((2u << first_head) - 1u);
            if ((ready & need) != need) continue;
            int a = ((int)((unsigned)(w >> 30) << 2)) >> 2;       // sign-extend 30 bits
            int b = ((int)((unsigned)w << 2)) >> 2;
            if ((int)lane > first_head) { a = 0; b = 0; }
            in_a += warp_sum(a);
            in_b += warp_sum(b);
            if (first_head < 32) break;
            newest -= 32;
        }
        if (lane == 0) {
            s_carry.seg[0] = in_a;
            s_carry.seg[1] = in_b;
            s_carry.has_head = 0;
        }
    } else if (threadIdx.x < 64) {
        long long packed = ((long long)tile_f << 32) + (long long)tile_d;
        unsigned long long ex = lookback_excl_warp(st_tot, tile, (unsigned long long)packed & LB_MASK);
        if (threadIdx.x == 32) {
            long long v = (long long)(ex << 2) >> 2;              // sign-extend 62 bits
            int d = (int)(unsigned)(v & 0xffffffffll);
            s_carry.tot[1] = d;
            s_carry.tot[0] = (int)((v - (long long)d) >> 32);
        }
    }
    __syncthreads();
    const RunAgg tin = s_carry;
    // second thread-local pass: decide emissions
    int run_f = tin.tot[0] + ex_f, run_d = tin.tot[1] + ex_d;
    Seg sg = carry.head ? carry : Seg{0, carry.a + tin.seg[0], carry.b + tin.seg[1]};
    // entries are kept at their event slot j (static indexing, no local memory);
    // bit j of emit[t] says that slot j of track t holds an entry
    int e_val[2][RUN_ITEMS];
    unsigned emit[2] = {0u, 0u};
    ppos = prev_key >> EVENT_POS_SHIFT;
#pragma unroll
    for (int j = 0; j < RUN_ITEMS; ++j) {
        e_val[0][j] = 0; e_val[1][j] = 0;
        if (valid[j]) {
            int df, dd;
            event_delta(key[j], df, dd);
            uint32_t pos = key[j] >> EVENT_POS_SHIFT;
            bool head = (base + j == 0) || pos != ppos;
            if (head) { sg.a = 0; sg.b = 0; }
            sg.a += df; sg.b += dd;
            run_f += df; run_d += dd;
            bool last = (base + j + 1 >= n) || (key[j + 1] >> EVENT_POS_SHIFT) != pos;
            if (last) {
                if (sg.a != 0 || pos == 0) { e_val[0][j] = run_f; emit[0] |= 1u << j; }
                if (sg.b != 0 || pos == 0) { e_val[1][j] = run_d; emit[1] |= 1u << j; }
            }
            ppos = pos;
        }
    }
    // an entry at position 0 always exists: synthesise it when no event sits there
    const bool synth = (tile == 0 && threadIdx.x == 0 && (key[0] >> EVENT_POS_SHIFT) != 0);
    const int c0 = __popc(emit[0]) + (synth ? 1 : 0), c1 = __popc(emit[1]) + (synth ? 1 : 0);
    // both output offsets through one packed scan and one chained scan (2 x 31 bits)
    long long tile_pk;
    long long o_pk = block_excl_sum<long long>(((long long)c0 << 31) | (long long)c1, s_scan64, tile_pk);
    if (threadIdx.x < 32) {
        unsigned long long b = lookback_excl_warp(st2, tile, (unsigned long long)tile_pk);
        if (threadIdx.x == 0) {
            s_obase[0] = b;
            if (tile == gridDim.x - 1) {
                unsigned long long tot = b + (unsigned long long)tile_pk;
                out_counts[0] = tot >> 31;
                out_counts[1] = tot & 0x7fffffffull;
            }
        }
    }
    __syncthreads();
    const unsigned long long w_pk = s_obase[0] + (unsigned long long)o_pk;
    unsigned long long w0 = w_pk >> 31, w1 = w_pk & 0x7fffffffull;
    if (synth) {
        if (w0 < cap) { frag_idx[w0] = 0; frag_val[w0] = 0; }
        ++w0;
        if (w1 < cap) { dir_idx[w1] = 0; dir_val[w1] = 0; }
        ++w1;
    }
#pragma unroll
    for (int j = 0; j < RUN_ITEMS; ++j) {
        const uint32_t pos = key[j] >> EVENT_POS_SHIFT;
        if ((emit[0] >> j) & 1u) {
            if (w0 < cap) { frag_idx[w0] = pos; frag_val[w0] = e_val[0][j]; }
            ++w0;
        }
        if ((emit[1] >> j) & 1u) {
            if (w1 < cap) { dir_idx[w1] = pos; dir_val[w1] = e_val[1][j]; }
            ++w1;
        }
    }
}


// ---- difference array -> two canonical run-length tracks ----------------------
//
// dense[pos] holds the net step of both pileups at graph position pos (packed as the
// walk deposits it).  One streaming pass: a tile of 4096 positions (32 KB) is loaded
// with 16-byte coalesced loads into shared memory laid out with a 16-byte pad per
// thread row (so that the row reads below are conflict-free), the words are zeroed
// behind the read (the array is clean again for the next call: no memset kernel),
// every thread walks its 16 consecutive positions, tile totals and entry counts go
// through two packed chained scans walked by two warps at once, and an entry
// (position, running sum) is written wherever the net step of a track is non-zero,
// plus always one at position 0 -- the canonical SparseValues of the reference
// (sparsediffs.py:13-20,137-141).  HBM-bound: 8 bytes read + 8 bytes of zeroes per
// position, 8 bytes per entry out.
//
// Measured alternatives (chr22-scale, 53.7 M positions; this kernel: 0.39 ms):
//   8192 positions per block                                    0.41 ms
//   one 512-position tile per WARP, persistent, no block barrier 0.50 ms (8x the look-backs)
//   bulk-copy engine, three tiles in flight per block            0.52 ms (below, GPC_DENSE_SCAN=3)

constexpr int DR_THREADS = 256;

__device__ __forceinline__ void dense_unpack(unsigned long long w, int& dfrag, int& ddir) {
    ddir = (int)(unsigned)(w & 0xffffffffull);
    dfrag = (int)(((long long)w - (long long)ddir) >> 32);
}

template <int ITEMS>
__global__ void __launch_bounds__(DR_THREADS, ITEMS == 16 ? 6 : 3)
dense_runs_block_kernel(unsigned long long* __restrict__ dense, unsigned n_pos,
                        uint32_t* __restrict__ frag_idx, int32_t* __restrict__ frag_val,
                        uint32_t* __restrict__ dir_idx, int32_t* __restrict__ dir_val,
                        unsigned long long cap,
                        unsigned long long* __restrict__ st_tot /* zeroed */,
                        unsigned long long* __restrict__ st_cnt /* zeroed */,
                        unsigned* __restrict__ ticket, unsigned long long* __restrict__ out_counts) {
    constexpr int TILE = DR_THREADS * ITEMS;
    constexpr int ROW = ITEMS * 8 + 16;
    extern __shared__ __align__(16) unsigned char s_rows[];      // DR_THREADS * ROW
    __shared__ long long s_scan64[33];
    __shared__ unsigned s_slot;
    __shared__ unsigned long long s_base[2];
    const unsigned tile = claim_tile(ticket, &s_slot);
    const unsigned tile_base = tile * TILE;
    // all loads of the tile first, then the zeroes: written as load / test / store per unit
    // the compiler kept that order (the stores go to the array the next load reads), i.e.
    // eight dependent DRAM round trips per tile -- 40 % of the kernel's stall samples.  The
    // array is padded to whole tiles (dense_acquire), so every unit may be loaded in full.
    ulonglong2 v[ITEMS / 2];
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j) {
        const unsigned u = j * DR_THREADS + threadIdx.x;        // 16-byte unit inside the tile
        v[j] = __ldcs(reinterpret_cast<const ulonglong2*>(dense + tile_base + 2 * u));
    }
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j) {
        const unsigned u = j * DR_THREADS + threadIdx.x;
        if (v[j].x | v[j].y)
            *reinterpret_cast<ulonglong2*>(dense + tile_base + 2 * u) = make_ulonglong2(0ull, 0ull);
        *reinterpret_cast<ulonglong2*>(s_rows + (u / (ITEMS / 2)) * ROW + (u % (ITEMS / 2)) * 16) = v[j];
    }
    __syncthreads();
    const unsigned char* row = s_rows + threadIdx.x * ROW;
    const unsigned base = tile_base + threadIdx.x * ITEMS;
    int tf = 0, td = 0;
    unsigned mf = 0, md = 0;
#pragma unroll
    for (int k = 0; k < ITEMS / 2; ++k) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + k * 16);
        int df, dd;
        dense_unpack(v.x, df, dd);
        tf += df; td += dd;
        if (df != 0) mf |= 1u << (2 * k);
        if (dd != 0) md |= 1u << (2 * k);
        dense_unpack(v.y, df, dd);
        tf += df; td += dd;
        if (df != 0) mf |= 2u << (2 * k);
        if (dd != 0) md |= 2u << (2 * k);
    }
    if (base == 0) { mf |= 1u; md |= 1u; }
    long long tile_tot, tile_cnt;
    const long long ex_tot = block_excl_sum<long long>(((long long)tf << 32) + (long long)td, s_scan64, tile_tot);
    const long long ex_cnt = block_excl_sum<long long>(((long long)__popc(mf) << 31) | (long long)__popc(md),
                                                       s_scan64, tile_cnt);
    if (threadIdx.x < 32) {
        unsigned long long ex = lookback_excl_warp(st_tot, tile, (unsigned long long)tile_tot & LB_MASK);
        if (threadIdx.x == 0) s_base[0] = ex;
    } else if (threadIdx.x < 64) {
        unsigned long long b = lookback_excl_warp(st_cnt, tile, (unsigned long long)tile_cnt);
        if (threadIdx.x == 32) {
            s_base[1] = b;
            if (tile == gridDim.x - 1) {
                const unsigned long long tot = b + (unsigned long long)tile_cnt;
                out_counts[0] = tot >> 31;
                out_counts[1] = tot & 0x7fffffffull;
            }
        }
    }
    __syncthreads();
    if ((mf | md) == 0) return;
    int run_f, run_d;
    {
        const long long v = ((long long)(s_base[0] << 2) >> 2) + ex_tot;
        run_d = (int)(unsigned)(v & 0xffffffffll);
        run_f = (int)((v - (long long)run_d) >> 32);
    }
    const unsigned long long w_pk = s_base[1] + (unsigned long long)ex_cnt;
    unsigned long long w0 = w_pk >> 31, w1 = w_pk & 0x7fffffffull;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        int df, dd;
        dense_unpack(*reinterpret_cast<const unsigned long long*>(row + k * 8), df, dd);
        run_f += df;
        run_d += dd;
        if ((mf >> k) & 1u) {
            if (w0 < cap) { frag_idx[w0] = base + k; frag_val[w0] = run_f; }
            ++w0;
        }
        if ((md >> k) & 1u) {
            if (w1 < cap) { dir_idx[w1] = base + k; dir_val[w1] = run_d; }
            ++w1;
        }
    }
}


// ---- the same scan, leaner: entries staged through shared memory, zero words skipped ------
//
// The kernel above executes ~1450 instructions per warp and tile (ncu: 35 % of the issue
// slots, nothing else above 41 %: it is bound by how much it executes between its waits).
// Most of them sit in the emission pass: 16 unrolled steps x 2 tracks, each a divergent
// region (a warp runs the body whenever ANY lane has an entry there, i.e. always) with a
// capacity check, 64-bit addressing and two stores at ~15 % active lanes.  Here
//   * a thread only visits its NON-ZERO words in the emission pass (a zero word changes
//     neither running sum and emits nothing): ~3 of 16 on the chr22-scale workload, the
//     warp-wide maximum ~8 -- a find-first-set loop instead of 16 unrolled steps;
//   * entries go to a staging area in shared memory at the block-wide rank the count scan
//     already provides (16-bit position in the tile + 32-bit value), branch-free (a lane
//     without an entry writes a dump slot), and the block copies them out with coalesced
//     stores: ~300 full sectors per tile instead of ~2500 partial ones;
//   * both block scans (running sums, entry counts) share their barriers;
//   * the tile total is the sum of the packed words, unpacked once.
// A tile with more than DS_CAP entries in a track (> 31 % of its positions) takes the
// direct path: same loop, stores straight to global memory with the capacity check.
constexpr int DS_CAP = 1280;
constexpr int DS_SLOTS = DS_CAP + 8;                   // + dump slot, keeps the arrays 16-byte aligned

// two exclusive block sums at once (same barriers); s needs 2 x 33 entries
__device__ __forceinline__ void block_excl_sum2(long long a, long long b, long long* s,
                                                long long& ex_a, long long& ex_b,
                                                long long& tot_a, long long& tot_b) {
    long long ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long oa = __shfl_up_sync(FULL, ia, d), ob = __shfl_up_sync(FULL, ib, d);
        if (lane_id() >= (unsigned)d) { ia += oa; ib += ob; }
    }
    const unsigned w = warp_id(), l = lane_id();
    if (l == 31) { s[w] = ia; s[33 + w] = ib; }
    __syncthreads();
    if (w == 0) {
        const unsigned nw = (blockDim.x + 31) >> 5;
        const long long xa = l < nw ? s[l] : 0ll, xb = l < nw ? s[33 + l] : 0ll;
        long long ja = xa, jb = xb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long oa = __shfl_up_sync(FULL, ja, d), ob = __shfl_up_sync(FULL, jb, d);
            if (l >= (unsigned)d) { ja += oa; jb += ob; }
        }
        s[l] = ja - xa;
        s[33 + l] = jb - xb;
        if (l == 31) { s[32] = ja; s[65] = jb; }
    }
    __syncthreads();
    ex_a = s[w] + ia - a;
    ex_b = s[33 + w] + ib - b;
    tot_a = s[32];
    tot_b = s[65];
}

// ---- reduce, then scan: no chained scan at all ------------------------------------------
//
// With its parts switched off one at a time (GPC_DENSE_DBG, chr22-scale array, 13 100 tiles)
// the single-pass kernel below takes 0.36 ms complete, 0.20 ms without the look-back, 0.19 ms
// for loads and zeroes alone and 0.08 ms for the loads alone (5.5 TB/s): the chained scan --
// every tile waiting for the aggregates of the tiles before it, ~600 of them in flight -- is
// 45 % of it (ncu: 38 % of all stall samples on the barrier behind the look-back), however
// many predecessors a round trip inspects.  Reading the array twice is cheaper than that:
//   dense_tile_sums   one streaming read (no stores, no barrier before the loads are all in
//                     flight): packed sum and entry counts of every tile, and of every
//                     group of 64 tiles (one atomic per tile)
//   dense_runs        the kernel below; a tile adds up the groups and the tiles before it
//                     (<= 270 independent loads spread over the block): tiles are independent
// 0.08 + 0.22 ms.
__global__ void __launch_bounds__(DR_THREADS)
dense_tile_sums_kernel(const unsigned long long* __restrict__ dense,
                       unsigned long long* __restrict__ sums /* [2][tiles] */,
                       unsigned long long* __restrict__ supers /* [2][groups], zeroed */, unsigned groups) {
    constexpr int ITEMS = 16;
    constexpr int TILE = DR_THREADS * ITEMS;
    __shared__ unsigned long long s_red[2][DR_THREADS / 32];
    const unsigned tile = blockIdx.x, tile_base = tile * TILE;
    ulonglong2 v[ITEMS / 2];
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j)
        v[j] = *reinterpret_cast<const ulonglong2*>(dense + tile_base + 2 * (j * DR_THREADS + threadIdx.x));
    unsigned long long wsum = 0ull;
    unsigned cf = 0, cd = 0;
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j) {
        wsum += v[j].x + v[j].y;
        const unsigned lx = (unsigned)v[j].x, ly = (unsigned)v[j].y;
        cd += (lx != 0u) + (ly != 0u);
        cf += ((unsigned)(v[j].x >> 32) + (lx >> 31) != 0u) + ((unsigned)(v[j].y >> 32) + (ly >> 31) != 0u);
    }
    if (tile == 0 && threadIdx.x == 0) {            // position 0 always carries an entry
        const unsigned lx = (unsigned)v[0].x;
        cd += (lx == 0u);
        cf += ((unsigned)(v[0].x >> 32) + (lx >> 31) == 0u);
    }
    unsigned long long cnt = ((unsigned long long)cf << 31) | cd;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        wsum += __shfl_xor_sync(FULL, wsum, d);
        cnt += __shfl_xor_sync(FULL, cnt, d);
    }
    if (lane_id() == 0) { s_red[0][warp_id()] = wsum; s_red[1][warp_id()] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, b = 0;
#pragma unroll
        for (int w = 0; w < DR_THREADS / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
        // two-level totals (scan.cuh): the emission kernel's tiles look their prefix up
        tile_total_publish(sums, supers, tile, a);
        tile_total_publish(sums + gridDim.x, supers + groups, tile, b);
    }
}

__global__ void __launch_bounds__(DR_THREADS, 4)
dense_runs_staged_kernel(unsigned long long* __restrict__ dense, unsigned n_pos,
                         uint32_t* __restrict__ frag_idx, int32_t* __restrict__ frag_val,
                         uint32_t* __restrict__ dir_idx, int32_t* __restrict__ dir_val,
                         unsigned long long cap,
                         unsigned long long* __restrict__ st_tot /* zeroed */,
                         unsigned long long* __restrict__ st_cnt /* zeroed */,
                         unsigned* __restrict__ ticket, unsigned long long* __restrict__ out_counts,
                         const unsigned long long* __restrict__ tile_base_in /* [2][tiles] tile totals of
                             dense_tile_sums_kernel, or NULL: chained scan in here */,
                         const unsigned long long* __restrict__ tile_supers /* [2][groups] */, unsigned groups,
                         int dbg /* experiments only (GPC_DENSE_DBG): results are wrong when set */) {
    constexpr int ITEMS = 16;
    constexpr int TILE = DR_THREADS * ITEMS;
    constexpr int ROW = ITEMS * 8 + 16;
    extern __shared__ __align__(16) unsigned char s_rows[];      // DR_THREADS * ROW, then the staging area
    int32_t* const s_val = reinterpret_cast<int32_t*>(s_rows + DR_THREADS * ROW);     // [2][DS_SLOTS]
    uint16_t* const s_pos = reinterpret_cast<uint16_t*>(s_val + 2 * DS_SLOTS);        // [2][DS_SLOTS]
    __shared__ long long s_scan64[66];
    __shared__ unsigned s_slot;
    __shared__ unsigned long long s_base[2];
    const unsigned tile = (tile_base_in || (dbg & 4)) ? blockIdx.x : claim_tile(ticket, &s_slot);
    const unsigned tile_base = tile * TILE;
    // all loads of the tile first, then the zeroes: written as load / test / store per unit
    // the compiler kept that order (the stores go to the array the next load reads), i.e.
    // eight dependent DRAM round trips per tile -- 40 % of the kernel's stall samples.  The
    // array is padded to whole tiles (dense_acquire), so every unit may be loaded in full.
    ulonglong2 v[ITEMS / 2];
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j) {
        const unsigned u = j * DR_THREADS + threadIdx.x;        // 16-byte unit inside the tile
        v[j] = __ldcs(reinterpret_cast<const ulonglong2*>(dense + tile_base + 2 * u));
    }
#pragma unroll
    for (int j = 0; j < ITEMS / 2; ++j) {
        const unsigned u = j * DR_THREADS + threadIdx.x;
        if ((v[j].x | v[j].y) && !(dbg & 8))
            *reinterpret_cast<ulonglong2*>(dense + tile_base + 2 * u) = make_ulonglong2(0ull, 0ull);
        *reinterpret_cast<ulonglong2*>(s_rows + (u / (ITEMS / 2)) * ROW + (u % (ITEMS / 2)) * 16) = v[j];
    }
    __syncthreads();
    if (dbg & 16) return;                                        // load + zero only
    const unsigned char* row = s_rows + threadIdx.x * ROW;
    const unsigned local = threadIdx.x * ITEMS;                  // first position of this thread in the tile
    // first pass: packed sum of the 16 words, one bit per word and track with a non-zero step
    // (word = (dfrag << 32) + ddir, signed: dfrag = high word + sign bit of the low word)
    unsigned long long wsum = 0ull;
    unsigned mf = 0, md = 0;
#pragma unroll
    for (int k = 0; k < ITEMS / 2; ++k) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + k * 16);
        wsum += v.x + v.y;
        const unsigned lx = (unsigned)v.x, ly = (unsigned)v.y;
        if (lx != 0u) md |= 1u << (2 * k);
        if ((unsigned)(v.x >> 32) + (lx >> 31) != 0u) mf |= 1u << (2 * k);
        if (ly != 0u) md |= 2u << (2 * k);
        if ((unsigned)(v.y >> 32) + (ly >> 31) != 0u) mf |= 2u << (2 * k);
    }
    if (tile_base + local == 0) { mf |= 1u; md |= 1u; }
    long long ex_tot, ex_cnt, tile_tot, tile_cnt;
    block_excl_sum2((long long)wsum, ((long long)__popc(mf) << 31) | (long long)__popc(md), s_scan64,
                    ex_tot, ex_cnt, tile_tot, tile_cnt);
    if (tile_base_in) {                                          // totals of all tiles are known
        __shared__ unsigned long long s_red[33];
        const unsigned long long b0 = tile_prefix_lookup(tile_base_in, tile_supers, tile, s_red);
        const unsigned long long b1 = tile_prefix_lookup(tile_base_in + gridDim.x, tile_supers + groups, tile, s_red);
        if (threadIdx.x == 0) {
            s_base[0] = b0;
            s_base[1] = b1;
            if (tile == gridDim.x - 1) {
                const unsigned long long tot = b1 + (unsigned long long)tile_cnt;
                out_counts[0] = tot >> 31;
                out_counts[1] = tot & 0x7fffffffull;
            }
        }
    } else if (dbg & 1) {                                        // no look-back: every tile starts at zero
        if (threadIdx.x == 0) { s_base[0] = 0; s_base[1] = (unsigned long long)tile * (2ull << 31 | 2ull) * 600; }
    } else if (threadIdx.x < 32) {
        unsigned long long ex = lookback_excl_warp(st_tot, tile, (unsigned long long)tile_tot & LB_MASK);
        if (threadIdx.x == 0) s_base[0] = ex;
    } else if (threadIdx.x < 64) {
        unsigned long long b = lookback_excl_warp(st_cnt, tile, (unsigned long long)tile_cnt);
        if (threadIdx.x == 32) {
            s_base[1] = b;
            if (tile == gridDim.x - 1) {
                const unsigned long long tot = b + (unsigned long long)tile_cnt;
                out_counts[0] = tot >> 31;
                out_counts[1] = tot & 0x7fffffffull;
            }
        }
    }
    __syncthreads();
    const unsigned long long g0 = s_base[1] >> 31, g1 = s_base[1] & 0x7fffffffull;   // tile's first entries
    const unsigned n0 = (unsigned)((unsigned long long)tile_cnt >> 31);
    const unsigned n1 = (unsigned)((unsigned long long)tile_cnt & 0x7fffffffull);
    const bool staged = n0 <= (unsigned)DS_CAP && n1 <= (unsigned)DS_CAP;            // (uniform over the block)
    int run_f, run_d;
    {
        const long long v = ((long long)(s_base[0] << 2) >> 2) + ex_tot;
        run_d = (int)(unsigned)(v & 0xffffffffll);
        run_f = (int)((v - (long long)run_d) >> 32);
    }
    unsigned r0 = (unsigned)((unsigned long long)ex_cnt >> 31);                      // rank inside the tile
    unsigned r1 = (unsigned)((unsigned long long)ex_cnt & 0x7fffffffull);
    unsigned m = mf | md;
    if (dbg & 2) return;                                         // no emission pass
    if (staged) {
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            int df, dd;
            dense_unpack(*reinterpret_cast<const unsigned long long*>(row + k * 8), df, dd);
            run_f += df;
            run_d += dd;
            const unsigned ef = (mf >> k) & 1u, ed = (md >> k) & 1u;
            const unsigned i0 = ef ? r0 : (unsigned)DS_CAP;                // dump slot
            const unsigned i1 = DS_SLOTS + (ed ? r1 : (unsigned)DS_CAP);
            s_pos[i0] = (uint16_t)(local + k);
            s_val[i0] = run_f;
            s_pos[i1] = (uint16_t)(local + k);
            s_val[i1] = run_d;
            r0 += ef;
            r1 += ed;
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < n0; i += DR_THREADS)
            if (g0 + i < cap) {
                frag_idx[g0 + i] = tile_base + s_pos[i];
                frag_val[g0 + i] = s_val[i];
            }
        for (unsigned i = threadIdx.x; i < n1; i += DR_THREADS)
            if (g1 + i < cap) {
                dir_idx[g1 + i] = tile_base + s_pos[DS_SLOTS + i];
                dir_val[g1 + i] = s_val[DS_SLOTS + i];
            }
    } else {
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            int df, dd;
            dense_unpack(*reinterpret_cast<const unsigned long long*>(row + k * 8), df, dd);
            run_f += df;
            run_d += dd;
            if ((mf >> k) & 1u) {
                if (g0 + r0 < cap) { frag_idx[g0 + r0] = tile_base + local + k; frag_val[g0 + r0] = run_f; }
                ++r0;
            }
            if ((md >> k) & 1u) {
                if (g1 + r1 < cap) { dir_idx[g1 + r1] = tile_base + local + k; dir_val[g1 + r1] = run_d; }
                ++r1;
            }
        }
    }
}
constexpr int DS_SMEM = DR_THREADS * (16 * 8 + 16) + 2 * DS_SLOTS * 6;


// ---- the same scan with the tile loads on the bulk-copy engine -------------------------
//
// Persistent blocks, DP_STAGES tile buffers each.  ONE cp.async.bulk (1-D TMA, SASS UBLKCP)
// per tile moves its 32 KB from the difference array into shared memory and completes on an
// mbarrier; a block keeps DP_STAGES - 1 tiles in flight while it scans the oldest one, so the
// load latency that the one-tile-per-block kernel waits out (long_scoreboard 13, barrier 11
// warps per issue) is hidden behind the scan, the look-back and the writes of the tile before.
// (First attempt: one 128-byte bulk copy per thread into padded rows -- 256 TMA operations per
// tile, ~170 ns each per SM: 3.75 ms for the pass.  The copy engine wants few, large copies.)
// Rows are NOT padded here (a bulk copy is contiguous): the order-free first pass reads the
// eight 16-byte units of a thread's row rotated by the lane, which is conflict-free; the
// emission pass needs them in order and pays the 8-way conflict.  Zeroes go back with plain
// coalesced stores, only where a word was non-zero.
constexpr int DP_STAGES = 3;
constexpr int DP_ITEMS = 16;
constexpr int DP_TILE = DR_THREADS * DP_ITEMS;            // 4096 positions
constexpr int DP_STAGE_BYTES = DP_TILE * 8;               // 32 KB
constexpr int DP_SMEM = DP_STAGES * DP_STAGE_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes,
                                          unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Tiles are dealt round-robin: block b scans tiles b, b + grid, b + 2 grid, ...  All blocks
// of a wave then work on neighbouring tiles at the same time, so the look-back of a tile
// finds its predecessors' aggregates published (the first version claimed tiles by ticket,
// three per block up front: the second and third tile of a block sat in its queue while
// the next block's first tile waited for their aggregates -- a serial chain through all
// blocks, 2.6 ms).  A tile only ever waits on tiles of earlier waves or of the same wave in
// blocks that are running: the launch is cooperative, i.e. refused unless the whole grid is
// resident.
__global__ void __launch_bounds__(DR_THREADS, 2)
dense_runs_pipelined_kernel(unsigned long long* __restrict__ dense, unsigned n_tiles,
                            uint32_t* __restrict__ frag_idx, int32_t* __restrict__ frag_val,
                            uint32_t* __restrict__ dir_idx, int32_t* __restrict__ dir_val,
                            unsigned long long cap,
                            unsigned long long* __restrict__ st_tot /* zeroed */,
                            unsigned long long* __restrict__ st_cnt /* zeroed */,
                            unsigned long long* __restrict__ out_counts) {
    extern __shared__ __align__(128) unsigned char dp_smem[];
    __shared__ __align__(8) unsigned long long s_bar[DP_STAGES];
    __shared__ long long s_scan64[33];
    __shared__ unsigned long long s_base[2];
    const unsigned t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < DP_STAGES; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // prologue: every stage gets its first tile
        for (int s = 0; s < DP_STAGES; ++s) {
            const unsigned tile = blockIdx.x + s * gridDim.x;
            if (tile < n_tiles) {
                mbar_arrive_expect_tx(&s_bar[s], DP_STAGE_BYTES);
                bulk_load(dp_smem + s * DP_STAGE_BYTES, dense + (size_t)tile * DP_TILE, DP_STAGE_BYTES,
                          &s_bar[s]);
            }
        }
    }
    __syncthreads();
    for (unsigned it = 0;; ++it) {
        const int s = it % DP_STAGES;
        const unsigned parity = (it / DP_STAGES) & 1u;
        const unsigned tile = blockIdx.x + it * gridDim.x;
        if (tile >= n_tiles) break;
        mbar_wait(&s_bar[s], parity);
        const unsigned char* stage = dp_smem + s * DP_STAGE_BYTES;
        // zero behind the read (striped: a warp's stores are contiguous), only non-zero words
        {
            unsigned long long* g = dense + (size_t)tile * DP_TILE;
#pragma unroll
            for (int j = 0; j < DP_ITEMS / 2; ++j) {
                const unsigned u = j * DR_THREADS + t;
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(stage + u * 16);
                if (v.x | v.y) *reinterpret_cast<ulonglong2*>(g + 2 * u) = make_ulonglong2(0ull, 0ull);
            }
        }
        const unsigned char* row = stage + t * (DP_ITEMS * 8);
        const unsigned base = tile * DP_TILE + t * DP_ITEMS;
        int tf = 0, td = 0;
        unsigned mf = 0, md = 0;
#pragma unroll
        for (int kk = 0; kk < DP_ITEMS / 2; ++kk) {
            const int k = (kk + (int)t) & 7;              // rotated by the lane: conflict-free
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + k * 16);
            int df, dd;
            dense_unpack(v.x, df, dd);
            tf += df; td += dd;
            if (df != 0) mf |= 1u << (2 * k);
            if (dd != 0) md |= 1u << (2 * k);
            dense_unpack(v.y, df, dd);
            tf += df; td += dd;
            if (df != 0) mf |= 2u << (2 * k);
            if (dd != 0) md |= 2u << (2 * k);
        }
        if (base == 0) { mf |= 1u; md |= 1u; }
        long long tile_tot, tile_cnt;
        const long long ex_tot = block_excl_sum<long long>(((long long)tf << 32) + (long long)td, s_scan64, tile_tot);
        const long long ex_cnt = block_excl_sum<long long>(((long long)__popc(mf) << 31) | (long long)__popc(md),
                                                           s_scan64, tile_cnt);
        if (t < 32) {
            unsigned long long ex = lookback_excl_warp(st_tot, tile, (unsigned long long)tile_tot & LB_MASK);
            if (t == 0) s_base[0] = ex;
        } else if (t < 64) {
            unsigned long long b = lookback_excl_warp(st_cnt, tile, (unsigned long long)tile_cnt);
            if (t == 32) {
                s_base[1] = b;
                if (tile == n_tiles - 1) {
                    const unsigned long long tot = b + (unsigned long long)tile_cnt;
                    out_counts[0] = tot >> 31;
                    out_counts[1] = tot & 0x7fffffffull;
                }
            }
        }
        __syncthreads();
        if (mf | md) {
            int run_f, run_d;
            {
                const long long v = ((long long)(s_base[0] << 2) >> 2) + ex_tot;
                run_d = (int)(unsigned)(v & 0xffffffffll);
                run_f = (int)((v - (long long)run_d) >> 32);
            }
            const unsigned long long w_pk = s_base[1] + (unsigned long long)ex_cnt;
            unsigned long long w0 = w_pk >> 31, w1 = w_pk & 0x7fffffffull;
#pragma unroll
            for (int k = 0; k < DP_ITEMS; ++k) {
                int df, dd;
                dense_unpack(*reinterpret_cast<const unsigned long long*>(row + k * 8), df, dd);
                run_f += df;
                run_d += dd;
                if ((mf >> k) & 1u) {
                    if (w0 < cap) { frag_idx[w0] = base + k; frag_val[w0] = run_f; }
                    ++w0;
                }
                if ((md >> k) & 1u) {
                    if (w1 < cap) { dir_idx[w1] = base + k; dir_val[w1] = run_d; }
                    ++w1;
                }
            }
        }
        // the stage is free once every thread has read its row for the last time
        __syncthreads();
        if (t == 0) {
            const unsigned next = tile + DP_STAGES * gridDim.x;
            if (next < n_tiles) {
                mbar_arrive_expect_tx(&s_bar[s], DP_STAGE_BYTES);
                bulk_load(dp_smem + s * DP_STAGE_BYTES, dense + (size_t)next * DP_TILE, DP_STAGE_BYTES, &s_bar[s]);
            }
        }
    }
}

// ---- duplicate filter (hash set keyed by start position) --------------------

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

constexpr unsigned DEDUP_EMPTY = 0xffffffffu;

// Pass 1: every slot of the table holds one read index -- the smallest index
// seen so far among the reads that share its key (start node, start offset);
// "first read wins", intervals.py:26-33.  The key of a slot is the key of the
// read it names, compared by looking that read up, so the table is 4 bytes per
// slot (30 MB for 5 M reads, L2-resident next to the two key arrays) and an
// insertion is one CAS in the common case.
__global__ void dedup_insert_kernel(const int32_t* __restrict__ start_node,
                                    const int32_t* __restrict__ start_off, unsigned n,
                                    unsigned* __restrict__ table, unsigned cap,
                                    unsigned* __restrict__ slot_of) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t node = start_node[i], off = start_off[i];
    unsigned long long key = ((unsigned long long)(uint32_t)node << 32) | (uint32_t)off;
    // any capacity (not only powers of two): multiply-shift range reduction
    unsigned h = __umulhi((unsigned)(mix64(key) >> 32), cap);
    while (true) {
        unsigned cur = table[h];
        if (cur == DEDUP_EMPTY) {
            cur = atomicCAS(&table[h], DEDUP_EMPTY, i);
            if (cur == DEDUP_EMPTY) { slot_of[i] = h; return; }
        }
        // occupied: the slot's key is the key of read `cur` (all reads that ever
        // own this slot share it)
        if (start_node[cur] == node && start_off[cur] == off) {
            atomicMin(&table[h], i);
            slot_of[i] = h;
            return;
        }
        if (++h == cap) h = 0;
    }
}

// Pass 2 (compaction): the reads that own their slot, in input order.
struct DedupOp {
    const unsigned* table_first;
    const unsigned* slot_of;
    uint32_t* out;
    __device__ __forceinline__ bool keep(unsigned i) const { return table_first[slot_of[i]] == i; }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const { out[w] = i; }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

// Reads arrive in arbitrary order; walking them in (coarse) graph order makes
// neighbouring threads gather the same node records, so the walk is served by
// L1/L2 instead of DRAM (814 MB of DRAM traffic per launch for 197 MB of
// compulsory bytes were measured without it -- profiles/traffic.json, r01).
__global__ void walk_order_keys_kernel(const int32_t* __restrict__ start_node,
                                       const uint32_t* __restrict__ read_index, unsigned n,
                                       int min_node, uint32_t n_nodes, int shift,
                                       uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned r = read_index ? read_index[i] : i;
    const int sn = start_node[r];
    long long v = (long long)(sn < 0 ? -(long long)sn : (long long)sn) - min_node;
    v = v < 0 ? 0 : (v >= (long long)n_nodes ? (long long)n_nodes - 1 : v);   // validated by the walk
    keys[i] = (uint32_t)v >> shift;
    vals[i] = r;
}

// gpc_reads_t.hdr: one 16-byte record per read (see include/gpc_b200.h)
__global__ void read_headers_kernel(gpc_reads_t reads, unsigned n, uint4* __restrict__ hdr) {
    unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long p0 = reads.path_ptr32 ? (long long)reads.path_ptr32[r] : reads.path_ptr[r];
    const long long p1 = reads.path_ptr32 ? (long long)reads.path_ptr32[r + 1] : reads.path_ptr[r + 1];
    const long long k = p1 - p0;
    bool bad = k <= 0 || k >= (1ll << 31) || p0 < 0 || p0 >= (1ll << 32);
    if (!bad)
        bad = reads.path[p0] != reads.start_node[r] || reads.path[p1 - 1] != reads.end_node[r];
    hdr[r] = make_uint4((unsigned)p0, (unsigned)(bad ? 0 : k) | (bad ? 0x80000000u : 0u),
                        (unsigned)reads.start_off[r], (unsigned)reads.end_off[r]);
}

// Compact reads (what crosses the host link) -> the arrays the kernels read.
//   offs[r]     start offset | end offset << 16
//   path_len[r] nodes on the read's path (1..255)
//   path_off[r] exclusive prefix sum of path_len (computed on the device)
__global__ void widen_reads_kernel(const uint32_t* __restrict__ offs, const uint8_t* __restrict__ path_len,
                                   const uint32_t* __restrict__ path_off, const int32_t* __restrict__ path,
                                   unsigned n, int32_t* __restrict__ start_node,
                                   int32_t* __restrict__ start_off, uint4* __restrict__ hdr) {
    unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t o = offs[r], k = path_len[r], p0 = path_off[r];
    start_node[r] = k ? path[p0] : 0;
    start_off[r] = (int32_t)(o & 0xffffu);
    hdr[r] = make_uint4(p0, k ? k : 0x80000000u, o & 0xffffu, o >> 16);
}

// The same with the paths delta-coded (10.3 bytes per 36-bp read on the link instead of
// 14.3): the first node id of every path as int32, every further node as the int8 step from
// the node before it (graphs with sorted ids: neighbours on a path are a few ids apart).
// Read r owns path_len[r] - 1 steps; the steps of all reads stand back to back, so those of
// read r begin at path_off[r] - r.  Also writes the int32 path array the walk reads.
__global__ void widen_reads_delta_kernel(const uint32_t* __restrict__ offs, const uint8_t* __restrict__ path_len,
                                         const uint32_t* __restrict__ path_off,
                                         const int32_t* __restrict__ first_node,
                                         const int8_t* __restrict__ step, unsigned n,
                                         int32_t* __restrict__ path, int32_t* __restrict__ start_node,
                                         int32_t* __restrict__ start_off, uint4* __restrict__ hdr) {
    unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t o = offs[r], k = path_len[r], p0 = path_off[r];
    int32_t v = first_node[r];
    start_node[r] = k ? v : 0;
    start_off[r] = (int32_t)(o & 0xffffu);
    hdr[r] = make_uint4(p0, k ? k : 0x80000000u, o & 0xffffu, o >> 16);
    if (k == 0) return;
    path[p0] = v;
    const int8_t* s = step + (p0 - r);
    for (uint32_t j = 1; j < k; ++j) {
        v += (int32_t)s[j - 1];
        path[p0 + j] = v;
    }
}

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_unique_reads(const int32_t* start_node, const int32_t* start_off,
                                uint64_t n_reads, uint32_t* out_index, uint64_t* out_count,
                                void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *out_count = 0;
    if (n_reads == 0) return GPC_OK;
    if (n_reads >= (1ull << 31)) { set_error("too many reads"); return GPC_ERR_ARG; }
    unsigned n = (unsigned)n_reads;
    // load factor 2/3, 4 bytes per slot: the table of a 5 M read batch (30 MB) stays
    // inside the 126 MB L2 instead of turning every probe into a DRAM access
    unsigned cap = (unsigned)max(1024ull, (3ull * n + 1) / 2);
    Scratch first, slot, total;
    GPC_CUDA(first.alloc((size_t)cap * 4, stream));
    GPC_CUDA(slot.alloc((size_t)n * 4, stream));
    GPC_CUDA(total.alloc(8, stream));
    GPC_CUDA(cudaMemsetAsync(first.p, 0xff, (size_t)cap * 4, stream));
    unsigned blocks = div_up(n, 256);
    GPC_PROF_BYTES("dedup_insert", 12.0 * n);
    dedup_insert_kernel<<<blocks, 256, 0, stream>>>(start_node, start_off, n,
                                                   first.as<unsigned>(), cap,
                                                   slot.as<unsigned>());
    GPC_LAUNCH_CHECK();
    GPC_TRY(select_if("dedup_select", DedupOp{first.as<unsigned>(), slot.as<unsigned>(), out_index}, n,
                      total.as<unsigned long long>(), stream, nullptr, /*keep_is_cheap=*/false));
    unsigned long long h = 0;
    GPC_READ_BACK({total.p, &h, (unsigned)(8)});
    *out_count = h;
    return GPC_OK;
}

// Walk order: read indices sorted by start node / 2^shift (16-bit keys, two passes), so
// that neighbouring threads walk neighbouring nodes.  The sorted index list lives in `keep`.
struct WalkOrder { Scratch k0, k1, v0, v1; };
static int walk_order(const gpc_graph_t* graph, const gpc_reads_t* reads, const uint32_t** read_index,
                      uint64_t n_reads, WalkOrder& keep, cudaStream_t stream) {
    if (!(n_reads >= (1u << 16) && n_reads < (1ull << 30)) || getenv("GPC_NO_WALK_ORDER")) return GPC_OK;
    int bits = 1;
    while ((1ull << bits) < graph->n_nodes) ++bits;
    const int shift = bits > 16 ? bits - 16 : 0;
    GPC_CUDA(keep.k0.alloc(n_reads * 4, stream));
    GPC_CUDA(keep.k1.alloc(n_reads * 4, stream));
    GPC_CUDA(keep.v0.alloc(n_reads * 4, stream));
    GPC_CUDA(keep.v1.alloc(n_reads * 4, stream));
    GPC_PROF_BYTES("walk_order_keys", 16.0 * n_reads);
    walk_order_keys_kernel<<<div_up(n_reads, 256), 256, 0, stream>>>(
        reads->start_node, *read_index, (unsigned)n_reads, graph->min_node, graph->n_nodes,
        shift, keep.k0.as<uint32_t>(), keep.v0.as<uint32_t>());
    GPC_LAUNCH_CHECK();
    bool in_tmp = false;
    GPC_TRY((radix_sort<uint32_t, true>(keep.k0.as<uint32_t>(), keep.k1.as<uint32_t>(),
                                        keep.v0.as<uint32_t>(), keep.v1.as<uint32_t>(), n_reads, 0,
                                        bits - shift, false, &in_tmp, stream)));
    *read_index = in_tmp ? keep.v1.as<uint32_t>() : keep.v0.as<uint32_t>();
    return GPC_OK;
}

// The pool behind Frontier::grow: 256 chunks (8 MB), enough for 256 reads of one launch to
// hold up to 4096 pending nodes each (GPC_SPILL_CHUNKS overrides; 0 = no pool).
struct SpillStore {
    Scratch mem;
    SpillPool pool{nullptr, nullptr, 0};
    int init(cudaStream_t stream) {
        const char* env = getenv("GPC_SPILL_CHUNKS");
        const int chunks = env ? atoi(env) : 256;
        if (chunks <= 0) return GPC_OK;
        const size_t words = (size_t)chunks * 2 * SPILL_CHUNK;
        GPC_CUDA(mem.alloc(words * 4 + 16, stream));
        GPC_CUDA(cudaMemsetAsync(mem.p, 0, 16, stream));
        pool.next = mem.as<unsigned>();
        pool.base = mem.as<uint32_t>() + 4;
        pool.n_chunks = (unsigned)chunks;
        return GPC_OK;
    }
};

static int walk_error(int e) {
    if (e == GPC_ERR_INVALID_INTERVAL) {
        set_error("Interval has node(s) not part of graph/pileup");
        return e;
    }
    if (e == GPC_ERR_FRONTIER_OVERFLOW) {
        set_error("extension frontier: a read holds more than %d pending nodes, or more than the "
                  "spill pool's reads hold more than %d (GPC_SPILL_CHUNKS)", SPILL_CHUNK, 4 + FRONTIER_CAP);
        return e;
    }
    return GPC_OK;
}

static int check_sample_args(const gpc_graph_t* graph, int32_t extension) {
    if (extension <= 0) {
        set_error("Invalid extension size %d used. Must be positive. Is fragment length < read length?", extension);
        return GPC_ERR_ARG;
    }
    if (graph->size_bp >= (1u << 29)) {
        set_error("graph of %u bp exceeds the 2^29 limit of the packed event key", graph->size_bp);
        return GPC_ERR_ARG;
    }
    return GPC_OK;
}

// bytes the walk must move per read: its header (or the six SoA fields), its index in
// the walk order; the path nodes and graph records are counted per visit in DESIGN.md
static double walk_read_bytes(const gpc_reads_t* reads) { return reads->hdr ? 20.0 : 36.0; }

extern "C" int gpc_build_read_headers(const gpc_reads_t* reads, uint32_t* hdr, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (reads->n_reads == 0) return GPC_OK;
    if (reads->n_reads >= (1ull << 31)) { set_error("too many reads"); return GPC_ERR_ARG; }
    GPC_PROF_BYTES("read_headers", 48.0 * reads->n_reads);
    read_headers_kernel<<<div_up(reads->n_reads, 256), 256, 0, stream>>>(
        *reads, (unsigned)reads->n_reads, reinterpret_cast<uint4*>(hdr));
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}


extern "C" int gpc_reads_from_compact(const uint32_t* offs, const uint8_t* path_len,
                                      const int32_t* path, uint64_t n_reads, int32_t* start_node,
                                      int32_t* start_off, uint32_t* hdr, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads == 0) return GPC_OK;
    if (n_reads >= (1ull << 31)) { set_error("too many reads"); return GPC_ERR_ARG; }
    Scratch off;
    GPC_CUDA(off.alloc(n_reads * 4, stream));
    GPC_TRY(exclusive_scan<unsigned char>(path_len, off.as<uint32_t>(), n_reads, nullptr, stream));
    GPC_PROF_BYTES("widen_reads", 37.0 * n_reads);
    widen_reads_kernel<<<div_up(n_reads, 256), 256, 0, stream>>>(
        offs, path_len, off.as<uint32_t>(), path, (unsigned)n_reads, start_node, start_off,
        reinterpret_cast<uint4*>(hdr));
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

extern "C" int gpc_reads_from_compact_delta(const uint32_t* offs, const uint8_t* path_len,
                                            const int32_t* first_node, const int8_t* step,
                                            uint64_t n_reads, int32_t* path, int32_t* start_node,
                                            int32_t* start_off, uint32_t* hdr, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_reads == 0) return GPC_OK;
    if (n_reads >= (1ull << 31)) { set_error("too many reads"); return GPC_ERR_ARG; }
    Scratch off;
    GPC_CUDA(off.alloc(n_reads * 4, stream));
    GPC_TRY(exclusive_scan<unsigned char>(path_len, off.as<uint32_t>(), n_reads, nullptr, stream));
    GPC_PROF_BYTES("widen_reads", 42.0 * n_reads);
    widen_reads_delta_kernel<<<div_up(n_reads, 256), 256, 0, stream>>>(
        offs, path_len, off.as<uint32_t>(), first_node, step, (unsigned)n_reads, path, start_node,
        start_off, reinterpret_cast<uint4*>(hdr));
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

extern "C" int gpc_sample_events(const gpc_graph_t* graph, const gpc_reads_t* reads,
                                 const uint32_t* read_index, uint64_t n_reads, int32_t extension,
                                 uint32_t* events, uint64_t event_capacity,
                                 uint64_t* n_events, uint8_t* touched, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_events = 0;
    GPC_TRY(check_sample_args(graph, extension));
    GPC_CUDA(cudaMemsetAsync(touched, 0, graph->n_nodes, stream));
    if (n_reads == 0) return GPC_OK;
    Scratch ctr;
    GPC_CUDA(ctr.alloc(16, stream));
    GPC_CUDA(cudaMemsetAsync(ctr.p, 0, 16, stream));
    unsigned long long* count = ctr.as<unsigned long long>();
    int* err = reinterpret_cast<int*>(count + 1);
    WalkOrder order;
    GPC_TRY(walk_order(graph, reads, &read_index, n_reads, order, stream));
    // The fast kernel holds up to 4 + 32 pending nodes per read; a launch in which some read
    // needs more (deeply nested bubbles with long branches) is run again with the kernel whose
    // spill list can move into a pool chunk of 4096 entries.
    SpillStore spill;
    unsigned long long h[2] = {0, 0};
    for (int big = 0; big < 2; ++big) {
        if (big) {
            GPC_TRY(spill.init(stream));
            GPC_CUDA(cudaMemsetAsync(ctr.p, 0, 16, stream));
        }
        GPC_PROF("walk_reads");
        if (graph->flags & GPC_GRAPH_UNORDERED)
            walk_reads_cyclic_kernel<false><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, events, event_capacity, count,
                nullptr, touched, err);
        else if (big)
            walk_reads_kernel<false, true><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, events, event_capacity, count,
                nullptr, touched, err, spill.pool);
        else
            walk_reads_kernel<false, false><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, events, event_capacity, count,
                nullptr, touched, err, spill.pool);
        GPC_LAUNCH_CHECK();
        GPC_READ_BACK({ctr.p, h, (unsigned)(16)});
        if ((int)(h[1] & 0xffffffffu) != GPC_ERR_FRONTIER_OVERFLOW || (graph->flags & GPC_GRAPH_UNORDERED))
            break;
    }
    *n_events = h[0];
    prof_add_bytes("walk_reads", walk_read_bytes(reads) * (double)n_reads + 4.0 * (double)h[0]);
    GPC_TRY(walk_error((int)(h[1] & 0xffffffffu)));
    if (h[0] > event_capacity) return GPC_ERR_CAPACITY;   // caller retries with *n_events
    return GPC_OK;
}

// ---- the difference array of the dense path: one per device, kept between calls and
// left zeroed by dense_runs_kernel (a failed call marks it dirty) -------------------------
namespace gpc {
struct DenseWorkspace {
    unsigned long long* p = nullptr;
    size_t words = 0;
    bool busy = false, dirty = false;
    cudaEvent_t done = nullptr;
    cudaStream_t last_stream = nullptr;
};
static DenseWorkspace g_dense[64];
static std::mutex g_dense_mutex;

// Hands out `words` zeroed 64-bit words valid in `stream` order.  *ws == nullptr means the
// cached array is in use by another host thread: `fallback` then owns a pool allocation.
static int dense_acquire(size_t words, cudaStream_t stream, DenseWorkspace** ws, Scratch& fallback,
                         unsigned long long** out) {
    int dev = 0;
    GPC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_dense_mutex);
    DenseWorkspace& w = g_dense[dev & 63];
    if (w.busy) {
        *ws = nullptr;
        GPC_CUDA(fallback.alloc(words * 8, stream));
        GPC_CUDA(cudaMemsetAsync(fallback.p, 0, words * 8, stream));
        *out = fallback.as<unsigned long long>();
        return GPC_OK;
    }
    if (w.done && w.last_stream != stream) GPC_CUDA(cudaStreamWaitEvent(stream, w.done, 0));
    if (words > w.words) {
        if (w.p) { GPC_CUDA(cudaDeviceSynchronize()); GPC_CUDA(cudaFree(w.p)); w.p = nullptr; w.words = 0; }
        const size_t want = words + words / 8 + 4096;           // some slack: chromosomes differ in size
        GPC_CUDA(cudaMalloc((void**)&w.p, want * 8));
        w.words = want;
        w.dirty = true;
    }
    if (w.dirty) {
        GPC_CUDA(cudaMemsetAsync(w.p, 0, w.words * 8, stream));
        w.dirty = false;
    }
    if (!w.done) GPC_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    w.busy = true;
    *ws = &w;
    *out = w.p;
    return GPC_OK;
}
static void dense_release(DenseWorkspace* ws, bool clean, cudaStream_t stream) {
    if (!ws) return;
    std::lock_guard<std::mutex> lock(g_dense_mutex);
    ws->busy = false;
    if (!clean) ws->dirty = true;
    ws->last_stream = stream;
    cudaEventRecord(ws->done, stream);
}
}  // namespace gpc

static int events_to_runs_impl(uint32_t* events, uint32_t* events_tmp, uint64_t n_events,
                               uint32_t size_bp, uint32_t* frag_idx, int32_t* frag_val,
                               uint64_t* n_frag, uint32_t* direct_idx, int32_t* direct_val,
                               uint64_t* n_direct, uint64_t capacity, cudaStream_t stream);

extern "C" int gpc_sample_pileup(const gpc_graph_t* graph, const gpc_reads_t* reads,
                                 const uint32_t* read_index, uint64_t n_reads, int32_t extension,
                                 uint32_t* frag_idx, int32_t* frag_val, uint32_t* direct_idx,
                                 int32_t* direct_val, uint64_t capacity, uint64_t* n_frag,
                                 uint64_t* n_direct, uint64_t* n_events, uint8_t* touched,
                                 int32_t mode, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_frag = *n_direct = *n_events = 0;
    GPC_TRY(check_sample_args(graph, extension));
    if (capacity < 1) { *n_frag = *n_direct = 1; return GPC_ERR_CAPACITY; }
    if (mode == 0) {
        // a read deposits ~2 steps (more across bubbles); the array costs ~16 bytes per
        // position, the list ~40 bytes per step (three sort passes + the run kernel)
        if (const char* e = getenv("GPC_SAMPLE_PILEUP_MODE")) mode = atoi(e);
        if (mode != 1 && mode != 2)
            mode = (20.0 * (double)n_reads >= (double)graph->size_bp) ? 1 : 2;
    }
    if (mode == 2 || n_reads == 0) {
        // event list: capacity guess of six steps per read, grown once if the walk needs more
        GPC_CUDA(cudaMemsetAsync(touched, 0, graph->n_nodes, stream));
        uint64_t cap = n_reads * 6 + 1024, n_ev = 0;
        for (int attempt = 0;; ++attempt) {
            Scratch ev, tmp;
            GPC_CUDA(ev.alloc(cap * 4, stream));
            int st = gpc_sample_events(graph, reads, read_index, n_reads, extension, ev.as<uint32_t>(),
                                       cap, &n_ev, touched, stream_);
            if (st == GPC_ERR_CAPACITY && attempt == 0) { cap = n_ev + 1024; continue; }
            GPC_TRY(st);
            *n_events = n_ev;
            GPC_CUDA(tmp.alloc(cap * 4, stream));
            // (runs beyond `capacity` are counted, not written: too small -> GPC_ERR_CAPACITY)
            return events_to_runs_impl(ev.as<uint32_t>(), tmp.as<uint32_t>(), n_ev, graph->size_bp,
                                       frag_idx, frag_val, n_frag, direct_idx, direct_val, n_direct,
                                       capacity, stream);
        }
    }
    // ---- difference array ------------------------------------------------------------
    GPC_CUDA(cudaMemsetAsync(touched, 0, graph->n_nodes, stream));
    const unsigned n_pos = graph->size_bp + 1;           // a fragment may close at size_bp
    // scan variant: 0 = tile sums, their scan, then independent 4096-position tiles with the
    // entries staged through shared memory (default); 2 = the same tiles in ONE pass with a
    // chained scan; 1 = single pass, every thread writing its own entries (the first version);
    // 3 = persistent blocks, tiles moved by the bulk-copy engine, three in flight (measured
    // slower, kept as the reference point for it)
    static const int variant = getenv("GPC_DENSE_SCAN") ? atoi(getenv("GPC_DENSE_SCAN")) : 0;
    const unsigned tile_size = DR_THREADS * 16;
    const unsigned tiles = div_up(n_pos, tile_size);
    DenseWorkspace* ws = nullptr;
    Scratch fallback, sc;
    unsigned long long* dense = nullptr;
    // (whole 4096-position tiles: the bulk-copy variant of the scan loads full rows)
    GPC_TRY(dense_acquire(((size_t)n_pos + 4095) & ~(size_t)4095, stream, &ws, fallback, &dense));
    struct Release {           // every way out gives the array back; only a finished scan leaves it clean
        DenseWorkspace* ws; cudaStream_t s; bool clean;
        ~Release() { dense_release(ws, clean, s); }
    } rel{ws, stream, false};
    const unsigned groups = (tiles >> RTS_GROUP_LOG2) + 1;
    const size_t bytes = ((size_t)tiles + groups) * 16 + 64;
    GPC_CUDA(sc.alloc(bytes, stream));
    char* p = sc.as<char>();
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(p);            // 2 run counts
    unsigned long long* n_steps = counts + 2;
    int* err = reinterpret_cast<int*>(counts + 3);
    unsigned* ticket = reinterpret_cast<unsigned*>(p + 32);
    unsigned long long* st_tot = reinterpret_cast<unsigned long long*>(p + 64);
    unsigned long long* st_cnt = st_tot + tiles;
    unsigned long long* st_super = st_cnt + tiles;                                    // [2][groups]
    WalkOrder order;
    GPC_TRY(walk_order(graph, reads, &read_index, n_reads, order, stream));
    SpillStore spill;
    unsigned long long h[4] = {0, 0, 0, 0};
    // (second round: only when the fast kernel's frontier overflowed, see gpc_sample_events;
    //  the scan of the first round has left the array zeroed again)
    for (int big = 0; big < 2; ++big) {
        if (big) GPC_TRY(spill.init(stream));
        GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
        rel.clean = false;
        GPC_PROF("walk_reads");
        if (graph->flags & GPC_GRAPH_UNORDERED)
            walk_reads_cyclic_kernel<true><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, nullptr, 0, n_steps, dense,
                touched, err);
        else if (big)
            walk_reads_kernel<true, true><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, nullptr, 0, n_steps, dense,
                touched, err, spill.pool);
        else
            walk_reads_kernel<true, false><<<div_up(n_reads, WALK_THREADS), WALK_THREADS, 0, stream>>>(
                *graph, *reads, read_index, (unsigned)n_reads, extension, nullptr, 0, n_steps, dense,
                touched, err, spill.pool);
        GPC_LAUNCH_CHECK();
        if (variant == 3) {
            GPC_PROF_BYTES("dense_runs", 16.0 * n_pos);
            static bool configured3 = false;
            if (!configured3) {
                GPC_CUDA(cudaFuncSetAttribute(dense_runs_pipelined_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
                configured3 = true;
            }
            static int per_sm = 0;
            if (!per_sm)
                GPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_runs_pipelined_kernel,
                                                                       DR_THREADS, DP_SMEM));
            if (per_sm < 1) { set_error("dense_runs_pipelined does not fit an SM"); return GPC_ERR_INTERNAL; }
            unsigned blocks = min(tiles, (unsigned)(sm_count() * per_sm));
            unsigned n_tiles_arg = tiles;
            unsigned long long cap_arg = capacity;
            void* args[] = {&dense, &n_tiles_arg, &frag_idx, &frag_val, &direct_idx, &direct_val, &cap_arg,
                            &st_tot, &st_cnt, &counts};
            // cooperative: refused unless every block of the grid is resident (the look-back
            // of a tile may wait on any block of its wave)
            GPC_CUDA(cudaLaunchCooperativeKernel((void*)dense_runs_pipelined_kernel, dim3(blocks),
                                                 dim3(DR_THREADS), args, DP_SMEM, stream));
        } else if (variant != 1) {
            static bool configured4 = false;
            if (!configured4) {
                GPC_CUDA(cudaFuncSetAttribute(dense_runs_staged_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM));
                configured4 = true;
            }
            static const int dbg = getenv("GPC_DENSE_DBG") ? atoi(getenv("GPC_DENSE_DBG")) : 0;
            const unsigned long long* bases = nullptr;
            if (variant != 2) {          // reduce, then scan (st_tot / st_cnt are one [2][tiles] array)
                GPC_PROF_BYTES("dense_tile_sums", 8.0 * n_pos);
                dense_tile_sums_kernel<<<tiles, DR_THREADS, 0, stream>>>(dense, st_tot, st_super, groups);
                GPC_LAUNCH_CHECK();
                bases = st_tot;
            }
            GPC_PROF_BYTES("dense_runs", 16.0 * n_pos);
            dense_runs_staged_kernel<<<tiles, DR_THREADS, DS_SMEM, stream>>>(
                dense, n_pos, frag_idx, frag_val, direct_idx, direct_val, capacity, st_tot, st_cnt,
                ticket, counts, bases, st_super, groups, dbg);
        } else {
            GPC_PROF_BYTES("dense_runs", 16.0 * n_pos);
            dense_runs_block_kernel<16><<<tiles, DR_THREADS, DR_THREADS * (16 * 8 + 16), stream>>>(
                dense, n_pos, frag_idx, frag_val, direct_idx, direct_val, capacity, st_tot, st_cnt,
                ticket, counts);
        }
        GPC_LAUNCH_CHECK();
        rel.clean = true;
        GPC_READ_BACK({counts, h, (unsigned)(32)});
        if ((int)(h[3] & 0xffffffffu) != GPC_ERR_FRONTIER_OVERFLOW || (graph->flags & GPC_GRAPH_UNORDERED))
            break;
    }
    *n_frag = h[0];
    *n_direct = h[1];
    *n_events = h[2];
    prof_add_bytes("walk_reads", walk_read_bytes(reads) * (double)n_reads + 8.0 * (double)h[2]);
    prof_add_bytes("dense_runs", 8.0 * (double)(h[0] + h[1]));
    GPC_TRY(walk_error((int)(h[3] & 0xffffffffu)));
    if (h[0] > capacity || h[1] > capacity) return GPC_ERR_CAPACITY;
    return GPC_OK;
}

extern "C" int gpc_events_to_runs(uint32_t* events, uint32_t* events_tmp, uint64_t n_events,
                                  uint32_t size_bp, uint32_t* frag_idx, int32_t* frag_val,
                                  uint64_t* n_frag, uint32_t* direct_idx, int32_t* direct_val,
                                  uint64_t* n_direct, void* stream_) {
    GPC_STAGE();
    return events_to_runs_impl(events, events_tmp, n_events, size_bp, frag_idx, frag_val, n_frag,
                               direct_idx, direct_val, n_direct, n_events + 1, (cudaStream_t)stream_);
}

static int events_to_runs_impl(uint32_t* events, uint32_t* events_tmp, uint64_t n_events,
                               uint32_t size_bp, uint32_t* frag_idx, int32_t* frag_val,
                               uint64_t* n_frag, uint32_t* direct_idx, int32_t* direct_val,
                               uint64_t* n_direct, uint64_t capacity, cudaStream_t stream) {
    if (n_events == 0) {   // empty track: one entry (0, 0)
        GPC_CUDA(cudaMemsetAsync(frag_idx, 0, 4, stream));
        GPC_CUDA(cudaMemsetAsync(frag_val, 0, 4, stream));
        GPC_CUDA(cudaMemsetAsync(direct_idx, 0, 4, stream));
        GPC_CUDA(cudaMemsetAsync(direct_val, 0, 4, stream));
        *n_frag = *n_direct = 1;
        return GPC_OK;
    }
    int bits = EVENT_POS_SHIFT;
    while (bits < 32 && (((unsigned long long)size_bp << EVENT_POS_SHIFT) >> bits)) ++bits;
    bool in_tmp = false;
    // only the position bits have to be ordered: the scan groups equal positions
    // and adds their deltas in any order
    GPC_TRY((radix_sort<uint32_t, false>(events, events_tmp, nullptr, nullptr, n_events,
                                         EVENT_POS_SHIFT, bits, false, &in_tmp, stream)));
    const uint32_t* sorted = in_tmp ? events_tmp : events;
    unsigned tiles = div_up(n_events, RUN_TILE);
    size_t bytes = (size_t)tiles * 32 + 64;
    Scratch sc;
    GPC_CUDA(sc.alloc(bytes, stream));
    GPC_CUDA(cudaMemsetAsync(sc.p, 0, bytes, stream));
    char* p = sc.as<char>();
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(p);            // 2
    unsigned* ticket = reinterpret_cast<unsigned*>(p + 16);
    unsigned long long* st_tot = reinterpret_cast<unsigned long long*>(p + 64);
    unsigned long long* st_seg = st_tot + tiles;
    unsigned long long* st2 = st_seg + tiles;
    unsigned long long* st3 = st2 + tiles;
    GPC_PROF_BYTES("events_to_runs", 4.0 * n_events);
    events_to_runs_kernel<<<tiles, RUN_THREADS, 0, stream>>>(
        sorted, (unsigned)n_events, frag_idx, frag_val, direct_idx, direct_val, st_tot, st_seg,
        st2, st3, ticket, counts, capacity);
    GPC_LAUNCH_CHECK();
    unsigned long long h[2];
    GPC_READ_BACK({counts, h, (unsigned)(16)});
    *n_frag = h[0];
    *n_direct = h[1];
    prof_add_bytes("events_to_runs", 8.0 * (double)(h[0] + h[1]));
    return (h[0] > capacity || h[1] > capacity) ? GPC_ERR_CAPACITY : GPC_OK;
}
