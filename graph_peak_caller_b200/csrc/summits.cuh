// Peak summits: PeakCollection.cut_around_summit of the reference
// (peakcollection.py:113-136) without the dense G-sized q-value array its caller
// builds (analysis/analysis_interface.py:388-406, mindense.py:25-59).
//
// For one peak -- start offset on its first node, end offset on its last node, node
// list -- the q-values along the peak are the concatenation of the run-length track
// over the pieces [node_off + start, node_off + len), [node_off, node_off + len), ...,
// [node_off, node_off + end).  The summit is the (count // 2)-th smallest of the
// positions (offsets along the peak) that hold the maximum value
// (np.partition(max_positions, size // 2)[size // 2]); the peak is then cut to
// [max(0, summit - w), min(summit + w, length)) with obg's get_subinterval rule.
//
// One sequential function per peak, shared by the kernel and by the CPU test build
// (tests/hostsim/summits_host.cpp), so that the exact code the kernel runs is checked
// against the live reference without a GPU.
#pragma once
#include <stdint.h>

#ifndef GPC_HD
#ifdef __CUDACC__
#define GPC_HD __host__ __device__ __forceinline__
#else
#define GPC_HD inline
#endif
#endif

namespace gpc {

struct SummitCut {
    int32_t summit;      // offset of the summit along the peak
    int32_t length;      // base pairs of the peak
    int32_t first;       // index (in the peak's node list) of the first node of the cut
    int32_t last;        // ... of its last node
    int32_t start_off;   // start offset on node `first`
    int32_t end_off;     // end offset on node `last`
};

// index of the last entry of pos[0..n) that is <= x, or -1
GPC_HD int64_t last_not_after(const uint32_t* pos, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (pos[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return (int64_t)lo - 1;
}

// Walks the run-length track over the pieces of one peak and calls
// seg(value, offset along the peak, number of positions) for every maximal stretch of
// equal track entries inside a piece; returns the peak's length.  seg returns false to
// stop the walk.
template <typename Seg>
GPC_HD int64_t walk_peak(const uint32_t* node_rec, int32_t min_node, int32_t start, int32_t end,
                         const int32_t* nodes, uint32_t n_peak_nodes, const uint32_t* q_pos,
                         const double* q_val, uint32_t n_q, Seg&& seg) {
    int64_t along = 0;
    for (uint32_t k = 0; k < n_peak_nodes; ++k) {
        const uint32_t v = (uint32_t)(nodes[k] - min_node);
        const uint32_t node_off = node_rec[4 * v], node_len = node_rec[4 * v + 3];
        const uint32_t a = node_off + (k == 0 ? (uint32_t)start : 0u);
        const uint32_t b = node_off + (k + 1 == n_peak_nodes ? (uint32_t)end : node_len);
        if (b <= a) continue;
        int64_t r = last_not_after(q_pos, n_q, a);
        uint32_t lo = a;
        while (lo < b) {
            const uint32_t next = (r + 1 < (int64_t)n_q) ? q_pos[r + 1] : 0xffffffffu;
            const uint32_t hi = next < b ? next : b;
            const double value = r >= 0 ? q_val[r] : 0.0;
            if (!seg(value, along, (int64_t)(hi - lo))) return -1;
            along += hi - lo;
            lo = hi;
            ++r;
        }
    }
    return along;
}

GPC_HD SummitCut peak_summit_cut(const uint32_t* node_rec, int32_t min_node, int32_t start,
                                 int32_t end, const int32_t* nodes, uint32_t n_peak_nodes,
                                 const uint32_t* q_pos, const double* q_val, uint32_t n_q,
                                 int32_t window) {
    SummitCut out = {0, 0, 0, 0, start, start};
    // pass 1: the maximum and how many positions hold it
    double best = 0.0;
    int64_t count = 0;
    bool any = false;
    const int64_t length = walk_peak(node_rec, min_node, start, end, nodes, n_peak_nodes, q_pos,
                                     q_val, n_q, [&](double v, int64_t, int64_t n) {
        if (!any || v > best) { best = v; count = n; any = true; }
        else if (v == best) count += n;
        return true;
    });
    if (length <= 0) return out;
    out.length = (int32_t)length;
    // pass 2: the (count // 2)-th position holding the maximum
    const int64_t target = count / 2;
    int64_t seen = 0, summit = 0;
    walk_peak(node_rec, min_node, start, end, nodes, n_peak_nodes, q_pos, q_val, n_q,
              [&](double v, int64_t along, int64_t n) {
        if (v != best) return true;
        if (seen + n > target) { summit = along + (target - seen); return false; }
        seen += n;
        return true;
    });
    out.summit = (int32_t)summit;
    // the cut [from, to) along the peak -> positions on nodes: the cut starts on the first
    // node that reaches beyond `from`, and ends on the first node that reaches `to`
    const int64_t from = summit - window > 0 ? summit - window : 0;
    const int64_t to = summit + window < length ? summit + window : length;
    int64_t walked = -(int64_t)start;
    bool have_first = false;
    for (uint32_t k = 0; k < n_peak_nodes; ++k) {
        const uint32_t v = (uint32_t)(nodes[k] - min_node);
        const int64_t size = node_rec[4 * v + 3];
        if (!have_first && walked + size > from) {
            out.first = (int32_t)k;
            out.start_off = (int32_t)(from - walked);
            have_first = true;
        }
        if (walked + size >= to) {
            out.last = (int32_t)k;
            out.end_off = (int32_t)(to - walked);
            break;
        }
        walked += size;
    }
    return out;
}

}  // namespace gpc
