// Fragment-length estimation: the sequential scans of the
// reference's MACS2-style peak model (shiftestimation/shiftestimation.py:65-336)
// and of its path choice through the graph (haplotyping.py:28-58), on the host.
//
// The model works on the sorted read-start positions of one chromosome and strand
// ("tags").  Every step is a single left-to-right sweep whose next window starts
// where the previous one ended, so there is no data-parallel form worth a kernel:
// 10^7 tags take a few 10 ms here against minutes of Python loops in the reference.
// What is left to the Python side (shiftestimation.py of this package) is the
// cross-correlation of two 1211-entry lines.
// Host code only: no kernel, all pointers are host pointers.
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gpc {
namespace {

inline int64_t floor_half(int64_t x) { return x >= 0 ? x / 2 : -((-x + 1) / 2); }

struct StrandPeak {
    int64_t pos;
    int64_t tags;
};

// Summit of one window of tags (shiftestimation.py:292-336): every distinct tag
// position covers [p, p + expansion) on a line that starts at the window's first
// tag (numpy's `line[idx] += 1` counts a repeated index once, so repeated tags
// count once); the summit is the middle one of the positions with the highest cover.
int64_t window_summit(const int64_t* tags, int64_t n, int64_t expansion, std::vector<int32_t>& line,
                      std::vector<int64_t>& top) {
    const int64_t first = tags[0], half = expansion / 2;
    const int64_t length = tags[n - 1] + 1 - first + expansion;
    line.assign((size_t)length, 0);
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || tags[i] != tags[i - 1]) line[(size_t)(tags[i] - first)] = 1;
    for (int64_t i = 0; i < n; ++i)
        if (i == 0 || tags[i] != tags[i - 1]) line[(size_t)(tags[i] - first + 2 * half)] -= 1;
    int32_t cover = 0, best = INT32_MIN;
    top.clear();
    for (int64_t x = 0; x < length; ++x) {
        cover += line[(size_t)x];
        if (cover > best) { best = cover; top.clear(); }
        if (cover == best) top.push_back(x);
    }
    return top[top.size() / 2] + first - half;
}

// shiftestimation.py:262-290: cut the tags into windows no wider than `peaksize`,
// each starting at the tag that closed the one before; a closed window with
// min_tags..max_tags tags is a peak.  The window still open at the end is dropped.
void strand_peaks(const int64_t* tags, int64_t n, int64_t peaksize, int64_t expansion,
                  int64_t min_tags, int64_t max_tags, std::vector<StrandPeak>& out) {
    out.clear();
    if (n < 2) return;
    std::vector<int32_t> line;
    std::vector<int64_t> top;
    int64_t open = 0;
    for (int64_t i = 1; i < n; ++i) {
        if (tags[i] - tags[open] + 1 > peaksize) {
            const int64_t count = i - open;
            if (count >= min_tags && count <= max_tags)
                out.push_back({window_summit(tags + open, count, expansion, line, top), count});
            open = i;
        }
    }
}

// shiftestimation.py:231-260: centres of (plus, minus) peak pairs no further apart
// than `peaksize` with comparable tag counts, plus peak to the left.
void pair_centers(const std::vector<StrandPeak>& plus, const std::vector<StrandPeak>& minus,
                  int64_t peaksize, std::vector<int64_t>& centers) {
    size_t ip = 0, im = 0, im_back = 0;
    bool overlapping = false;
    while (ip < plus.size() && im < minus.size()) {
        const StrandPeak p = plus[ip], m = minus[im];
        if (p.pos - peaksize > m.pos) {
            ++im;
        } else if (p.pos + peaksize < m.pos) {
            ++ip;
            im = im_back;
            overlapping = false;
        } else {
            if (!overlapping) { overlapping = true; im_back = im; }
            const double ratio = (double)p.tags / (double)m.tags;
            if (ratio < 2 && ratio > 0.5 && p.pos < m.pos) centers.push_back(floor_half(p.pos + m.pos));
            ++im;
        }
    }
}

// shiftestimation.py:180-218: every tag within peaksize + expansion/2 of a centre
// marks where its expanded copy starts and ends on the model line around the centre.
void add_to_line(const int64_t* centers, size_t n_centers, const int64_t* tags, int64_t n_tags,
                 int64_t peaksize, int64_t expansion, int32_t* start, int32_t* end, int64_t window) {
    const int64_t reach = peaksize + expansion / 2;
    size_t ic = 0;
    int64_t it = 0, it_back = 0;
    bool overlapping = false;
    while (ic < n_centers && it < n_tags) {
        const int64_t c = centers[ic], t = tags[it];
        if (c - reach > t) {
            ++it;
        } else if (c + reach < t) {
            ++ic;
            it = it_back;
            overlapping = false;
        } else {
            if (!overlapping) { overlapping = true; it_back = it; }
            start[std::max<int64_t>(t - c + peaksize, 0)] += 1;
            end[std::min<int64_t>(t + expansion - c + peaksize, window - 1)] -= 1;
            ++it;
        }
    }
}

}  // namespace
}  // namespace gpc

using namespace gpc;

extern "C" int gpc_shift_model_host(const int64_t* plus_tags, uint64_t n_plus,
                                    const int64_t* minus_tags, uint64_t n_minus, int64_t peaksize,
                                    int64_t tag_expansion, int64_t min_tags, int64_t max_tags,
                                    int32_t* plus_start, int32_t* plus_end, int32_t* minus_start,
                                    int32_t* minus_end, int64_t* centers, uint64_t* n_centers) {
    if (peaksize <= 0 || tag_expansion < 0 || !n_centers || !plus_start || !plus_end ||
        !minus_start || !minus_end) {
        set_error("gpc_shift_model_host: bad arguments");
        return GPC_ERR_ARG;
    }
    for (uint64_t i = 1; i < n_plus; ++i)
        if (plus_tags[i] < plus_tags[i - 1]) { set_error("gpc_shift_model_host: plus tags not sorted"); return GPC_ERR_ARG; }
    for (uint64_t i = 1; i < n_minus; ++i)
        if (minus_tags[i] < minus_tags[i - 1]) { set_error("gpc_shift_model_host: minus tags not sorted"); return GPC_ERR_ARG; }
    std::vector<StrandPeak> plus, minus;
    strand_peaks(plus_tags, (int64_t)n_plus, peaksize, tag_expansion, min_tags, max_tags, plus);
    strand_peaks(minus_tags, (int64_t)n_minus, peaksize, tag_expansion, min_tags, max_tags, minus);
    std::vector<int64_t> found;
    if (!plus.empty() && !minus.empty()) pair_centers(plus, minus, peaksize, found);
    const uint64_t capacity = *n_centers;
    *n_centers = found.size();
    if (centers) {
        if (capacity < found.size()) return GPC_ERR_CAPACITY;
        std::copy(found.begin(), found.end(), centers);
    }
    const int64_t window = 1 + 2 * peaksize + tag_expansion;
    add_to_line(found.data(), found.size(), plus_tags, (int64_t)n_plus, peaksize, tag_expansion,
                plus_start, plus_end, window);
    add_to_line(found.data(), found.size(), minus_tags, (int64_t)n_minus, peaksize, tag_expansion,
                minus_start, minus_end, window);
    return GPC_OK;
}

// haplotyping.py:28-58: from the graph's only source, follow the successor most reads
// pass through (the first of equals in adjacency order) down to a sink.  A node without
// any edge is not a source: the reference's tests/test_haplotyping.py:10-25,50-66 holds
// such a node (11) and still finds one start node.
extern "C" int gpc_most_covered_path_host(const int64_t* succ_ptr, const int32_t* succ,
                                          const int64_t* pred_ptr, const int64_t* node_count,
                                          uint64_t n_nodes, int32_t* path, uint64_t* n_path) {
    if (!succ_ptr || !pred_ptr || !node_count || !n_path) {
        set_error("gpc_most_covered_path_host: bad arguments");
        return GPC_ERR_ARG;
    }
    int64_t source = -1;
    for (uint64_t v = 0; v < n_nodes; ++v)
        if (pred_ptr[v + 1] == pred_ptr[v] && (succ_ptr[v + 1] > succ_ptr[v] || n_nodes == 1)) {
            if (source >= 0) { set_error("Only works when graph has one start node"); return GPC_ERR_ARG; }
            source = (int64_t)v;
        }
    if (source < 0) { set_error("Only works when graph has one start node"); return GPC_ERR_ARG; }
    const uint64_t capacity = *n_path;
    uint64_t n = 0;
    for (int64_t v = source;;) {
        if (n >= n_nodes) { set_error("gpc_most_covered_path_host: the graph has a cycle"); return GPC_ERR_ARG; }
        if (path && n < capacity) path[n] = (int32_t)v;
        ++n;
        if (succ_ptr[v + 1] == succ_ptr[v]) break;
        int64_t next = -1, best = -1;
        for (int64_t e = succ_ptr[v]; e < succ_ptr[v + 1]; ++e)
            if (node_count[succ[e]] > best) { best = node_count[succ[e]]; next = succ[e]; }
        v = next;
    }
    *n_path = n;
    return (path && capacity < n) ? GPC_ERR_CAPACITY : GPC_OK;
}
