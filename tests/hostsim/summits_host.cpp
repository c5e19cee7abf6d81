// TEST INFRASTRUCTURE: compiles the per-peak function of
// graph_peak_caller_b200/csrc/summits.cuh -- the code the summit kernel runs, one
// thread per peak -- for the CPU, so that the CPU test suite can hold it against the
// oracle and the live reference.
#include "../../graph_peak_caller_b200/csrc/summits.cuh"

extern "C" void peak_summits_host(const uint32_t* node_rec, int32_t min_node,
                                  const int32_t* peak_start, const int32_t* peak_end,
                                  const uint32_t* peak_node_ptr, const int32_t* peak_nodes,
                                  long n_peaks, const uint32_t* q_pos, const double* q_val,
                                  long n_q, int32_t window, int32_t* out) {
    for (long i = 0; i < n_peaks; ++i) {
        const uint32_t lo = peak_node_ptr[i], hi = peak_node_ptr[i + 1];
        const gpc::SummitCut c = gpc::peak_summit_cut(node_rec, min_node, peak_start[i],
                                                      peak_end[i], peak_nodes + lo, hi - lo, q_pos,
                                                      q_val, (uint32_t)n_q, window);
        int32_t* o = out + 6 * i;
        o[0] = c.summit; o[1] = c.length; o[2] = c.first; o[3] = c.last;
        o[4] = c.start_off; o[5] = c.end_off;
    }
}
