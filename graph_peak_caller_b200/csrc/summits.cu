// Peak summits on the device: one thread per peak runs peak_summit_cut (summits.cuh).
// Peaks number 10^3..10^5 per chromosome and touch a handful of track runs each, so
// the kernel is latency-bound by its two binary searches per piece; what it removes
// is the reference's dense q-value array (8 bytes per base pair of the graph).
#include "common.cuh"
#include "summits.cuh"

using namespace gpc;

namespace {

__global__ void __launch_bounds__(128)
peak_summits_kernel(gpc_graph_t g, const int32_t* __restrict__ peak_start,
                    const int32_t* __restrict__ peak_end,
                    const uint32_t* __restrict__ peak_node_ptr,
                    const int32_t* __restrict__ peak_nodes, unsigned n_peaks,
                    const uint32_t* __restrict__ q_pos, const double* __restrict__ q_val,
                    unsigned n_q, int window, int32_t* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_peaks) return;
    const uint32_t lo = peak_node_ptr[i], hi = peak_node_ptr[i + 1];
    const SummitCut c = peak_summit_cut(g.node_rec, g.min_node, peak_start[i], peak_end[i],
                                        peak_nodes + lo, hi - lo, q_pos, q_val, n_q, window);
    int32_t* o = out + 6 * (size_t)i;
    o[0] = c.summit;
    o[1] = c.length;
    o[2] = c.first;
    o[3] = c.last;
    o[4] = c.start_off;
    o[5] = c.end_off;
}

}  // namespace

extern "C" int gpc_peak_summits(const gpc_graph_t* graph, const int32_t* peak_start,
                                const int32_t* peak_end, const uint32_t* peak_node_ptr,
                                const int32_t* peak_nodes, uint64_t n_peaks,
                                const uint32_t* q_pos, const double* q_val, uint64_t n_q,
                                int32_t window, int32_t* out, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_peaks == 0) return GPC_OK;
    if (window < 0) { set_error("gpc_peak_summits: negative window"); return GPC_ERR_ARG; }
    GPC_PROF("peak_summits");
    peak_summits_kernel<<<div_up(n_peaks, 128), 128, 0, stream>>>(
        *graph, peak_start, peak_end, peak_node_ptr, peak_nodes, (unsigned)n_peaks, q_pos, q_val,
        (unsigned)n_q, (int)window, out);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}
