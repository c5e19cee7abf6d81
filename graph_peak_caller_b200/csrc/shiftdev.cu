// Fragment-length estimation, the whole-array steps on sm_100a.
//
// The estimate has three parts.  Two of them are one operation per read / per path node
// and run here, on the reads as they already sit in HBM for the callpeaks pass:
//   HaploTyper.build                     haplotyping.py:18-26   visits per node
//   LinearFilter.find_start_positions    linear_filter.py:22-46 read starts projected onto
//                                        the most-covered path, per strand, in read order
// The third -- the walk along the most-covered path (haplotyping.py:28-58) and the sweeps
// of the peak model over the SORTED starts (shiftestimation.py:231-336) -- is sequential
// by construction (csrc/shift.cu, host); the sort in between is gpc_sort_u64_pairs.
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace gpc {

// counts[v] += 1 for every forward visit (path id > 0) of node v: the reference keys its
// counter by signed id and the most-covered path asks for forward ids only
__global__ void node_read_counts_kernel(const int32_t* __restrict__ path, unsigned long long n,
                                        int32_t min_node, uint32_t n_nodes, uint32_t* __restrict__ counts) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t id = path[i];
        if (id <= 0) continue;
        const long long v = (long long)id - min_node;
        if (v >= 0 && v < (long long)n_nodes) atomicAdd(&counts[v], 1u);     // (RED: nobody reads the result)
    }
}

__global__ void path_lengths_kernel(gpc_graph_t g, const int32_t* __restrict__ path, unsigned n_path,
                                    uint32_t* __restrict__ len, int* __restrict__ bad) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_path) return;
    const int32_t v = path[i];
    if (v < 0 || (uint32_t)v >= g.n_nodes) {
        *bad = 1;
        len[i] = 0;
        return;
    }
    len[i] = node_record(g, (uint32_t)v).w;
}

__global__ void path_distance_kernel(const int32_t* __restrict__ path, unsigned n_path,
                                     const uint32_t* __restrict__ offs, uint32_t n_nodes,
                                     uint32_t* __restrict__ dist /* 0xffffffff: not on the path */) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_path) return;
    const int32_t v = path[i];
    if (v >= 0 && (uint32_t)v < n_nodes) dist[v] = offs[i];
}

// linear_filter.py:27-44: a start on a forward node lies `offset` behind the node's start on
// the path; a start on a reverse node (id < 0) is the same base pair counted on the forward
// strand, node size - offset (the rule the reference's known answer allows)
struct StrandStartsOp {
    gpc_graph_t g;
    const int32_t* start_node;
    const int32_t* start_off;
    const uint32_t* dist;
    long long start_offset;
    int forward;
    long long* out;
    __device__ __forceinline__ long long node_of(unsigned i) const {
        const int32_t id = start_node[i];
        const long long a = id < 0 ? -(long long)id : (long long)id;
        return a - g.min_node;
    }
    __device__ __forceinline__ bool keep(unsigned i) const {
        if ((start_node[i] > 0) != (forward != 0)) return false;
        const long long v = node_of(i);
        return v >= 0 && v < (long long)g.n_nodes && dist[v] != 0xffffffffu;
    }
    __device__ __forceinline__ void emit(unsigned i, unsigned long long w) const {
        const long long v = node_of(i);
        const long long d = (long long)dist[v] - start_offset;
        out[w] = forward ? d + start_off[i]
                         : d + (long long)node_record(g, (uint32_t)v).w - start_off[i];
    }
    __device__ __forceinline__ void finish(unsigned long long) const {}
};

}  // namespace gpc

using namespace gpc;

extern "C" int gpc_node_read_counts(const int32_t* path, uint64_t n_path_nodes, int32_t min_node,
                                    uint32_t n_nodes, uint32_t* counts, void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    GPC_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_nodes * 4, stream));
    if (n_path_nodes == 0 || n_nodes == 0) return GPC_OK;
    const unsigned blocks = (unsigned)std::min<uint64_t>(div_up(n_path_nodes, 256), (uint64_t)sm_count() * 16);
    GPC_PROF_BYTES("node_read_counts", (double)n_path_nodes * 4);
    node_read_counts_kernel<<<blocks, 256, 0, stream>>>(path, n_path_nodes, min_node, n_nodes, counts);
    GPC_LAUNCH_CHECK();
    return GPC_OK;
}

extern "C" int gpc_path_start_positions(const gpc_graph_t* graph, const int32_t* path, uint64_t n_path,
                                        int64_t start_offset, const int32_t* start_node,
                                        const int32_t* start_off, uint64_t n_reads, int64_t* out_plus,
                                        uint64_t* n_plus, int64_t* out_minus, uint64_t* n_minus,
                                        void* stream_) {
    GPC_STAGE();
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_plus = *n_minus = 0;
    const uint32_t n = graph->n_nodes;
    if (n_path > n) {
        set_error("gpc_path_start_positions: a path of %llu nodes through a graph of %u",
                  (unsigned long long)n_path, n);
        return GPC_ERR_ARG;
    }
    if (n_reads >= (1ull << 32)) {
        set_error("gpc_path_start_positions: %llu reads, at most 2^32 - 1", (unsigned long long)n_reads);
        return GPC_ERR_ARG;
    }
    Scratch dist, len, offs, counts;
    GPC_CUDA(dist.alloc((size_t)n * 4, stream));
    GPC_CUDA(cudaMemsetAsync(dist.p, 0xff, (size_t)n * 4, stream));
    GPC_CUDA(counts.alloc(32, stream));
    GPC_CUDA(cudaMemsetAsync(counts.p, 0, 32, stream));
    unsigned long long* c_plus = counts.as<unsigned long long>();
    unsigned long long* c_minus = c_plus + 1;
    int* bad = reinterpret_cast<int*>(c_plus + 2);
    if (n_path) {
        GPC_CUDA(len.alloc((size_t)n_path * 4, stream));
        GPC_CUDA(offs.alloc((size_t)n_path * 4, stream));
        GPC_PROF("path_lengths");
        path_lengths_kernel<<<div_up(n_path, 256), 256, 0, stream>>>(*graph, path, (unsigned)n_path,
                                                                      len.as<uint32_t>(), bad);
        GPC_LAUNCH_CHECK();
        GPC_TRY(exclusive_scan<uint32_t>(len.as<uint32_t>(), offs.as<uint32_t>(), n_path, nullptr, stream));
        GPC_PROF("path_distance");
        path_distance_kernel<<<div_up(n_path, 256), 256, 0, stream>>>(path, (unsigned)n_path,
                                                                       offs.as<uint32_t>(), n,
                                                                       dist.as<uint32_t>());
        GPC_LAUNCH_CHECK();
    }
    StrandStartsOp plus{*graph, start_node, start_off, dist.as<uint32_t>(), (long long)start_offset, 1,
                        reinterpret_cast<long long*>(out_plus)};
    StrandStartsOp minus = plus;
    minus.forward = 0;
    minus.out = reinterpret_cast<long long*>(out_minus);
    GPC_TRY(select_if("path_starts_plus", plus, n_reads, c_plus, stream));
    GPC_TRY(select_if("path_starts_minus", minus, n_reads, c_minus, stream));
    unsigned long long h[2] = {0, 0};
    int h_bad = 0;
    GPC_READ_BACK({c_plus, h, 16u}, {bad, &h_bad, 4u});
    if (h_bad) {
        set_error("gpc_path_start_positions: a path node is outside the graph");
        return GPC_ERR_ARG;
    }
    *n_plus = h[0];
    *n_minus = h[1];
    return GPC_OK;
}
