/* gpc_b200.h -- C ABI of the B200-native callpeaks core.
 *
 * The reference (uio-bmi/graph_peak_caller v1.2.3) is pure Python and has no
 * FFI of its own; its drop-in boundary is the Python API of SURVEY.md §8(b).
 * This header is the boundary one level below it: every entry point replaces
 * the body of the reference function(s) named in its comment, takes plain
 * device pointers and sizes, returns an int status (GPC_OK == 0) and never
 * throws.  graph_peak_caller_b200/ binds it with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers unless the name ends in _host
 *     or the parameter is a scalar out-parameter (uint64_t* n_xxx): those are
 *     host pointers, written before the function returns (the function
 *     synchronises the stream when it has to report a count);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - outputs are caller-allocated with the documented capacity; temporary
 *     scratch comes from the stream-ordered pool (cudaMallocAsync);
 *   - positions are 0-based offsets into the id-ordered concatenation of the
 *     nodes (the reference's graph.node_indexes coordinates), uint32;
 *   - integer pileups are int32, lambda / p / q values are float64.
 */
#ifndef GPC_B200_H
#define GPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    GPC_OK = 0,
    GPC_ERR_CUDA = 1,              /* a CUDA runtime call failed            */
    GPC_ERR_ARG = 2,               /* invalid argument                      */
    GPC_ERR_CAPACITY = 3,          /* output capacity too small: retry      */
    GPC_ERR_INVALID_INTERVAL = 4,  /* reference InvalidPileupInterval       */
    GPC_ERR_FRONTIER_OVERFLOW = 5, /* too many simultaneous branches        */
    GPC_ERR_INTERNAL = 6
};

/* Thread-local text of the last error of the calling thread. */
const char* gpc_last_error_string(void);
/* Library / build identification, e.g. "gpc_b200 0.1 sm_100a". */
const char* gpc_version_string(void);

/* Graph in HBM.  node_rec holds n_nodes+1 records of four uint32:
 *   { node_off, succ_ptr, pred_ptr, node_len }  (record n: {size_bp, m, m, 0})
 * i.e. obg's node_indexes / adj_list / reverse_adj_list of one node in one
 * 16-byte load.  succ / pred are 0-based node indexes; every edge ascends. */
typedef struct gpc_graph {
    uint32_t n_nodes;
    uint32_t n_edges;
    uint32_t size_bp;
    int32_t min_node;
    const uint32_t* node_rec;
    const uint32_t* succ;
    const uint32_t* pred;
    const double* lin_start; /* LinearMap._node_starts (may be NULL) */
    const double* lin_end;   /* LinearMap._node_ends   (may be NULL) */
} gpc_graph_t;

/* Reads in HBM, structure of arrays (signed node ids: negative == reverse
 * strand, as obg Interval.region_paths). */
typedef struct gpc_reads {
    uint64_t n_reads;
    const int32_t* start_node;
    const int32_t* start_off;
    const int32_t* end_node;
    const int32_t* end_off;
    const int64_t* path_ptr; /* n_reads + 1 */
    const int32_t* path;
} gpc_reads_t;

/* ---- S1  UniqueIntervals.__iter__  (graph_peak_caller/intervals.py:23-33) ---
 * Keeps the first read of every start position (signed node id, offset).
 * out_index: capacity n_reads, receives the kept read indexes in ascending
 * order; *n_unique their number. */
int gpc_unique_reads(const int32_t* start_node, const int32_t* start_off, uint64_t n_reads,
                     uint32_t* out_index, uint64_t* n_unique, void* stream);

/* ---- S2-S4  ReadsAdderWDirect.add_reads, SparseExtender.run_linear,
 *            ReverseSparseExtender  (sample/sparsegraphpileup.py:37-211) ------
 * Walks every read (read_index[i], or i when read_index is NULL), extends its
 * end by `extension` = fragment_length - read_length along all paths and
 * appends packed +1/-1 events (position << 3 | tag << 1 | is_minus; tag 0:
 * both pileups, 1: direct only, 2: fragment only).  touched: n_nodes bytes,
 * set to 1 for every node a read or an extension reached.
 * Returns GPC_ERR_CAPACITY with *n_events = required capacity when `events`
 * is too small (nothing else is lost: call again). */
int gpc_sample_events(const gpc_graph_t* graph, const gpc_reads_t* reads,
                      const uint32_t* read_index, uint64_t n_reads, int32_t extension,
                      uint32_t* events, uint64_t event_capacity, uint64_t* n_events,
                      uint8_t* touched, void* stream);

/* ---- T1/T2  SparseDiffs.from_pileup + get_sparse_values
 *            (sparsediffs.py:137-141, 180-190), also get_direct_pileup
 *            (sample/sparsegraphpileup.py:226-237) -------------------------
 * Radix-sorts the events (events_tmp: same size, ping-pong buffer; both are
 * clobbered) and writes the two canonical run-length tracks: strictly
 * increasing positions starting at 0, neighbouring values different.
 * Capacity of each output: n_events + 1. */
int gpc_events_to_runs(uint32_t* events, uint32_t* events_tmp, uint64_t n_events,
                       uint32_t size_bp, uint32_t* frag_idx, int32_t* frag_val, uint64_t* n_frag,
                       uint32_t* direct_idx, int32_t* direct_val, uint64_t* n_direct,
                       void* stream);

/* ---- building blocks (np.argsort / np.cumsum of the reference's
 *      sparsediffs.py, exported for the host side and the parity tests) ------
 * LSD radix sorts over bits [begin_bit, end_bit); keys/vals ping-pong with
 * the *_tmp buffers, *result_in_tmp tells where the sorted data is.  The
 * 64-bit variant skips digits that are equal for all keys (vals may be NULL).
 * n < 2^30. */
int gpc_sort_u32(uint32_t* keys, uint32_t* keys_tmp, uint64_t n, int32_t begin_bit,
                 int32_t end_bit, int32_t* result_in_tmp, void* stream);
int gpc_sort_u64_pairs(uint64_t* keys, uint64_t* keys_tmp, uint32_t* vals, uint32_t* vals_tmp,
                       uint64_t n, int32_t begin_bit, int32_t end_bit, int32_t* result_in_tmp,
                       void* stream);
/* out[i] = sum of in[0..i); *total = sum of all (host pointer). */
int gpc_exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint64_t* total,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPC_B200_H */
