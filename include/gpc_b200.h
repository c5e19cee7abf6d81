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
 *     host pointers, written before the function returns (the function waits
 *     for its own kernels when it has to report a count); output ARRAYS are
 *     complete in stream order, like any other work enqueued on `stream`;
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
    /* Optional (may be NULL): n_nodes records of eight uint32, 32-byte aligned,
     *   { node_off, node_len, n_succ | n_pred << 16, succ0, succ1, pred0, pred1, 0 }
     * so that the read walk needs one 32-byte load per visited node; nodes with
     * more than two successors / predecessors fall back to the CSR arrays. */
    const uint32_t* walk_rec;
    /* bit 0 (GPC_GRAPH_UNORDERED): some edge does not ascend in node id (a cyclic graph,
     * or an unsorted one).  Only the sample pileup accepts such graphs: it then walks
     * every read with a label-correcting worklist (visited[node] = max remaining, the
     * reference's legacy/extender.py:273-284) instead of the id-ordered frontier. */
    uint32_t flags;
} gpc_graph_t;
#define GPC_GRAPH_UNORDERED 1u
/* bit 1: the linear span (lin_end - lin_start) of every node is a whole multiple of its
 * length, so every read start projects to an integer (control/linearmap.py:43-74): the
 * control stage then sorts 32-bit integer keys.  Set by whoever uploads the linear map. */
#define GPC_GRAPH_INTEGER_LINEAR 2u

/* Reads in HBM, structure of arrays (signed node ids: negative == reverse
 * strand, as obg Interval.region_paths). */
typedef struct gpc_reads {
    uint64_t n_reads;
    const int32_t* start_node;
    const int32_t* start_off;
    const int32_t* end_node;
    const int32_t* end_off;
    const int64_t* path_ptr; /* n_reads + 1 (may be NULL when path_ptr32 is given) */
    const int32_t* path;
    const uint32_t* path_ptr32; /* optional: the same offsets in 32 bits (fewer bytes to
                                   upload when the paths hold < 2^32 nodes); used if non-NULL */
    /* Optional (may be NULL): n_reads records of four uint32, written by
     * gpc_build_read_headers:
     *   { path offset, path nodes | inconsistent << 31, start offset, end offset }
     * When present the read walk loads this one record per read (its start / end node
     * are the first / last node of its path) and ignores end_node, end_off and path_ptr*. */
    const uint32_t* hdr;
} gpc_reads_t;

/* Builds gpc_reads_t.hdr (capacity 4 * n_reads uint32) from the arrays of `reads`
 * (path_ptr or path_ptr32, start / end node and offset, path); the record of a read
 * whose start / end node is not the first / last node of its path, or whose path is
 * empty, is marked inconsistent and reported as GPC_ERR_INVALID_INTERVAL by the walk
 * (reference: InvalidPileupInterval, sample/sparsegraphpileup.py:47-62). */
int gpc_build_read_headers(const gpc_reads_t* reads, uint32_t* hdr, void* stream);

/* Reads as they cross the host link (15 bytes per read instead of 30 for 36-bp reads:
 * the start / end node repeat the ends of the path, offsets fit 16 bits, path offsets
 * are a prefix sum) -> the arrays the kernels read:
 *   offs[r] = start offset | end offset << 16,  path_len[r] = nodes on the path (1..255),
 *   path = the signed node ids of all paths, back to back.
 * Writes start_node / start_off (n_reads each: the duplicate filter and the control
 * projection read them) and hdr (4 * n_reads, see gpc_reads_t.hdr).  A read with
 * path_len 0 is marked inconsistent (-> GPC_ERR_INVALID_INTERVAL from the walk). */
int gpc_reads_from_compact(const uint32_t* offs, const uint8_t* path_len, const int32_t* path,
                           uint64_t n_reads, int32_t* start_node, int32_t* start_off,
                           uint32_t* hdr, void* stream);

/* The same with delta-coded paths (10.3 bytes per 36-bp read): first_node[r] = the first
 * node id of read r's path (int32), step = for every further path node the int8 difference
 * to the node before it, the steps of all reads back to back (read r owns path_len[r] - 1 of
 * them).  Also writes `path` (capacity: sum of path_len), the array gpc_reads_t.path points
 * at.  A host that meets a step outside int8 ships the batch in the plain compact form. */
int gpc_reads_from_compact_delta(const uint32_t* offs, const uint8_t* path_len,
                                 const int32_t* first_node, const int8_t* step, uint64_t n_reads,
                                 int32_t* path, int32_t* start_node, int32_t* start_off,
                                 uint32_t* hdr, void* stream);

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

/* ---- T1/T2, generic form: tracks from ARBITRARY event lists ---------------------------
 * What the reference's container does with any (index, diff) list a caller hands it --
 * float diffs, unsorted, indices repeated -- as opposed to the packed +-1 events of the
 * callpeaks path above.  indices: int64 (0 <= index < 2^32 - 1, else GPC_ERR_ARG), diffs:
 * float64, fewer than 2^30 entries per call; outputs need room for every input entry
 * (n, or n_a + n_b).  Running sums are float64 and parallel: bit-identical to np.cumsum
 * for integer-valued / exactly representable steps, equal to rounding otherwise
 * (csrc/tracks.cu).
 *
 * gpc_diffs_to_runs: SparseDiffs.get_sparse_values (sparsediffs.py:137-141: stable sort
 *   by index, running sum, SparseValues._sanitize :13-20) -> canonical runs; with the
 *   event list of SparseDiffs.from_pileup (:180-190) as input, the pileup itself. */
int gpc_diffs_to_runs(const int64_t* indices, const double* diffs, uint64_t n, uint32_t* out_pos,
                      double* out_val, uint64_t* n_out, void* stream);
/* gpc_diffs_merge: the front half of SparseDiffs.maximum / apply_binary_func
 *   (sparsediffs.py:143-154, 208-219): the distinct indices of both lists in ascending
 *   order and BOTH tracks' running sums there (a track is 0 before its first entry) --
 *   the two arrays the reference hands to the caller's function.  As there (:149-153)
 *   the k-th slot of a track in merged order takes that track's k-th diff: for
 *   index-sorted lists, the only ones the reference itself makes, every diff thereby stays
 *   with its index. */
int gpc_diffs_merge(const int64_t* idx_a, const double* diff_a, uint64_t n_a, const int64_t* idx_b,
                    const double* diff_b, uint64_t n_b, uint32_t* out_pos, double* out_a,
                    double* out_b, uint64_t* n_out, void* stream);
/* gpc_diffs_combine: merge + op(a, b) + drop of entries repeating their predecessor's
 *   value.  op GPC_TRACK_MAX with drop_leading_zero = 1 is SparseDiffs.maximum
 *   (sparsediffs.py:143-159: np.max, ediff1d, _sanitize :161-165), returned as runs
 *   (position, value); against the single event (first index, min_value) it is
 *   SparseDiffs.clip_min (:130-135).  NaN propagates as in numpy. */
enum { GPC_TRACK_MAX = 0, GPC_TRACK_MIN = 1, GPC_TRACK_ADD = 2, GPC_TRACK_SUB = 3, GPC_TRACK_MUL = 4 };
int gpc_diffs_combine(const int64_t* idx_a, const double* diff_a, uint64_t n_a, const int64_t* idx_b,
                      const double* diff_b, uint64_t n_b, int32_t op, int32_t drop_leading_zero,
                      uint32_t* out_pos, double* out_val, uint64_t* n_out, void* stream);

/* ---- S2-S5 + T1/T2 in one call: get_fragment_pileup incl. the direct pileup
 *      (sample/__init__.py:5-9, sample/sparsegraphpileup.py:37-250,
 *       sparsediffs.py:137-141,180-190) ---------------------------------------------
 * Same results as gpc_sample_events followed by gpc_events_to_runs, without the
 * caller holding the event list.  mode 0 picks by density: read sets with at least
 * ~0.1 steps per base pair of the graph go through a difference array over the
 * graph's positions (64-bit reductions straight from the walk, then ONE streaming
 * scan that emits both run-length tracks and leaves the array zeroed for the next
 * call; no sort); sparser ones (whole-genome depth) through the event list and the
 * radix sort.  mode 1 / 2 force the array / the event list.
 * Capacity of each of the four output arrays: `capacity` entries;
 * GPC_ERR_CAPACITY reports the needed size in *n_frag / *n_direct (call again).
 * *n_events: +1/-1 steps deposited. */
int gpc_sample_pileup(const gpc_graph_t* graph, const gpc_reads_t* reads,
                      const uint32_t* read_index, uint64_t n_reads, int32_t extension,
                      uint32_t* frag_idx, int32_t* frag_val, uint32_t* direct_idx,
                      int32_t* direct_val, uint64_t capacity, uint64_t* n_frag,
                      uint64_t* n_direct, uint64_t* n_events, uint8_t* touched, int32_t mode,
                      void* stream);

/* ---- C2/C3  LinearMap.map_interval_collection (control/linearmap.py:51-74) and
 *            SparseControl.create up to the LinearPileup
 *            (control/controlgenerator.py:21-41, control/linearpileup.py:130-143)
 * Projects the start position of every control read onto the linear axis
 * (reference float64 operation order), builds the +-1 events of each
 * extension window (half width extensions[w] / 2, weight fragment_length /
 * (2 * half)), and writes the linear background
 *     max(min_value, max_w count_w * weight_w)
 * as breakpoints (out_pos, float64 linear coordinate) and values (out_val).
 * Capacity of out_pos / out_val: 2 * n_extensions * n_reads.  x_out (may be
 * NULL): the projected starts, n_reads doubles.  graph->lin_start/lin_end
 * must be set. */
int gpc_control_linear_track(const gpc_graph_t* graph, const int32_t* start_node,
                             const int32_t* start_off, const uint32_t* read_index,
                             uint64_t n_reads, const int32_t* extensions /* host */,
                             int32_t n_extensions, int32_t fragment_length, double min_value,
                             double* out_pos, double* out_val, uint64_t* n_out, double* x_out,
                             void* stream);

/* ---- C4/C5  LinearPileup.to_sparse_pileup -> LinearMap.to_sparse_pileup
 *            (control/linearpileup.py:56-103, control/linearmap.py:76-99) ----
 * Maps the linear background back onto graph positions: an untouched node
 * gets one entry (node start, min_value); a touched node gets the value in
 * force at its linear start plus every breakpoint inside its linear span,
 * positions through numpy's float floor-division.  Values are multiplied by
 * value_scale (scale_tracks, control/__init__.py:32-42: ratio when < 1).
 * touched may be NULL (all nodes touched).  out_pos is non-decreasing and may
 * repeat a position (the last entry wins).  GPC_ERR_CAPACITY reports the
 * needed capacity in *n_out. */
int gpc_control_backproject(const gpc_graph_t* graph, const uint8_t* touched, const double* bp,
                            const double* val, uint64_t n_bp, double min_value,
                            double value_scale, uint32_t* out_pos, double* out_val,
                            uint64_t out_capacity, uint64_t* n_out, void* stream);

/* ---- P1  PValuesFinder.get_p_values_pileup (sparsepvalues.py:16-31) ---------
 * Merges sample runs and control entries; per distinct position
 * p = -log10 PoissonSF(count * sample_scale, lambda) with the reference's
 * clamp rules (count 0 -> 0, underflow -> 1000); canonical output.
 * sample_scale: 1/ratio when ratio > 1, else 1 (scale_tracks); control_scale
 * multiplies every lambda (ratio when < 1 and not already applied by
 * gpc_control_backproject, else 1).
 * Capacity of out_pos / out_p: n_sample + n_control. */
int gpc_pvalues(const uint32_t* sample_pos, const int32_t* sample_val, uint64_t n_sample,
                double sample_scale, const uint32_t* control_pos, const double* control_val,
                uint64_t n_control, double control_scale, uint32_t* out_pos, double* out_p,
                uint64_t* n_out, void* stream);

/* ---- Q1  PToQValuesMapper.__get_sub_counts + _from_subcounts, first half
 *          (sparsepvalues.py:50-74) ------------------------------------------
 * Distinct non-zero p of one chromosome with the base pairs they cover
 * (unordered).  This table is what ranks exchange in a multi-GPU run.
 * GPC_ERR_CAPACITY reports the needed capacity in *n_distinct. */
int gpc_ptable_local(const uint32_t* pos, const double* p, uint64_t n_runs, uint32_t track_size,
                     double* table_p, uint64_t* table_count, uint64_t table_capacity,
                     uint64_t* n_distinct, void* stream);

/* ---- Q1  PToQValuesMapper._from_subcounts second half + get_p_to_q_values
 *          (sparsepvalues.py:53-59, 95-103) -----------------------------------
 * Input: any concatenation of local tables (p may repeat); total_bp = sum of
 * the track sizes of all chromosomes.  Output: distinct p descending and
 * their q (capacity n_entries each). */
int gpc_qtable_build(const double* table_p, const uint64_t* table_count, uint64_t n_entries,
                     uint64_t total_bp, double* out_p, double* out_q, uint64_t* n_out,
                     void* stream);

/* ---- Q2  QValuesFinder.get_q_values (sparsepvalues.py:116-125) --------------
 * Exact-key lookup of every run's p in the table (p == 0 -> q = 0). */
int gpc_qvalues_apply(const double* p, uint64_t n_runs, const double* table_p,
                      const double* table_q, uint64_t n_table, double* out_q, void* stream);

/* ---- H1  SparseValues.threshold_copy (sparsediffs.py:58-62) -----------------
 * Canonical boolean track of q >= cutoff.  Capacity n_runs. */
int gpc_threshold(const uint32_t* pos, const double* q, uint64_t n_runs, double cutoff,
                  uint32_t* out_pos, uint8_t* out_val, uint64_t* n_out, void* stream);

/* ---- H2-H4  HolesCleaner(graph, track, max_size, touched).run()
 *            (postprocess/holecleaner.py:9-109, segmentanalyzer.py:5-94,
 *             graphs.py:14-267) -------------------------------------------
 * Input: canonical boolean track (pos strictly increasing from 0, val
 * alternating).  Fills every hole (False run between peaks) whose shortest
 * way through the graph is <= max_size (the read length); holes on untouched
 * nodes, on graph sources/sinks, in line-graph components of more than 36
 * pieces or without an entry piece are never filled.  touched: n_nodes bytes
 * or NULL (every node touched).  Output: canonical boolean track; capacity
 * n_runs + 2 is always enough (cleaning only removes runs); GPC_ERR_CAPACITY
 * reports the needed size in *n_out. */
int gpc_clean_holes(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                    uint64_t n_runs, int32_t max_size, const uint8_t* touched, uint32_t* out_pos,
                    uint8_t* out_val, uint64_t out_capacity, uint64_t* n_out, void* stream);

/* ---- M1-M3  SparseMaxPaths(track, graph, score_pileup).run()
 *            (postprocess/maxpaths.py:11-191, graphs.py:270-355,
 *             subgraphanalyzer.py:4-39) and the scoring tail of
 *            CallPeaksFromQvalues.__get_max_paths (callpeaks.py:186-212,
 *            mindense.py:25-59) --------------------------------------------
 * Input: cleaned boolean track; score track (position + values, either
 * float64 in score_val or int32 in score_val_i32, the other NULL: the direct
 * pileup, or the q-values when q_values_max_path); q-value track for
 * the final score (max q over the path).
 * Output, one entry per peak in the reference's order (line-graph components
 * first, then single-node peaks):
 *   peak_start / peak_end   offsets in the first / last node
 *   peak_length             bases covered (obg Interval.length())
 *   peak_score              max q over the path (0 for empty peaks)
 *   peak_info               bit0 is_diff, bit1 is_ambiguous
 *   peak_node_ptr[n+1], peak_nodes   node ids of each path
 *   order                   peak indexes stably sorted by score, descending
 * internal_id_shift: single-node peaks get id (1-based node index +
 * internal_id_shift); min_node-1 is the sane value, -(min_node-1) is what the
 * reference computes (maxpaths.py:34).
 * GPC_ERR_CAPACITY: *n_peaks / *n_nodes_total hold the needed capacities. */
int gpc_max_paths(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                  uint64_t n_runs, const uint32_t* score_pos, const double* score_val,
                  const int32_t* score_val_i32, uint64_t n_score, const uint32_t* q_pos,
                  const double* q_val, uint64_t n_q,
                  int32_t internal_id_shift, int32_t* peak_start, int32_t* peak_end,
                  int32_t* peak_length, double* peak_score, uint8_t* peak_info,
                  uint32_t* peak_node_ptr, uint64_t peak_capacity, int32_t* peak_nodes,
                  uint64_t node_capacity, uint32_t* order, uint64_t* n_peaks,
                  uint64_t* n_nodes_total, void* stream);

/* ---- sub-graphs of the peaks: the second result of SparseMaxPaths.run()
 *          (postprocess/maxpaths.py:96-124, graphs.py:106-116, 157-186, 317-338), which
 *          CallPeaksFromQvalues hands to Reporter.sub_graphs (callpeaks.py:204-210,
 *          reporter.py:11-15) ---------------------------------------------------------
 * Same inputs as gpc_max_paths (cleaned boolean track, score track).  The line graph
 * over the peak pieces -- vertices in the reference's order [tails, whole nodes, heads],
 * an edge i -> j for every graph edge between their nodes -- split into its connected
 * components, numbered as gpc_max_paths numbers its peaks (component c = peak c):
 *   comp_ptr[n_comp + 1]      members of component c: comp_ptr[c] .. comp_ptr[c + 1]
 *   member_node / member_weight / member_exit   node id, int32 piece score and "has an
 *                             edge to the sink" per member, in ascending vertex order
 *   row_ptr[n_members + 1], col   successor lists (CSR) with columns local to the
 *                             component
 * The reference's matrix of component c (k members) is (k + 1) x (k + 1) with
 * M[i, j] = -weight_i for every successor j of i, and M[i, k] = -weight_i for exits.
 * Capacities: comp_capacity >= n_comp + 1, member_capacity >= n_members + 1 (row_ptr;
 * the member arrays take n_members), col_capacity >= n_cols; GPC_ERR_CAPACITY reports the
 * three numbers (call again). */
int gpc_peak_subgraphs(const gpc_graph_t* graph, const uint32_t* pos, const uint8_t* val,
                       uint64_t n_runs, const uint32_t* score_pos, const double* score_val,
                       const int32_t* score_val_i32, uint64_t n_score, uint32_t* comp_ptr,
                       uint64_t comp_capacity, int32_t* member_node, int32_t* member_weight,
                       uint8_t* member_exit, uint32_t* row_ptr, uint64_t member_capacity,
                       uint32_t* col, uint64_t col_capacity, uint64_t* n_comp,
                       uint64_t* n_members, uint64_t* n_cols, void* stream);

/* ---- summits  PeakCollection.cut_around_summit
 *          (peakcollection.py:113-136) with the q-values read from the run-length
 *          track instead of the dense array of analysis_interface.py:388-406 /
 *          mindense.py:25-59 ------------------------------------------------------
 * Peaks as gpc_max_paths returns them (start offset on the first node, end offset
 * on the last node, node ids peak_nodes[peak_node_ptr[i] .. peak_node_ptr[i+1])).
 * Per peak, out[6*i ..] = { summit, length, first, last, start_off, end_off }:
 * summit = the (count // 2)-th smallest offset along the peak among the positions
 * holding the peak's maximum q; the cut [max(0, summit - window), min(summit +
 * window, length)) runs from start_off on node peak_nodes[ptr + first] to end_off
 * on node peak_nodes[ptr + last].  Everything is device memory. */
int gpc_peak_summits(const gpc_graph_t* graph, const int32_t* peak_start,
                     const int32_t* peak_end, const uint32_t* peak_node_ptr,
                     const int32_t* peak_nodes, uint64_t n_peaks, const uint32_t* q_pos,
                     const double* q_val, uint64_t n_q, int32_t window, int32_t* out,
                     void* stream);

/* ---- measurement hooks (no reference counterpart) ---------------------------
 * gpc_launch_count: kernels launched by this library since it was loaded.
 * gpc_profile_enable(1): time every kernel launch with a CUDA event pair on
 * its stream; gpc_profile_report writes "name calls total_ms" lines for the
 * launches since the previous report and returns the bytes needed. */
uint64_t gpc_launch_count(void);
/* Guard mode (GPC_GUARD=1 in the environment when the library is loaded): scratch buffers
 * are poisoned and fenced by canaries; gpc_guard_violations counts buffers found written
 * outside their bounds since the library was loaded. */
int gpc_guard_mode(void);
uint64_t gpc_guard_violations(void);
void gpc_profile_enable(int on);
uint64_t gpc_profile_report(char* buf, uint64_t capacity);

/* ---- ingest: IntervalCollection text file -> reads SoA, on
 *      the HOST ----------------------------------------------------------------------
 * One JSON object per line, as offsetbasedgraph's Interval.to_file_line writes it
 * and the reference reads it back (callpeaks_interface.py:64-66 after vg JSON
 * conversion; reporter.py:25-33 for peaks):
 *     {"start": 3, "end": 17, "region_paths": [12, 13, 15], "direction": 1}
 * keys in any order, unknown keys skipped, blank lines ignored.  A read's start /
 * end node is its first / last region path.  `text` need not be 0-terminated.
 * All pointers are host pointers; `n_threads` <= 0 uses GPC_INGEST_THREADS from the
 * environment if set, else every CPU the process may run on; each worker is placed on
 * its own CPU when it starts.
 * Call once with all output arrays NULL to get the counts, allocate
 * (path_ptr: n_reads + 1), call again with *n_reads / *n_path_nodes set to the
 * capacities.  The counting call keeps its parse and the filling call for the same
 * (unchanged) buffer copies it out, so the text is scanned once; a filling call
 * without a matching counting call parses on its own.  GPC_ERR_ARG on a malformed
 * line (see gpc_last_error_string). */
int gpc_parse_intervals_host(const char* text, uint64_t n_bytes, int32_t n_threads,
                             int32_t* start_node, int32_t* start_off, int32_t* end_node,
                             int32_t* end_off, int64_t* path_ptr, int32_t* path,
                             uint64_t* n_reads, uint64_t* n_path_nodes);

/* ---- ingest: vg alignments as JSON (`vg view -aj`, one Alignment
 *      per line) -> reads SoA, on the HOST ------------------------------------------
 * Replaces pyvg.conversion.vg_json_file_to_interval_collection as the reference
 * calls it (callpeaks_interface.py:24,64, multiplegraphscallpeaks.py:97-98,
 * preprocess_interface.py:44).  Per line: region paths = the mappings' node ids
 * (negated where position.is_reverse), start = offset of the first mapping, end =
 * offset of the last mapping + sum of its edits' from_length (the reference's
 * tests/examples.py:5-31 pins this: offset 10, from_length 4 + 1 -> Interval(10, 15)).
 * Integers may be quoted (protobuf JSON writes 64-bit fields as strings).  Lines
 * without a path or without mappings (unaligned reads) yield no read and are
 * counted in *n_without_path (may be NULL).  When max_node > 0 a node id outside
 * [min_node, max_node] is GPC_ERR_INVALID_INTERVAL (pyvg: IntervalNotInGraphException).
 * Same two-call protocol and threading as gpc_parse_intervals_host. */
int gpc_parse_vg_alignments_host(const char* text, uint64_t n_bytes, int32_t n_threads,
                                 int32_t min_node, int32_t max_node, int32_t* start_node,
                                 int32_t* start_off, int32_t* end_node, int32_t* end_off,
                                 int64_t* path_ptr, int32_t* path, uint64_t* n_reads,
                                 uint64_t* n_path_nodes, uint64_t* n_without_path);

/* ---- ingest: vg graph as JSON (`vg view -Vj`) -> node and edge
 *      lists, on the HOST ------------------------------------------------------------
 * Replaces the parsing half of pyvg.conversion.json_file_to_obg_numpy_graph, which
 * the reference's create_ob_graph calls (preprocess_interface.py:51-57): every line is
 * an object {"node": [{"id", "sequence"}], "edge": [{"from", "to", "from_start",
 * "to_end"}], ...}; node_len = length of the sequence; an edge is reported as
 * (-from if from_start else from, -to if to_end else to), paths and other keys are
 * skipped.  Nodes and edges come out in file order.  One pass over the text either
 * way: with all arrays NULL it counts; with arrays, *n_nodes / *n_edges are their
 * capacities on entry and the numbers found on return -- entries beyond the capacity
 * are counted, not stored, and the call returns GPC_ERR_CAPACITY (a caller that knows
 * an upper bound, e.g. from the file size, needs a single call). */
int gpc_parse_vg_graph_host(const char* text, uint64_t n_bytes, int32_t* node_id,
                            int32_t* node_len, uint64_t* n_nodes, int32_t* edge_from,
                            int32_t* edge_to, uint64_t* n_edges);

/* ---- alignment file -> one file per chromosome, on the HOST ---------------------------
 * split_vg_json_reads_into_chromosomes (preprocess_interface.py:83-145): line i goes,
 * byte for byte, to fds[k] of the first range [range_lo[k], range_hi[k]] holding the
 * first node id of the line (the reference's regular expression, :109), else to
 * fds[n_ranges] ("unmapped").  fds: n_ranges + 1 open descriptors; n_lines[n_ranges + 1]
 * receives the lines written to each. */
int gpc_split_alignments_host(const char* text, uint64_t n_bytes, const int64_t* range_lo,
                              const int64_t* range_hi, uint64_t n_ranges, const int32_t* fds,
                              uint64_t* n_lines);

/* ---- node sequences of a vg graph in JSON form, on the HOST ------------------------
 * For the sequence graph behind the reference's FASTA output
 * (SequenceGraph.set_sequences_using_vg_json_graph at preprocess_interface.py:60-66,
 * PeakFasta at peakfasta.py:8-18): node ids in file order, seq_ptr[n_nodes + 1]
 * offsets into `seq`, the bases as they stand in the file.  One pass and capacities
 * as above (*n_nodes, *n_bases = room on entry, numbers found on return). */
int gpc_parse_vg_sequences_host(const char* text, uint64_t n_bytes, int32_t* node_id,
                                int64_t* seq_ptr, char* seq, uint64_t* n_nodes,
                                uint64_t* n_bases);

/* ---- fragment-length estimation, on the HOST ----------------------
 * The sequential scans of PeakModel (shiftestimation/shiftestimation.py:65-336) for
 * one chromosome: __naive_find_peaks (:262-290) + __naive_peak_pos (:292-336) on both
 * strands, __find_pair_center (:231-260), __model_add_line (:180-218) for both
 * strands.  plus_tags / minus_tags: SORTED read-start positions (Treatment, :26-43).
 * The four int32 lines of 1 + 2 * peaksize + tag_expansion entries are ACCUMULATED
 * into (the caller zeroes them once and calls per chromosome).  centers may be NULL;
 * else *n_centers is its capacity on entry.  *n_centers = pair centres found. */
int gpc_shift_model_host(const int64_t* plus_tags, uint64_t n_plus, const int64_t* minus_tags,
                         uint64_t n_minus, int64_t peaksize, int64_t tag_expansion,
                         int64_t min_tags, int64_t max_tags, int32_t* plus_start,
                         int32_t* plus_end, int32_t* minus_start, int32_t* minus_end,
                         int64_t* centers, uint64_t* n_centers);

/* ---- fragment-length estimation, the whole-array steps on the DEVICE --
 * gpc_node_read_counts: HaploTyper.build (haplotyping.py:18-26).  path: the signed node
 * ids of all read paths back to back (gpc_reads_t.path); counts[v] (n_nodes uint32, zeroed
 * by the call) = forward visits of node index v -- the counter the most-covered path is
 * chosen by. */
int gpc_node_read_counts(const int32_t* path, uint64_t n_path_nodes, int32_t min_node,
                         uint32_t n_nodes, uint32_t* counts, void* stream);
/* gpc_path_start_positions: LinearFilter.find_start_positions (linear_filter.py:22-46).
 * path: n_path 0-based node indexes in path order (device memory), start_offset: where the
 * path starts on its first node.  A read whose start node lies on the path contributes
 * (distance of the node from the path's start - start_offset + start offset) to out_plus
 * when its start node id is positive, and (... + node size - start offset) to out_minus
 * when it is negative; both lists come out in read order (capacity n_reads each). */
int gpc_path_start_positions(const gpc_graph_t* graph, const int32_t* path, uint64_t n_path,
                             int64_t start_offset, const int32_t* start_node,
                             const int32_t* start_off, uint64_t n_reads, int64_t* out_plus,
                             uint64_t* n_plus, int64_t* out_minus, uint64_t* n_minus, void* stream);

/* HaploTyper.get_maximum_interval_through_graph (haplotyping.py:28-58): the path from
 * the graph's single source that always continues on the successor with the highest
 * node_count (first of equals in adjacency order).  0-based node indexes; path may be
 * NULL (count only); *n_path = capacity in, length out.  Nodes without any edge are
 * not sources (tests/test_haplotyping.py:10-25).  A graph with no or several
 * sources is GPC_ERR_ARG ("Only works when graph has one start node", :33). */
int gpc_most_covered_path_host(const int64_t* succ_ptr, const int32_t* succ,
                               const int64_t* pred_ptr, const int64_t* node_count,
                               uint64_t n_nodes, int32_t* path, uint64_t* n_path);

/* ---- C1  LinearMap.from_graph / find_starts / find_ends
 *          (control/linearmap.py:102-154) ------------------------------------
 * HOST function on HOST arrays (one-off preprocessing that the reference
 * caches in *_linear_map.npz): starts[v] = longest distance from the source
 * to the start of node v; ends[v] = linear length - longest distance from
 * the end of v to the sink. */
int gpc_linear_map_from_graph_host(const int64_t* node_len, const int64_t* succ_ptr,
                                   const int32_t* succ, const int64_t* pred_ptr,
                                   const int32_t* pred, uint64_t n_nodes, double* starts,
                                   double* ends);

/* The same on the DEVICE (graph in HBM as for every other stage; starts / ends: n_nodes
 * float64 each, device pointers; *n_barriers (host, may be NULL): number of nodes no edge
 * jumps over).  The recurrence is cut at those barrier nodes: one thread per segment
 * between two barriers solves it locally, a scan of x -> max(x + a, b) functions over the
 * segments carries it across (csrc/linearmap.cu).  Integer arithmetic, bit-identical to
 * the host sweep and to the reference's float64 sums.  Graphs whose ids do not ascend
 * along every edge are GPC_ERR_ARG. */
int gpc_linear_map_from_graph(const gpc_graph_t* graph, double* starts, double* ends,
                              uint64_t* n_barriers, void* stream);

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
