"""TEST INFRASTRUCTURE ONLY (build container) -- run the *unmodified* reference
from /root/reference on flat inputs, through oracle/refshim.

Used by tests/test_oracle_vs_reference.py (oracle pinned against the live
reference) and oracle/make_golden.py (golden vectors committed under
tests/golden/).  /root/reference does not exist on the GPU box; everything
here is skipped there.
"""
import os
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "refshim"))
import refcompat  # noqa: E402


def available():
    return refcompat.available()


def _ref():
    refcompat.install()
    import offsetbasedgraph as obg
    import graph_peak_caller as gpc
    return obg, gpc


def to_ref_graph(graph):
    obg, _ = _ref()
    mn = graph.min_node
    blocks = {mn + i: obg.Block(int(s)) for i, s in enumerate(graph.node_len)}
    adj = {}
    for i in range(graph.node_len.size):
        lo, hi = graph.succ_ptr[i], graph.succ_ptr[i + 1]
        if hi > lo:
            adj[mn + i] = [int(s) + mn for s in graph.succ[lo:hi]]
    return obg.GraphWithReversals(blocks, adj)


def to_ref_intervals(reads, ref_graph):
    obg, _ = _ref()
    out = []
    for i in range(len(reads.start_node)):
        lo, hi = reads.path_ptr[i], reads.path_ptr[i + 1]
        out.append(obg.Interval(
            obg.Position(int(reads.start_node[i]), int(reads.start_off[i])),
            obg.Position(int(reads.end_node[i]), int(reads.end_off[i])),
            [int(p) for p in reads.path[lo:hi]], graph=ref_graph))
    return out


def ref_linear_map(graph):
    _ref()
    from graph_peak_caller.control.linearmap import LinearMap
    lm = LinearMap.from_graph(to_ref_graph(graph))
    return np.asarray(lm._node_starts), np.asarray(lm._node_ends)


def _peaks_to_tuples(peaks):
    return [(int(p.start_position.offset), int(p.end_position.offset),
             [int(r) for r in p.region_paths],
             (bool(p.info[0]), bool(p.info[1])), float(p.score)) for p in peaks]


def ref_callpeaks(graph, sample_reads, control_reads, read_length,
                  fragment_length, has_control=True, global_min=None,
                  q_threshold=0.05, unique=False):
    """CallPeaks.run of the reference; returns every intermediate track."""
    obg, gpc = _ref()
    from graph_peak_caller.control.linearmap import LinearMap
    from graph_peak_caller.intervals import Intervals, UniqueIntervals
    from graph_peak_caller.reporter import Reporter
    from graph_peak_caller.sparsediffs import SparseValues
    rg = to_ref_graph(graph)
    wrap = UniqueIntervals if unique else Intervals
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            LinearMap.from_graph(rg).to_file("lm.npz")
            config = gpc.Configuration()
            config.read_length = read_length
            config.fragment_length = fragment_length
            config.linear_map_name = "lm.npz"
            config.has_control = has_control
            config.global_min = global_min
            config.q_values_threshold = q_threshold
            caller = gpc.CallPeaks(rg, config, Reporter("ref_"))
            s = wrap(to_ref_intervals(sample_reads, rg))
            c = wrap(to_ref_intervals(control_reads, rg))
            caller.run_pre_callpeaks(s, c)
            out = {}
            sv = caller.sample_pileup.get_sparse_values()
            out["sample"] = (sv.indices.copy(), sv.values.copy())
            out["sample_diffs"] = (caller.sample_pileup._indices.copy(),
                                   caller.sample_pileup._diffs.copy())
            out["control_diffs"] = (caller.control_pileup._indices.copy(),
                                    caller.control_pileup._diffs.copy())
            cv = caller.control_pileup.get_sparse_values()
            out["control"] = (cv.indices.copy(), cv.values.copy())
            out["touched"] = np.array(sorted(caller.touched_nodes), dtype=np.int64)
            d = SparseValues.from_sparse_files("ref_direct_pileup")
            out["direct"] = (d.indices, d.values)
            out["n_sample"], out["n_control"] = s.n_reads, c.n_reads
            caller.get_p_values()
            out["p_values"] = (caller.p_values_pileup.indices.copy(),
                               caller.p_values_pileup.values.copy())
            caller.get_p_to_q_values_mapping()
            out["p_to_q"] = dict(caller.p_to_q_values_mapping)
            caller.get_q_values()
            out["q_values"] = (caller.q_values_pileup.indices.copy(),
                               caller.q_values_pileup.values.copy())
            caller.call_peaks_from_q_values()
            h = SparseValues.from_sparse_files("ref_hole_cleaned")
            out["hole_cleaned"] = (h.indices, h.values)
            out["peaks"] = _peaks_to_tuples(caller.max_path_peaks)
            all_paths = obg.IntervalCollection.from_file(
                "ref_all_max_paths.intervalcollection")
            out["all_max_paths"] = [
                (int(p.start_position.offset), int(p.end_position.offset),
                 [int(r) for r in p.region_paths]) for p in all_paths.intervals]
            out["track_size"] = int(rg.node_indexes[-1])
        finally:
            os.chdir(cwd)
    return out


def ref_fragment_pileup(graph, reads, read_length, fragment_length):
    """get_fragment_pileup of the reference (sample/__init__.py:5-9)."""
    _, gpc = _ref()
    from graph_peak_caller.intervals import Intervals
    from graph_peak_caller.sample import get_fragment_pileup
    rg = to_ref_graph(graph)
    config = gpc.Configuration()
    config.read_length = read_length
    config.fragment_length = fragment_length
    sd = get_fragment_pileup(rg, Intervals(to_ref_intervals(reads, rg)), config)
    sv = sd.get_sparse_values()
    return (sv.indices, sv.values,
            np.array(sorted(sd.touched_nodes), dtype=np.int64))


def ref_clean_holes(graph, indices, values, max_size, touched_ids):
    _ref()
    from graph_peak_caller.postprocess.holecleaner import HolesCleaner
    from graph_peak_caller.sparsediffs import SparseValues
    sv = SparseValues(np.array(indices), np.array(values))
    sv.track_size = int(graph.node_off[-1])
    touched = None if touched_ids is None else set(int(t) for t in touched_ids)
    out = HolesCleaner(to_ref_graph(graph), sv, max_size, touched).run()
    return out.indices, out.values


def ref_max_paths(graph, indices, values, score_idx, score_val, track_size,
                  with_subgraphs=False):
    _ref()
    from graph_peak_caller.postprocess.maxpaths import SparseMaxPaths
    from graph_peak_caller.sparsediffs import SparseValues
    sv = SparseValues(np.array(indices), np.array(values))
    sv.track_size = track_size
    sc = SparseValues(np.array(score_idx), np.array(score_val))
    sc.track_size = track_size
    peaks, subs = SparseMaxPaths(sv, to_ref_graph(graph), sc).run()
    out = [(int(p.start_position.offset), int(p.end_position.offset),
            [int(r) for r in p.region_paths],
            (bool(p.info[0]), bool(p.info[1]))) for p in peaks]
    if with_subgraphs:
        return out, [(np.asarray(s._node_ids), s._graph) for s in subs]
    return out
