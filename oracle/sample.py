"""ORACLE (test infrastructure, CPU) -- sample / fragment pileup.

Restates graph_peak_caller/sample/sparsegraphpileup.py and
graph_peak_caller/intervals.py of the reference on flat arrays.  The algorithm
is the reference's own (per-read bookkeeping, then one sweep over the nodes in
topological order with a {read id -> remaining} dict per node), not the
read-centric form the CUDA kernels use -- that equivalence is exactly what the
parity tests check.

``graph`` is any object exposing node_len, node_off, min_node, succ_ptr, succ,
pred_ptr, pred (numpy arrays, 0-based node indexes); ``reads`` any object
exposing start_node, start_off, end_node, end_off, path_ptr, path.
"""
from collections import defaultdict

import numpy as np

from . import tracks


class InvalidPileupInterval(Exception):
    """reference custom_exceptions.py:14-15"""


def unique_read_mask(reads):
    """UniqueIntervals.__iter__ (intervals.py:23-33): keep the first read per
    start position (signed node id, offset); obg Interval.hash(ignore_end_pos=True)
    is pinned to that key by tests/test_filter_duplicates.py."""
    seen = set()
    keep = np.zeros(len(reads.start_node), dtype=bool)
    for i, key in enumerate(zip(reads.start_node.tolist(), reads.start_off.tolist())):
        if key in seen:
            continue
        seen.add(key)
        keep[i] = True
    return keep


class _Pileup:
    """SparseGraphPileup (sparsegraphpileup.py:25-30)."""

    def __init__(self, n_nodes):
        self.starts = []
        self.ends = []
        self.node_starts = np.zeros(n_nodes + 2)
        self.touched = np.zeros(n_nodes + 2, dtype=bool)


def _add_reads(graph, reads, pile):
    """ReadsAdderWDirect.add_reads (sparsegraphpileup.py:37-119).  Returns the
    per-node read-end offsets for both strands and the direct-pileup closing
    events."""
    n = graph.node_len.size
    off = graph.node_off
    mn = graph.min_node
    fwd_ends = [[] for _ in range(n)]       # _pos_ends, keyed by 0-based node
    rev_ends = [[] for _ in range(n)]       # _neg_ends
    direct_closers = []                     # pos_read_ends  (-1 events)
    direct_openers = []                     # neg_read_ends  (+1 events)
    ptr = reads.path_ptr
    for r in range(len(reads.start_node)):
        rps = np.abs(reads.path[ptr[r]:ptr[r + 1]]).astype(np.int64) - mn
        if rps.size and (rps.min() < 0 or rps.max() >= n):
            raise InvalidPileupInterval(
                "Interval has node(s) not part of graph/pileup: %s"
                % reads.path[ptr[r]:ptr[r + 1]])
        end_node = int(reads.end_node[r])
        start_node = int(reads.start_node[r])
        pile.touched[rps] = True
        if end_node < 0:
            # add_open_neg_interval (:87-92)
            pile.node_starts[1:][rps[1:]] -= 1
            pile.node_starts[rps[:-1]] += 1
        else:
            # add_open_pos_interval (:76-85)
            pile.node_starts[rps[1:]] += 1
            pile.node_starts[1:][rps[:-1]] -= 1
        # _add_start_pos (:69-74)
        s = abs(start_node) - mn
        if start_node > 0:
            pile.starts.append(off[s] + int(reads.start_off[r]))
        else:
            pile.ends.append(off[s + 1] - int(reads.start_off[r]))
        # read end bookkeeping (:49-57, :109-119)
        e = abs(end_node) - mn
        if end_node < 0:
            rev_ends[e].append(int(reads.end_off[r]))
            direct_openers.append(off[e + 1] - int(reads.end_off[r]))
        else:
            fwd_ends[e].append(int(reads.end_off[r]))
            direct_closers.append(off[e] + int(reads.end_off[r]))
    return fwd_ends, rev_ends, direct_closers, direct_openers


def _extend_forward(graph, pile, end_lists, extension):
    """SparseExtender.run_linear (sparsegraphpileup.py:151-184)."""
    n = graph.node_len.size
    off = graph.node_off
    carried = defaultdict(dict)            # node -> {read id -> remaining}
    next_id = 0
    for v in range(n):                     # id order == topological order
        incoming = carried.pop(v, {})
        if incoming:
            pile.touched[v] = True
        size = int(graph.node_len[v])
        pile.node_starts[v] += len(incoming)
        out = {}
        own = end_lists[v]
        candidates = [(next_id + k, e + extension) for k, e in enumerate(own)]
        candidates += list(incoming.items())
        for rid, end in candidates:
            if end <= size:
                pile.ends.append(off[v] + end)
            else:
                out[rid] = end - size
        next_id += len(own)
        pile.node_starts[v + 1] -= len(out)
        for w in graph.succ[graph.succ_ptr[v]:graph.succ_ptr[v + 1]]:
            tgt = carried[int(w)]
            for rid, rem in out.items():     # NodeInfo.update (:14-16)
                tgt[rid] = max(tgt.get(rid, 0), rem)


def _extend_reverse(graph, pile, end_lists, extension):
    """ReverseSparseExtender (sparsegraphpileup.py:190-211): the same sweep on
    the reversed graph, emitting +1 events in forward coordinates."""
    n = graph.node_len.size
    off = graph.node_off
    G = int(off[-1])
    carried = defaultdict(dict)
    next_id = 0
    for v in range(n - 1, -1, -1):
        incoming = carried.pop(v, {})
        if incoming:
            pile.touched[v] = True
        size = int(graph.node_len[v])
        mirrored_start = G - int(off[v + 1])
        pile.node_starts[v + 1] -= len(incoming)      # _add_node_start (:207-208)
        out = {}
        own = end_lists[v]
        candidates = [(next_id + k, e + extension) for k, e in enumerate(own)]
        candidates += list(incoming.items())
        for rid, end in candidates:
            if end <= size:
                pile.starts.append(G - (mirrored_start + end))   # _add_end (:210-211)
            else:
                out[rid] = end - size
        next_id += len(own)
        pile.node_starts[v] += len(out)               # _add_node_end (:204-205)
        for w in graph.pred[graph.pred_ptr[v]:graph.pred_ptr[v + 1]]:
            tgt = carried[int(w)]
            for rid, rem in out.items():
                tgt[rid] = max(tgt.get(rid, 0), rem)


def pileup_to_diffs(pile, node_off):
    """SparseDiffs.from_pileup (sparsediffs.py:180-190).  The index vector has
    n+1 node entries, node_starts has n+2: the last slot is never read."""
    indices = np.r_[pile.starts, pile.ends, node_off]
    diffs = np.r_[np.ones(len(pile.starts)), -np.ones(len(pile.ends)),
                  pile.node_starts]
    order = np.argsort(indices)
    return indices[order], diffs[order]


def fragment_pileup(graph, reads, extension):
    """SamplePileupGenerator.run (sparsegraphpileup.py:214-250) with
    extension = fragment_length - read_length (sample/__init__.py:8).

    Returns a dict:
      sample_diffs  (indices, diffs) raw event list (not canonical)
      sample        (indices, values) canonical run-length fragment pileup
      direct        (indices, values) canonical run-length direct (read-only) pileup
      touched       sorted array of touched node ids
    """
    if extension < 0:
        raise Exception("Invalid extension size %d used. Must be positive. "
                        "Is fragment length < read length?" % extension)
    if extension <= 0:
        raise Exception("Invalid fragment length %d used in SparseExtender" % extension)
    n = graph.node_len.size
    pile = _Pileup(n)
    fwd_ends, rev_ends, closers, openers = _add_reads(graph, reads, pile)

    # get_direct_pileup (:226-237)
    direct = _Pileup(n)
    direct.ends = pile.ends + closers
    direct.starts = pile.starts + openers
    direct.node_starts = pile.node_starts.copy()
    d_idx, d_diffs = pileup_to_diffs(direct, graph.node_off)
    d_idx, d_val = tracks.diffs_to_runs(d_idx, d_diffs)

    _extend_forward(graph, pile, fwd_ends, extension)
    _extend_reverse(graph, pile, rev_ends, extension)
    s_idx, s_diffs = pileup_to_diffs(pile, graph.node_off)
    touched = np.flatnonzero(pile.touched[:-2]) + graph.min_node
    r_idx, r_val = tracks.diffs_to_runs(s_idx, s_diffs)
    return dict(sample_diffs=(s_idx, s_diffs), sample=(r_idx, r_val),
                direct=(d_idx, d_val), touched=touched)
