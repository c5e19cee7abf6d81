"""ORACLE (test infrastructure, CPU) -- sample pileup on graphs with cycles.

The live path of the reference sweeps the nodes in topological order
(sample/sparsegraphpileup.py:131-137) and has no answer for a graph with cycles; its
legacy path has: ``Extender.get_areas_from_node`` (legacy/extender.py:273-284) over obg's
``GraphTraverser.extend_from_block`` -- a worklist with ``visited[node] = max(remaining)``,
a node covered ``[0, min(remaining, size))`` -- and ``BinaryContinousAreas``
(legacy/areas.py:12-72), where what ONE read covers on a node is a set of positions
(full node, prefix, suffix, internal interval), so overlapping pieces count once.
offsetbasedgraph's traverser is not in the reference tree: this restatement is pinned
on the reference's own known answers, tests/legacy/test_reversal.py:49-70 (self-loop
node of 100 bp, read 70-90, fragment 40 -> prefix 10 and suffix 30; two-node cycle),
and is otherwise PARITY UNPINNED, as SURVEY.md §8(a) row S-cyc states.

Plain Python loops: small cases only.
"""
import numpy as np

from . import tracks


def read_areas(graph, reads, i, extension):
    """{0-based node index: [(a, b), ...]} in the read's direction of travel (offsets
    from the node's start as the read sees it), body pieces and extension pieces apart.
    Returns (body, extension, forward?)."""
    lo, hi = reads.path_ptr[i], reads.path_ptr[i + 1]
    path = [int(p) for p in reads.path[lo:hi]]
    fwd = path[0] > 0
    mn = graph.min_node
    idx = [abs(p) - mn for p in path]
    so, eo = int(reads.start_off[i]), int(reads.end_off[i])
    size = lambda v: int(graph.node_len[v])
    body = []
    for j, v in enumerate(idx):
        a = so if j == 0 else 0
        b = min(eo, size(v)) if j == len(idx) - 1 else size(v)
        if b > a:
            body.append((v, a, b))
    ext = []
    last = idx[-1]
    total = eo + extension
    if min(total, size(last)) > min(eo, size(last)):
        ext.append((last, min(eo, size(last)), min(total, size(last))))
    if fwd:
        nxt = lambda v: graph.succ[graph.succ_ptr[v]:graph.succ_ptr[v + 1]]
    else:
        nxt = lambda v: graph.pred[graph.pred_ptr[v]:graph.pred_ptr[v + 1]]
    visited = {}

    def extend_from_block(v, length):          # obg GraphTraverser.extend_from_block
        if visited.get(v, 0) >= length:
            return
        visited[v] = length
        rem = length - size(v)
        if rem <= 0:
            return
        for w in nxt(v):
            extend_from_block(int(w), rem)

    if total > size(last):
        for w in nxt(last):                     # legacy/extender.py:275-276
            extend_from_block(int(w), total - size(last))
    for v, l in visited.items():                # :278-284 (full node or prefix)
        if l > 0:
            ext.append((v, 0, min(l, size(v))))
    return body, ext, fwd


def _union(pieces):
    out = []
    for v, a, b in sorted(pieces):
        if out and out[-1][0] == v and a <= out[-1][2]:
            out[-1][2] = max(out[-1][2], b)
        else:
            out.append([v, a, b])
    return out


def fragment_pileup(graph, reads, extension):
    """Canonical fragment and direct pileups + touched node ids, like
    oracle.sample.fragment_pileup, for any graph (cycles included)."""
    G = int(graph.node_off[-1])
    starts = {"sample": [], "direct": []}
    ends = {"sample": [], "direct": []}
    touched = set()
    for i in range(len(reads.start_node)):
        body, ext, fwd = read_areas(graph, reads, i, extension)
        for key, pieces in (("sample", body + ext), ("direct", body)):
            for v, a, b in _union(pieces):
                touched.add(v)
                off, n = int(graph.node_off[v]), int(graph.node_len[v])
                if fwd:
                    starts[key].append(off + a)
                    ends[key].append(off + b)
                else:                           # offsets count from the node's END
                    starts[key].append(off + n - b)
                    ends[key].append(off + n - a)
    out = {}
    for key in ("sample", "direct"):
        idx = np.r_[np.array(starts[key], dtype=np.int64), np.array(ends[key], dtype=np.int64)]
        d = np.r_[np.ones(len(starts[key])), -np.ones(len(ends[key]))]
        # an entry at position 0 always exists (SparseValues of a pileup built from
        # node_indexes, sparsediffs.py:180-190)
        idx, d = np.r_[0, idx], np.r_[0.0, d]
        out[key] = tracks.diffs_to_runs(idx, d)
    out["touched"] = np.array(sorted(touched), dtype=np.int64) + graph.min_node
    out["track_size"] = G
    return out
