"""ORACLE (test infrastructure, CPU) -- threshold consumers: hole cleaning,
max paths and final peak scoring.

Restates graph_peak_caller/postprocess/{holecleaner,segmentanalyzer,graphs,
maxpaths,subgraphanalyzer}.py, graph_peak_caller/mindense.py:25-59 and
graph_peak_caller/callpeaks.py:150-212 of the reference on flat arrays.  The
graph searches use the same scipy.sparse.csgraph routines as the reference
(third-party, unpinned), so their tie-breaking is inherited, not re-invented.

Positions are global (id-order) coordinates; "node index" below is the
1-based index the reference gets from np.digitize on node_indexes.
"""
import numpy as np
import scipy.sparse.csgraph as csgraph
from scipy.sparse import csr_matrix

from . import tracks


# ---------------------------------------------------------------------------
# segment classification (segmentanalyzer.py)
# ---------------------------------------------------------------------------

def _segment_nodes(segments, node_off):
    """holecleaner.py:50-55 / maxpaths.py:52-57."""
    ids = np.empty_like(segments)
    ids[:, 0] = np.digitize(segments[:, 0], node_off)
    ids[:, 1] = np.digitize(segments[:, 1], node_off, True)
    return ids


def _between(first, last):
    """SegmentAnalyzer._get_spanning_nodes (segmentanalyzer.py:14-18)."""
    out = []
    for a, b in zip(first, last):
        if b > a + 1:
            out.extend(range(a + 1, b))
    return out


def classify_segments(segments, ids, node_off):
    """SegmentAnalyzer.run (segmentanalyzer.py:86-94, 53-84).  Returns
    (to_node_end, whole_nodes, from_node_start, internals):
      to_node_end      2 x k : node index, bases up to the node end   (get_ends)
      whole_nodes      2 x k : node index, node size                  (get_fulls)
      from_node_start  2 x k : node index, bases from the node start  (get_starts)
      internals        k x 2 : segments touching neither node border
    """
    multi = ids[:, 0] != ids[:, 1]
    wholes, tails, heads = [], [], []        # fulls, ends, starts of the reference

    seg, sid = segments[multi], ids[multi]
    wholes.extend(_between(sid[:, 0], sid[:, 1]))
    at_start = seg[:, 0] == node_off[sid[:, 0] - 1]
    at_end = seg[:, 1] == node_off[sid[:, 1]]
    wholes.extend(sid[:, 0][at_start])
    wholes.extend(sid[:, 1][at_end])
    tails.extend(zip(sid[:, 0][~at_start], seg[:, 0][~at_start]))
    heads.extend(zip(sid[:, 1][~at_end], seg[:, 1][~at_end]))

    seg, sid = segments[~multi], ids[~multi]
    at_start = seg[:, 0] == node_off[sid[:, 0] - 1]
    at_end = seg[:, 1] == node_off[sid[:, 1]]
    whole = at_start & at_end
    wholes.extend(sid[:, 0][whole])
    only_start = at_start & ~whole
    heads.extend(zip(sid[:, 1][only_start], seg[:, 1][only_start]))
    only_end = at_end & ~whole
    tails.extend(zip(sid[:, 0][only_end], seg[:, 0][only_end]))
    internals = seg[~at_start & ~at_end]

    def as_pairs(lst):
        a = np.transpose(np.array(lst, dtype="int"))
        return a if a.size else np.empty((2, 0), dtype="int")

    heads = as_pairs(heads)
    if heads.size:
        heads[1] = heads[1] - node_off[heads[0] - 1]
    tails = as_pairs(tails)
    if tails.size:
        tails[1] = node_off[tails[0]] - tails[1]
    w = np.array(wholes, dtype="int")
    if w.size:
        wholes = np.vstack((w, node_off[w] - node_off[w - 1])).astype("int")
    else:
        wholes = np.empty((2, 0), dtype="int")
    return tails, wholes, heads, internals


def split_segments(segments, ids, node_off):
    """SegmentSplitter.run (segmentanalyzer.py:97-134): one piece per node.
    Returns (pieces k x 2, node index per piece, masks dict)."""
    multi = ids[:, 0] != ids[:, 1]
    span_seg, span_ids = segments[multi], ids[multi]
    between = np.array(_between(span_ids[:, 0], span_ids[:, 1]), dtype="int")
    n_old, n_new, n_mid = segments.shape[0], span_seg.shape[0], between.size
    pieces = np.empty((n_old + n_new + n_mid, 2), dtype="int")
    if n_old:
        pieces[:n_old] = segments
        pieces[:n_old, 1][multi] = node_off[span_ids[:, 0]]
    pieces[n_old:n_old + n_new, 0] = node_off[span_ids[:, 1] - 1]
    pieces[n_old:n_old + n_new, 1] = span_seg[:, 1]
    if n_mid:
        pieces[n_old + n_new:, 0] = node_off[between - 1]
        pieces[n_old + n_new:, 1] = node_off[between]
    node = np.r_[ids[:, 0], ids[:, 1][multi], between].astype("int")
    at_start = pieces[:, 0] == node_off[node - 1]
    at_end = pieces[:, 1] == node_off[node]
    whole = at_start & at_end
    masks = dict(whole=whole, head=at_start & ~whole, tail=at_end & ~whole,
                 internal=~at_start & ~at_end)
    return pieces, node, masks


# ---------------------------------------------------------------------------
# line graph over hole / peak pieces (graphs.py)
# ---------------------------------------------------------------------------

class _AlwaysTouched:
    def __contains__(self, item):
        return True


class _PieceGraph:
    """LineGraph (graphs.py:118-186) with StubsFilter (:14-80) or
    PosStubFilter (:83-106).  tails/wholes/heads are 2 x k arrays whose row 0
    holds 1-based node indexes (converted to node ids here) and row 1 the
    weight (bases for holes, score for peaks)."""

    def __init__(self, graph, tails, wholes, heads, positive, last_node=None,
                 touched=None):
        shift = graph.min_node - 1
        self.graph = graph
        self.shift = shift
        if last_node is not None:
            last_node = last_node + shift
        t_id, w_id, h_id = tails[0] + shift, wholes[0] + shift, heads[0] + shift
        touched = _AlwaysTouched() if touched is None else touched

        def succs(v):
            i = int(v) - graph.min_node
            return graph.succ[graph.succ_ptr[i]:graph.succ_ptr[i + 1]] + graph.min_node

        def preds(v):
            i = int(v) - graph.min_node
            return graph.pred[graph.pred_ptr[i]:graph.pred_ptr[i + 1]] + graph.min_node

        self._succs = succs
        t_keep = np.ones(t_id.shape, dtype=bool)
        w_keep = np.ones(w_id.shape, dtype=bool)
        h_keep = np.ones(h_id.shape, dtype=bool)
        if not positive:
            has_pred = lambda ids: np.array([len(preds(v)) for v in ids], dtype=bool)
            has_succ = lambda ids: np.array([len(succs(v)) for v in ids], dtype=bool)
            w_keep &= has_pred(w_id)
            h_keep &= has_pred(h_id)
            t_keep &= has_succ(t_id)
            w_keep &= has_succ(w_id)
        self.t_keep, self.w_keep, self.h_keep = t_keep, w_keep, h_keep
        into = set(w_id.tolist()) | set(h_id.tolist())       # _pos_to_nodes
        outof = set(t_id.tolist()) | set(w_id.tolist())      # _pos_from_nodes
        ft, fw, fh = t_id[t_keep], w_id[w_keep], h_id[h_keep]
        if positive:
            entry = lambda ids: np.array(
                [not any(p in outof for p in preds(v)) for v in ids], dtype=bool)
            leave = lambda ids: np.array(
                [not any(s in into for s in succs(v)) for v in ids], dtype=bool)
        else:
            entry = lambda ids: np.array(
                [not all(p in outof or p not in touched for p in preds(v))
                 for v in ids], dtype=bool)
            leave = lambda ids: np.array(
                [not all(s in into or s > last_node or s not in touched
                         for s in succs(v)) for v in ids], dtype=bool)
        self.whole_entries = np.flatnonzero(entry(fw))
        self.head_entries = np.flatnonzero(entry(fh))
        self.whole_exits = np.flatnonzero(leave(fw))
        self.tail_exits = np.flatnonzero(leave(ft))

        self.forced = (np.vstack((t_id[~t_keep], tails[1][~t_keep])),
                       w_id[~w_keep].astype("int"),
                       np.vstack((h_id[~h_keep], heads[1][~h_keep])))
        self.nodes = np.r_[ft, fw, fh].astype("int")
        self.weights = np.r_[tails[1][t_keep], wholes[1][w_keep],
                             heads[1][h_keep]].astype("int")
        self.n_tails, self.n_heads, self.n = ft.size, fh.size, self.nodes.size
        self.sink = self.n
        self.matrix = self._build()
        self.entries = np.r_[np.arange(self.n_tails),
                             self.n_tails + self.whole_entries,
                             self.n - self.n_heads + self.head_entries].astype("int")

    def _build(self):
        """LineGraph.make_graph (graphs.py:157-186)."""
        nt, nh = self.n_tails, self.n_heads
        target = {int(v): nt + i for i, v in enumerate(self.nodes[nt:])}
        rows, cols, data = [], [], []
        for i, v in enumerate(self.nodes[:self.n - nh]):
            nxt = [target[int(s)] for s in self._succs(v) if int(s) in target]
            rows.extend([i] * len(nxt))
            cols.extend(nxt)
            data.extend([self.weights[i]] * len(nxt))
        to_sink = np.r_[self.tail_exits, nt + self.whole_exits,
                        np.arange(self.n - nh, self.n)].astype("int")
        rows.extend(to_sink.tolist())
        cols.extend([self.sink] * to_sink.size)
        data.extend(self.weights[to_sink])
        return csr_matrix((np.array(data, dtype="int32"),
                           (np.array(rows, dtype="int32"),
                            np.array(cols, dtype="int32"))),
                          [self.sink + 1, self.sink + 1])

    def keep_mask(self, max_size):
        """DividedLinegraph.filter_small (graphs.py:232-267)."""
        if not self.entries.size:
            return np.ones(self.n, dtype=bool)
        is_entry = np.zeros(self.sink, dtype=bool)
        is_entry[self.entries] = True
        n_comp, label = csgraph.connected_components(
            self.matrix[:self.sink, :self.sink])
        keep = np.zeros(self.sink, dtype=bool)
        lookup = np.empty(self.sink + 1, dtype="int")
        # members of every component in ascending order (= np.flatnonzero(label == comp),
        # without one pass over all vertices per component: whole-genome chromosomes have
        # 10^5 components)
        order = np.argsort(label, kind="stable")
        bounds = np.searchsorted(label[order], np.arange(n_comp + 1))
        for comp in range(n_comp):
            members = np.r_[order[bounds[comp]:bounds[comp + 1]], self.sink]
            if members.size - 1 > 36:
                keep[members[:-1]] = True
                continue
            local_entries = np.flatnonzero(is_entry[members[:-1]])
            if not local_entries.size:
                keep[members[:-1]] = True
                continue
            lookup[members] = np.arange(members.size)
            rows = self.matrix[members]
            sub = csr_matrix((rows.data, lookup[rows.indices], rows.indptr),
                             shape=(members.size, members.size))
            dist = csgraph.shortest_path(sub)
            reach = np.min(dist[local_entries], axis=0)
            keep[members[:-1]] = (reach + dist[:, -1])[:-1] > max_size
        return keep

    def masked(self, mask):
        """LineGraph.get_masked (graphs.py:188-206): back to 1-based indexes."""
        nt, nh, n = self.n_tails, self.n_heads, self.n
        nodes = self.nodes - self.shift
        tails = np.vstack((nodes[:nt][mask[:nt]], self.weights[:nt][mask[:nt]]))
        wholes = nodes[nt:n - nh][mask[nt:n - nh]]
        heads = np.vstack((nodes[n - nh:][mask[n - nh:]],
                           self.weights[n - nh:][mask[n - nh:]]))
        f_t, f_w, f_h = self.forced
        f_t = f_t.copy()
        f_h = f_h.copy()
        f_t[0] -= self.shift
        f_h[0] -= self.shift
        return (np.hstack((tails, f_t)), np.r_[wholes, f_w - self.shift],
                np.hstack((heads, f_h)))

    def best_paths(self):
        """PosDividedLineGraph.max_paths / _backtrace (graphs.py:270-355) and
        SubGraphAnalyzer.get_info (subgraphanalyzer.py:4-39)."""
        self.subgraphs = []      # (node ids, csr matrix) per component, graphs.py:336-338
        if not self.entries.size:
            return [], []
        is_entry = np.zeros(self.sink, dtype=bool)
        is_entry[self.entries] = True
        self.matrix.data += 1
        n_comp, label = csgraph.connected_components(
            self.matrix[:self.sink, :self.sink])
        paths, infos = [], []
        order = np.argsort(label, kind="stable")        # members per component, ascending
        bounds = np.searchsorted(label[order], np.arange(n_comp + 1))
        for comp in range(n_comp):
            members = np.r_[order[bounds[comp]:bounds[comp + 1]], self.sink]
            sub = -self.matrix[members][:, members]
            sub.data = sub.data + 1
            self.subgraphs.append((self.nodes[members[:-1]], sub))
            if members.size == 2:
                paths.append([members[0]])
                infos.append((0, 0))
                continue
            dist, pred = csgraph.shortest_path(sub, return_predecessors=True,
                                               method="BF")
            local_entries = np.flatnonzero(is_entry[members[:-1]])
            first = local_entries[np.argmin(dist[local_entries, -1])]
            row = pred[first]
            cur = row[-1]
            walk = []
            while cur > 0:
                walk.append(cur)
                cur = row[cur]
            if (not walk) or walk[-1] != first:
                walk += [first]
            paths.append(members[walk][::-1])
            infos.append(_component_info(sub, walk[::-1], dist))
        return paths, infos


def _component_info(sub, path, dist):
    """SubGraphAnalyzer (subgraphanalyzer.py:4-39): (has_two_bindings, is_ambiguous)."""
    score = np.min(dist[:, -1])
    if len(path) <= 1:
        return (False, False)
    two = np.sum(np.min(sub, 1)) != score
    ambiguous = False
    on_path = set(int(p) for p in path)
    for v in path[:-1]:
        reach = dist[path[0], v]
        row = sub[v]
        for nxt, d in zip(row.indices, row.data):
            if int(nxt) in on_path:
                continue
            if reach + d + dist[nxt, -1] == score:
                ambiguous = True
                break
        if ambiguous:
            break
    return (bool(two), bool(ambiguous))


# ---------------------------------------------------------------------------
# hole cleaning (holecleaner.py)
# ---------------------------------------------------------------------------

def clean_holes(graph, indices, values, max_size, touched_ids=None):
    """HolesCleaner(graph, sparse_values, max_size, touched_nodes).run()
    (holecleaner.py:9-109).  ``indices``/``values`` is the canonical boolean
    thresholded track; returns the cleaned canonical boolean track."""
    indices = np.asarray(indices).astype("int")
    values = np.asarray(values)
    node_off = graph.node_off
    lo = 1 if values[0] != 0 else 0
    hi = indices.size
    last_node = node_off.size
    if values[-1] == 0:
        last_node = np.searchsorted(node_off, indices[-1])
        hi -= 1
    holes = indices[lo:hi].reshape((hi - lo) // 2, 2)
    ids = _segment_nodes(holes, node_off)
    tails, wholes, heads, internals = classify_segments(holes, ids, node_off)
    kept_internals = internals[(internals[:, 1] - internals[:, 0]) > max_size]

    untouched = np.zeros(0, dtype="int")
    touched = None
    if touched_ids is not None:
        touched = set(int(t) for t in touched_ids)
    if touched:
        shift = graph.min_node - 1
        is_touched = np.array([int(v) + shift in touched for v in wholes[0]],
                              dtype=bool)
        untouched = wholes[0][~is_touched].astype("int")
        wholes = wholes[:, is_touched]
    pg = _PieceGraph(graph, tails, wholes, heads, positive=False,
                     last_node=last_node, touched=touched)
    k_tails, k_wholes, k_heads = pg.masked(pg.keep_mask(max_size))

    # build_kept_holes (:87-109)
    k_wholes = np.r_[k_wholes, untouched].astype("int")
    nt, nw, nh = k_tails.shape[1], k_wholes.size, k_heads.shape[1]
    ni = kept_internals.shape[0]
    kept = np.empty((nt + nw + nh + ni, 2), dtype="int")
    if kept.shape == kept_internals.shape:
        kept = kept_internals
    else:
        if nt:
            kept[:nt, 0] = node_off[k_tails[0]] - k_tails[1]
            kept[:nt, 1] = node_off[k_tails[0]]
        if nw:
            kept[nt:nt + nw, 0] = node_off[k_wholes - 1]
            kept[nt:nt + nw, 1] = node_off[k_wholes]
        if k_heads.size:
            kept[nt + nw:nt + nw + nh, 0] = node_off[k_heads[0] - 1]
            kept[nt + nw:nt + nw + nh, 1] = node_off[k_heads[0] - 1] + k_heads[1]
        kept[nt + nw + nh:] = kept_internals
        kept = np.sort(kept, axis=0)          # column-wise, as the reference

    # build_sparse_values (:36-48)
    idx = np.sort(kept.ravel())
    val = np.arange(idx.size) % 2
    if (not idx.size) or idx[0] != 0:
        idx = np.r_[0, idx]
        val = np.r_[1, val]
    if values[-1] == 0:
        idx = np.r_[idx, indices[-1]]
        val = np.r_[val, 0]
    return tracks.canonical_runs(idx.astype("int"), val.astype(bool))


# ---------------------------------------------------------------------------
# max paths (maxpaths.py) and final scoring (callpeaks.py:167-212)
# ---------------------------------------------------------------------------

def piece_scores(pieces, score_idx, score_val, track_size):
    """SparseMaxPaths.get_segment_scores (maxpaths.py:175-191)."""
    at = np.empty_like(pieces)
    at[:, 0] = np.digitize(pieces[:, 0], score_idx)
    at[:, 1] = np.digitize(pieces[:, 1], score_idx, True)
    weighted = np.ediff1d(score_idx, to_end=track_size - score_idx[-1]) * score_val
    prefix = np.r_[0, np.cumsum(weighted)]
    base = prefix[at[:, 1] - 1] - prefix[at[:, 0] - 1]
    partial = (pieces - score_idx[at - 1]) * score_val[at - 1]
    return base + (partial[:, 1] - partial[:, 0])


def max_paths(graph, indices, values, score_idx, score_val, score_track_size,
              replicate_internal_id_quirk=True, with_subgraphs=False):
    """SparseMaxPaths(sparse_values, graph, score_pileup).run() without variant
    maps (maxpaths.py:11-65, 96-163).  Returns a list of peaks, each a tuple
    (start_offset, end_offset, [node ids], (is_diff, is_ambiguous)), in the
    reference's order: line-graph components first, then single-node peaks.

    ``replicate_internal_id_quirk``: maxpaths.py:34 subtracts (min_node-1)
    from the 1-based node index of single-node peaks instead of adding it;
    harmless for min_node == 1.  True follows the reference literally.

    ``with_subgraphs``: also return the second element of the reference's result, one
    SubGraph per peak (graphs.py:106-116, 336-338; maxpaths.py:120-124) as a
    (node ids, scipy csr matrix) pair: the line-graph component of the peak with its sink
    as last row / column and minus the piece score on every edge; a single-node peak has
    its region path and an empty 1 x 1 matrix."""
    indices = np.asarray(indices)
    values = np.asarray(values)
    node_off = graph.node_off
    start = 0 if values[0] == 1 else 1
    edges = indices[start:]
    if values[-1] == 1:
        edges = np.r_[edges, node_off[-1]]
    segments = edges.reshape(edges.size // 2, 2)
    ids = _segment_nodes(segments, node_off)
    pieces, node, masks = split_segments(segments, ids, node_off)
    scores = piece_scores(pieces, np.asarray(score_idx), np.asarray(score_val),
                          score_track_size)

    def scored(mask):
        return np.vstack((node[mask], scores[mask]))

    shift = graph.min_node - 1
    internal_nodes = node[masks["internal"]]
    internal_pos = pieces[masks["internal"]]
    internal_off = internal_pos - node_off[internal_nodes - 1, None]
    if replicate_internal_id_quirk:
        internal_ids = internal_nodes - shift
    else:
        internal_ids = internal_nodes + shift
    internal_peaks = [(int(o[0]), int(o[1]), [int(i)], (False, False))
                      for o, i in zip(internal_off, internal_ids)]

    pg = _PieceGraph(graph, scored(masks["tail"]), scored(masks["whole"]),
                     scored(masks["head"]), positive=True)
    paths, infos = pg.best_paths()
    back = np.concatenate([np.flatnonzero(masks[k])
                           for k in ("tail", "whole", "head")])
    peaks = []
    for path, info in zip(paths, infos):
        pi = back[path]
        nn = node[pi]
        peaks.append((int(pieces[pi[0], 0] - node_off[nn[0] - 1]),
                      int(pieces[pi[-1], 1] - node_off[nn[-1] - 1]),
                      [int(v) for v in nn + shift], info))
    if with_subgraphs:
        small = [(list(p[2]), csr_matrix(([], ([], [])), shape=(1, 1))) for p in internal_peaks]
        return peaks + internal_peaks, pg.subgraphs + small
    return peaks + internal_peaks


def peak_length(graph, peak):
    """obg Interval.length for a forward peak."""
    s, e, nodes = peak[0], peak[1], peak[2]
    sizes = graph.node_len[np.asarray(nodes) - graph.min_node]
    return int(sizes.sum() - s - (sizes[-1] - e))


def score_and_filter_peaks(graph, peaks, q_idx, q_val, fragment_length):
    """CallPeaksFromQvalues.__get_max_paths tail (callpeaks.py:186-212) with
    DensePileup.get_interval_values (mindense.py:25-59): score = max q over
    the path, stable sort by score descending, keep length >= fragment length.
    Returns a list of (start, end, nodes, info, score)."""
    G = int(graph.node_off[-1])
    dense = tracks.runs_to_dense(np.asarray(q_idx), np.asarray(q_val), G)
    scored = []
    for p in peaks:
        length = peak_length(graph, p)
        assert length >= 0
        if length == 0:
            scored.append(p + (0.0,))
            continue
        best = -np.inf
        nodes = p[2]
        for k, v in enumerate(nodes):
            lo = 0
            hi = int(graph.node_len[v - graph.min_node])
            if k == 0:
                lo = p[0]
            if k == len(nodes) - 1:
                hi = p[1]
            assert hi > lo
            base = int(graph.node_off[v - graph.min_node])
            best = max(best, float(np.max(dense[base + lo:base + hi])))
        scored.append(p + (best,))
    scored.sort(key=lambda t: t[4], reverse=True)
    return [t for t in scored if peak_length(graph, t) >= fragment_length]
