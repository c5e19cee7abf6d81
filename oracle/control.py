"""ORACLE (test infrastructure, CPU) -- background (lambda) track.

Restates graph_peak_caller/control/{__init__,controlgenerator,linearmap,
linearpileup,linearintervals}.py and graph_peak_caller/eventsorter.py of the
reference on flat arrays, keeping the reference's float64 operation order
(sequential running sums of scaled +-1 steps), so that lambda values agree with
the reference to the last bit where this restatement is pinned.
"""
import numpy as np

from . import tracks


def linear_map_from_graph(graph):
    """LinearMap.from_graph / find_starts / find_ends (linearmap.py:102-154):
    longest distance from the source to each node start, and linear length
    minus longest distance from each node end to the sink."""
    n = graph.node_len.size
    size = graph.node_len
    starts = np.zeros(n)
    for v in range(n):
        reach = starts[v] + size[v]
        for w in graph.succ[graph.succ_ptr[v]:graph.succ_ptr[v + 1]]:
            if reach > starts[w]:
                starts[w] = reach
    back = np.zeros(n)
    for v in range(n - 1, -1, -1):
        reach = back[v] + size[v]
        for w in graph.pred[graph.pred_ptr[v]:graph.pred_ptr[v + 1]]:
            if reach > back[w]:
                back[w] = reach
    length = back[0] + size[0]            # linearmap.py:153: node index 0 is the source
    return starts, length - back


def project_read_starts(graph, lin_starts, lin_ends, reads):
    """LinearMap.map_interval_collection / graph_position_to_linear
    (linearmap.py:51-74) for the start position of every read (the only
    coordinate the control track uses, linearintervals.py:23-25)."""
    node = np.abs(reads.start_node.astype(np.int64)) - graph.min_node
    ns = lin_starts[node]
    ne = lin_ends[node]
    scale = (ne - ns) / graph.node_len[node]
    offset = reads.start_off.astype(np.int64)
    return np.where(reads.start_node > 0, ns + scale * offset, ne - scale * offset)


def linear_background(x, extensions, fragment_length, min_value):
    """SparseControl.create up to the LinearPileup (controlgenerator.py:27-41,
    linearpileup.py:130-143).  Returns (breakpoints float64, values float64)."""
    top = None
    for ext in extensions:
        half = ext // 2
        idx, diffs = tracks.starts_ends_to_diffs(
            np.add.outer(np.array([-half, half]), x))
        diffs = diffs / (half * 2 / fragment_length)
        if top is None:
            top = tracks.clip_min_diffs(idx, diffs, min_value)
        else:
            top = tracks.maximum_diffs(top[0], top[1], idx, diffs)
    idx, diffs = tracks.drop_zero_diffs(*top)
    values = np.cumsum(diffs)
    # sanitize_indices (keep last of equal index), sanitize_values (drop repeats)
    keep = np.r_[np.flatnonzero(np.diff(idx)), values.size - 1]
    idx, values = idx[keep], values[keep]
    keep = np.r_[0, np.flatnonzero(np.diff(values)) + 1]
    return idx[keep], values[keep]


def back_project(graph, lin_starts, lin_ends, bp, val, touched_ids, min_value):
    """LinearPileup.to_sparse_pileup -> get_event_sorter -> from_event_sorter ->
    LinearMap.to_sparse_pileup (linearpileup.py:56-103, eventsorter.py:7-19,
    linearmap.py:76-99).  ``touched_ids`` are node ids (None = every node).
    Returns the control track as (indices int64, diffs float64) -- indices are
    non-decreasing, duplicates possible."""
    mn = graph.min_node
    n = graph.node_len.size
    if touched_ids is None:
        nodes = np.arange(n, dtype=np.int64)
    else:
        nodes = np.asarray(sorted(int(t) for t in touched_ids), dtype=np.int64) - mn
    # event codes: NODE_END 0 < PILEUP_CHANGE 1 < NODE_START 2 at equal index
    ev_idx = np.concatenate([lin_ends[nodes], bp, lin_starts[nodes]])
    ev_val = np.concatenate([nodes.astype(np.float64), val,
                             nodes.astype(np.float64)])
    ev_code = np.concatenate([np.zeros(nodes.size, dtype=np.uint8),
                              np.ones(bp.size, dtype=np.uint8),
                              np.full(nodes.size, 2, dtype=np.uint8)])
    order = np.lexsort((ev_code, ev_idx))
    per_node = {}
    open_nodes = set()
    cur_index, cur_value = 0, 0
    for index, code, value in zip(ev_idx[order], ev_code[order], ev_val[order]):
        if code == 2:
            v = int(value)
            open_nodes.add(v)
            per_node[v] = ([cur_index], [cur_value])
        elif code == 1:
            for v in open_nodes:
                per_node[v][0].append(index)
                per_node[v][1].append(value)
            cur_index, cur_value = index, value
        else:
            open_nodes.remove(int(value))
    off = graph.node_off
    out_idx, out_val = [], []
    for v in range(n):
        if v not in per_node:
            out_idx.append(off[v])
            out_val.append(min_value)
            continue
        idxs, vals = per_node[v]
        scale = (lin_ends[v] - lin_starts[v]) / graph.node_len[v]
        shift = lin_starts[v]
        mapped = [(i - shift) // scale + off[v] for i in idxs]
        mapped[0] = max(off[v], mapped[0])
        out_idx.extend(mapped)
        out_val.extend(vals)
    return (np.array(out_idx, dtype="int"),
            np.diff(np.r_[0, out_val]))


def background_track(graph, lin_starts, lin_ends, reads, extensions,
                     fragment_length, touched_ids, global_min=None):
    """get_background_track (control/__init__.py:7-14) -> SparseControl.create
    (controlgenerator.py:21-44).  Returns dict(control_diffs=(idx, diffs),
    linear=(bp, val), min_value=...)."""
    x = project_read_starts(graph, lin_starts, lin_ends, reads)
    if global_min is not None:
        min_value = global_min
    else:
        min_value = x.size * fragment_length / lin_ends[-1]
    bp, val = linear_background(x, extensions, fragment_length, min_value)
    idx, diffs = back_project(graph, lin_starts, lin_ends, bp, val,
                              touched_ids, min_value)
    return dict(control_diffs=(idx, diffs), linear=(bp, val),
                min_value=min_value, x=x)


def scale_tracks(sample_diffs, control_diffs, ratio):
    """scale_tracks (control/__init__.py:32-42).  Returns new diff arrays."""
    if ratio == 1:
        return sample_diffs, control_diffs
    if ratio > 1:
        return sample_diffs * (1 / ratio), control_diffs
    return sample_diffs, control_diffs * ratio
