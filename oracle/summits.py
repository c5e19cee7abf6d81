"""ORACLE (test infrastructure, CPU, numpy) -- peak summits.

Restates PeakCollection.cut_around_summit (graph_peak_caller/peakcollection.py:
113-136 of the reference, v1.2.3) with DensePileup.get_interval_values
(mindense.py:25-59) on flat arrays.  Only tests/ may import this module.

Parity status: PINNED for the summit position -- compared with the unmodified
reference (through oracle/refshim) on random peaks and tracks by
tests/test_summits.py where /root/reference exists, and with the reference's
own known answer (tests/test_command_line_interface.py:121-137: constant q over
Peak(0, 2, [1, 2]) on the 7/4/7/4 graph, window 2 -> Peak(2, 6, [1])).  The cut
itself is obg's Interval.get_subinterval; offsetbasedgraph is absent, its rule
is the one inferred for oracle/refshim and is anchored on that same known answer.
"""
import numpy as np


def dense_track(indices, values, size):
    """SparseValues.to_dense_pileup (sparsediffs.py:64-76): 0 before the first
    index, every value held until the next index."""
    out = np.zeros(size, dtype=np.float64)
    idx = np.r_[np.asarray(indices, dtype=np.int64), size]
    for i, v in enumerate(values):
        out[idx[i]:idx[i + 1]] = v
    return out


def interval_values(node_off, node_len, min_node, dense, start, end, nodes):
    """DensePileup.get_interval_values (mindense.py:25-59), forward intervals."""
    parts = []
    for k, node in enumerate(nodes):
        i = node - min_node
        a = start if k == 0 else 0
        b = end if k == len(nodes) - 1 else int(node_len[i])
        parts.append(dense[node_off[i] + a:node_off[i] + b])
    return np.concatenate(parts)


def summit_position(values):
    """peakcollection.py:115-117"""
    max_positions = np.flatnonzero(values == np.max(values))
    return int(np.partition(max_positions, max_positions.size // 2)[max_positions.size // 2])


def subinterval(node_len, min_node, start, nodes, from_offset, to_offset):
    """obg Interval.get_subinterval: (start offset, end offset, nodes) of the
    part [from_offset, to_offset) along the interval."""
    walked = -start
    first = None
    out = []
    for node in nodes:
        size = int(node_len[node - min_node])
        if first is None and walked + size > from_offset:
            first = from_offset - walked
        if first is not None:
            out.append(node)
        if walked + size >= to_offset:
            return first, to_offset - walked, out
        walked += size
    raise ValueError("sub-interval outside interval")


def cut_around_summit(node_off, node_len, min_node, dense, start, end, nodes, window=60):
    """-> (summit, length, (start offset, end offset, nodes) of the cut);
    peakcollection.py:114-120."""
    values = interval_values(node_off, node_len, min_node, dense, start, end, nodes)
    summit = summit_position(values)
    cut = subinterval(node_len, min_node, start, nodes, max(0, summit - window),
                      min(summit + window, values.size))
    return summit, int(values.size), cut
