"""Flat (CSR) graph container with the ``offsetbasedgraph`` surface the
callpeaks path reads.

The reference receives an ``offsetbasedgraph.GraphWithReversals``; the only
things its hot path asks of it are ``blocks``, ``adj_list``,
``reverse_adj_list``, ``node_indexes``, ``min_node``, ``node_size`` and a
topological node order (reference sample/sparsegraphpileup.py:131-137,
control/linearmap.py:102-154, postprocess/graphs.py:47-66).  Here those are
views over six flat arrays, which is also exactly what is uploaded to HBM:

    node_len[n]    int64   bases per node (node ``min_node + i``)
    node_off[n+1]  int64   prefix sum of node_len == obg ``node_indexes``
    succ_ptr[n+1]  int64   CSR row pointers, succ[m] int32 0-based targets
    pred_ptr[n+1], pred[m] the transposed CSR

Node ids ascend along every edge in vg / obg graphs; the kernels pop their
per-read frontier in id order, which is then a topological order (reference
sample/sparsegraphpileup.py:136-137 only needs *a* topological order, and its
coordinates are laid out in id order anyway).  A graph built with the obg-style
constructor may break that rule -- cycles as in the reference's
tests/cyclic_graph.py:6-22 -- and is then marked ``is_ordered = False``: the sample
pileup walks such a graph with a label-correcting worklist (the reference's
legacy/extender.py:273-284); every stage that needs the order (linear map, hole
cleaning, max paths) refuses it, as the live reference has no path for it either.
"""
import numpy as np


class Block:
    """obg ``Block``: just a length."""

    def __init__(self, length):
        self._length = int(length)

    def length(self):
        return self._length


class _Blocks:
    """dict-like view: node id -> Block."""

    def __init__(self, graph):
        self._g = graph

    def __len__(self):
        return self._g.n_nodes

    def __contains__(self, node_id):
        return self._g.min_node <= node_id < self._g.min_node + self._g.n_nodes

    def __iter__(self):
        return iter(self.keys())

    def keys(self):
        return range(self._g.min_node, self._g.min_node + self._g.n_nodes)

    def values(self):
        return (Block(s) for s in self._g.node_len)

    def __getitem__(self, node_id):
        if node_id not in self:
            raise KeyError(node_id)
        return Block(self._g.node_len[node_id - self._g.min_node])


class _Adjacency:
    """dict-like view: node id -> list of successor ids (or, reversed,
    ``-id`` -> list of negated predecessor ids, as obg ``reverse_adj_list``)."""

    def __init__(self, graph, reverse):
        self._g = graph
        self._rev = reverse

    def __getitem__(self, node_id):
        g = self._g
        if self._rev:
            i = -node_id - g.min_node
            if not 0 <= i < g.n_nodes:
                return []
            return [-(int(p) + g.min_node)
                    for p in g.pred[g.pred_ptr[i]:g.pred_ptr[i + 1]]]
        i = node_id - g.min_node
        if not 0 <= i < g.n_nodes:
            return []
        return [int(s) + g.min_node
                for s in g.succ[g.succ_ptr[i]:g.succ_ptr[i + 1]]]

    def keys(self):
        g = self._g
        if self._rev:
            deg = np.diff(g.pred_ptr)
            return [-(int(i) + g.min_node) for i in np.flatnonzero(deg)]
        deg = np.diff(g.succ_ptr)
        return [int(i) + g.min_node for i in np.flatnonzero(deg)]

    def items(self):
        return ((k, self[k]) for k in self.keys())

    def __contains__(self, node_id):
        return len(self[node_id]) > 0


class Graph:
    uses_numpy_backend = True

    def __init__(self, blocks, adj_list):
        """obg-style constructor: ``blocks`` maps id -> Block (or int length),
        ``adj_list`` maps id -> iterable of successor ids."""
        ids = sorted(blocks.keys())
        if not ids:
            raise ValueError("empty graph")
        min_node = ids[0]
        n = ids[-1] - ids[0] + 1
        node_len = np.zeros(n, dtype=np.int64)
        for i in ids:
            b = blocks[i]
            node_len[i - min_node] = b.length() if hasattr(b, "length") else int(b)
        src, dst = [], []
        for a, succs in adj_list.items():
            for b in succs:
                src.append(a - min_node)
                dst.append(b - min_node)
        self._init_flat(node_len, np.asarray(src, dtype=np.int64),
                        np.asarray(dst, dtype=np.int64), min_node, allow_unordered=True)

    @classmethod
    def from_edges(cls, node_len, src, dst, min_node=1, allow_unordered=False):
        """``src``/``dst`` are 0-based node indexes."""
        self = cls.__new__(cls)
        self._init_flat(np.asarray(node_len, dtype=np.int64),
                        np.asarray(src, dtype=np.int64),
                        np.asarray(dst, dtype=np.int64), int(min_node), allow_unordered)
        return self

    @classmethod
    def from_obg(cls, g):
        """Convert any object with the obg surface (duck-typed)."""
        if isinstance(g, cls):
            return g
        n = len(g.node_indexes) - 1
        node_len = np.diff(np.asarray(g.node_indexes, dtype=np.int64))
        src, dst = [], []
        for a in g.blocks.keys():
            for b in g.adj_list[a]:
                src.append(a - g.min_node)
                dst.append(b - g.min_node)
        assert node_len.size == n
        return cls.from_edges(node_len, src, dst, g.min_node, allow_unordered=True)

    def require_ordered(self, what):
        if not self.is_ordered:
            raise ValueError(
                "%s needs node ids that ascend along every edge (topological numbering); "
                "cyclic or unsorted graphs are only supported by the sample pileup" % what)

    def _init_flat(self, node_len, src, dst, min_node, allow_unordered=False):
        n = node_len.size
        if src.size and (src.min() < 0 or dst.min() < 0
                         or src.max() >= n or dst.max() >= n):
            raise ValueError("edge endpoint outside the graph")
        self.is_ordered = bool(src.size == 0 or np.all(dst > src))
        if not self.is_ordered and not allow_unordered:
            raise ValueError(
                "node ids must ascend along every edge (topological numbering); "
                "cyclic or unsorted graphs are not supported on this path")
        self.n_nodes = n
        self.min_node = min_node
        self.node_len = node_len
        self.node_off = np.r_[0, np.cumsum(node_len)].astype(np.int64)
        if self.node_off[-1] >= 2 ** 31:
            raise ValueError("graph longer than 2^31 bases")
        order = np.lexsort((dst, src))
        self.succ = dst[order].astype(np.int32)
        self.succ_ptr = np.r_[0, np.cumsum(np.bincount(src, minlength=n))].astype(np.int64)
        order = np.lexsort((src, dst))
        self.pred = src[order].astype(np.int32)
        self.pred_ptr = np.r_[0, np.cumsum(np.bincount(dst, minlength=n))].astype(np.int64)
        self.blocks = _Blocks(self)
        self.adj_list = _Adjacency(self, False)
        self.reverse_adj_list = _Adjacency(self, True)
        self._device = None

    # --- obg surface ---------------------------------------------------------
    @property
    def node_indexes(self):
        return self.node_off

    def node_size(self, node_id):
        return int(self.node_len[abs(node_id) - self.min_node])

    def get_topological_sorted_node_ids(self):
        return range(self.min_node, self.min_node + self.n_nodes)

    def get_sorted_node_ids(self, reverse=False):
        ids = range(self.min_node, self.min_node + self.n_nodes)
        return list(reversed(ids)) if reverse else list(ids)

    def convert_to_numpy_backend(self):
        return self

    @property
    def n_edges(self):
        return int(self.succ.size)

    @property
    def size_bp(self):
        return int(self.node_off[-1])

    def __repr__(self):
        return "Graph(n=%d, m=%d, bp=%d, min_node=%d)" % (
            self.n_nodes, self.n_edges, self.size_bp, self.min_node)

    # --- .npz persistence of the flat form (the CSR the kernels read) --------
    def to_file(self, file_name):
        np.savez(file_name, node_len=self.node_len, min_node=self.min_node,
                 succ_ptr=self.succ_ptr, succ=self.succ)

    @classmethod
    def from_file(cls, file_name):
        o = np.load(file_name)
        src = np.repeat(np.arange(o["node_len"].size), np.diff(o["succ_ptr"]))
        return cls.from_edges(o["node_len"], src, o["succ"], int(o["min_node"]))

    from_numpy_file = from_file


GraphWithReversals = Graph
