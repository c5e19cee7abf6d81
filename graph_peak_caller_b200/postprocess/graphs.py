"""``SubGraph`` of a peak, as the reference's ``SparseMaxPaths.run()`` returns it next to
every peak (graph_peak_caller/postprocess/graphs.py:106-116, 317-338;
postprocess/maxpaths.py:96-124) and ``Reporter.sub_graphs`` saves it (reporter.py:11-15).

The line graph over the peak pieces and its connected components are made on the device
(``gpc_peak_subgraphs``: the front half of ``gpc_max_paths`` up to the successor lists);
what is left for the host is to wrap the CSR slices of one component as the scipy matrix
the reference's file format holds.  Nothing is computed unless somebody looks at a
sub-graph: the callpeaks pass itself only carries the lazy list."""
from collections.abc import Sequence

import numpy as np


class SubGraph:
    """graphs.py:106-116: the node ids of a line-graph component and its matrix --
    ``(k + 1) x (k + 1)``, the sink last, ``-score`` of the source piece on every edge."""

    def __init__(self, node_ids, graph):
        self._node_ids = node_ids
        self._graph = graph

    def __str__(self):
        return "Subgraph (%s)" % self._node_ids

    __repr__ = __str__


class SubGraphList(Sequence):
    """The sub-graphs of all peaks of one call, in the order of the peaks (line-graph
    components first, then the single-node peaks, whose sub-graph is their region path and
    an empty 1 x 1 matrix, maxpaths.py:120-124)."""

    def __init__(self, fetch, peak_set, index=None, shared=None):
        self._fetch = fetch                  # () -> arrays of kernels.peak_subgraphs
        self._ps = peak_set
        self._index = None if index is None else np.asarray(index)
        self._shared = shared if shared is not None else {}

    def _arrays(self):
        if "arrays" not in self._shared:
            self._shared["arrays"] = self._fetch()
        return self._shared["arrays"]

    def __len__(self):
        return self._ps.n_all if self._index is None else len(self._index)

    def _sub_graph(self, k):
        from scipy.sparse import csr_matrix
        comp_ptr, node, weight, is_exit, row_ptr, col = self._arrays()
        n_comp = comp_ptr.size - 1
        if k >= n_comp:                      # single-node peak
            lo, hi = int(self._ps.node_ptr[k]), int(self._ps.node_ptr[k + 1])
            return SubGraph(self._ps.nodes[lo:hi].tolist(), csr_matrix(([], ([], [])), shape=(1, 1)))
        lo, hi = int(comp_ptr[k]), int(comp_ptr[k + 1])
        size = hi - lo
        deg = (row_ptr[lo + 1:hi + 1] - row_ptr[lo:hi]).astype(np.int64)
        to_sink = is_exit[lo:hi].astype(np.int64)
        indptr = np.r_[0, np.cumsum(deg + to_sink), 0]
        indptr[-1] = indptr[-2]                              # the sink's own (empty) row
        indices = np.empty(int(indptr[-1]), dtype=np.int32)
        data = np.empty(indices.size, dtype=np.int32)
        w = -weight[lo:hi].astype(np.int32)
        a0 = int(row_ptr[lo])
        for i in range(size):                # rows are short (the out-degree of a node)
            b, d = int(indptr[i]), int(deg[i])
            s = int(row_ptr[lo + i]) - a0
            indices[b:b + d] = col[a0 + s:a0 + s + d]
            if to_sink[i]:
                indices[b + d] = size
            data[b:int(indptr[i + 1])] = w[i]
        m = csr_matrix((data, indices, indptr.astype(np.int32)), shape=(size + 1, size + 1))
        return SubGraph(node[lo:hi].astype(np.int64), m)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("sub-graph index out of range")
        return self._sub_graph(i if self._index is None else int(self._index[i]))

    def subset(self, index):
        """The sub-graphs of the peaks ``index`` (positions in the full output)."""
        return SubGraphList(self._fetch, self._ps, index, self._shared)
