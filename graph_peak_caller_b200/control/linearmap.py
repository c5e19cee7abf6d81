"""``LinearMap`` (reference graph_peak_caller/control/linearmap.py:20-154):
every node gets a [start, end) span on one linear axis: the longest distance
from the graph source to the node start, and the linear length minus the
longest distance from the node end to the sink.

``from_graph`` is the one-off preprocessing step that writes
``*_linear_map.npz``: a longest-path recurrence, run as kernels
(gpc_linear_map_from_graph) when a device is present and as a native sequential sweep on
hosts without one.  The projection and back-projection that use the map are kernels
(gpc_control_linear_track / gpc_control_backproject)."""
import numpy as np

from ..graph import Graph


class LinearMap:
    def __init__(self, node_starts, node_ends, graph):
        self._node_starts = np.asanyarray(node_starts, dtype=np.float64)
        self._node_ends = np.asanyarray(node_ends, dtype=np.float64)
        assert np.all(self._node_starts < self._node_ends), \
            np.where(self._node_starts >= self._node_ends)
        self._length = self._node_ends[-1]
        self._graph = graph

    def __eq__(self, other):
        return bool(np.all(self._node_starts == other._node_starts)
                    and np.all(self._node_ends == other._node_ends))

    def __repr__(self):
        return "%s:%s" % (self._node_starts, self._node_ends)

    def get_node_start(self, node_id):
        return self._node_starts[node_id - self._graph.min_node]

    def get_node_end(self, node_id):
        return self._node_ends[node_id - self._graph.min_node]

    def get_scale_and_offset(self, node_id):
        span = self.get_node_end(node_id) - self.get_node_start(node_id)
        return span / self._graph.node_size(node_id), self.get_node_start(node_id)

    def as_arrays(self):
        return self._node_starts, self._node_ends

    @classmethod
    def from_graph(cls, graph):
        """Longest distance from the source to every node start / from every node end to
        the sink (linearmap.py:102-154).  With a CUDA device: ``gpc_linear_map_from_graph``
        -- the recurrence cut at the nodes no edge jumps over, one thread per segment, a
        scan of max-plus functions across segments (csrc/linearmap.cu).  On a host without
        a device (the reference runs this step once and caches ``*_linear_map.npz``) the
        library's sequential sweep ``from_graph_host`` does it; both are bit-identical."""
        from ..device import cuda_available
        if not cuda_available():
            return cls.from_graph_host(graph)
        import ctypes
        from .. import _lib
        from ..device import device_graph, empty, ptr, stream_ptr, torch_cuda
        torch = torch_cuda()
        g = Graph.from_obg(graph)
        g.require_ordered("LinearMap.from_graph")
        dg = device_graph(g)
        starts, ends = empty(g.n_nodes, torch.float64), empty(g.n_nodes, torch.float64)
        barriers = ctypes.c_uint64(0)
        _lib.check(_lib.load().gpc_linear_map_from_graph(
            ctypes.byref(dg.c), ptr(starts), ptr(ends), ctypes.byref(barriers), stream_ptr()))
        both = torch.stack((starts, ends)).cpu().numpy()
        lm = cls(both[0], both[1], graph)
        lm.n_barriers = int(barriers.value)
        return lm

    @classmethod
    def from_graph_host(cls, graph):
        """The same as a plain sequential sweep over the CSR in id (= topological) order,
        natively on the host (gpc_linear_map_from_graph_host)."""
        import ctypes
        from .. import _lib
        g = Graph.from_obg(graph)
        g.require_ordered("LinearMap.from_graph")
        lib = _lib.load()
        n = g.n_nodes
        starts, ends = np.zeros(n), np.zeros(n)
        node_len = np.ascontiguousarray(g.node_len, dtype=np.int64)
        arrays = [node_len, g.succ_ptr, g.succ, g.pred_ptr, g.pred]
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(lib.gpc_linear_map_from_graph_host(
            p(node_len), p(g.succ_ptr), p(g.succ), p(g.pred_ptr), p(g.pred), n,
            p(starts), p(ends)))
        del arrays
        return cls(starts, ends, graph)

    @classmethod
    def from_file(cls, filename, graph):
        obj = np.load(filename)
        return cls(obj["starts"], obj["ends"], graph)

    def to_file(self, filename):
        np.savez(filename, starts=self._node_starts, ends=self._node_ends)
