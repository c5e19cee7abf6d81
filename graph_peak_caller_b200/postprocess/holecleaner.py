"""``HolesCleaner`` with the reference's contract
(graph_peak_caller/postprocess/holecleaner.py:9-85) over gpc_clean_holes."""
import numpy as np

from .. import kernels
from ..device import device_graph
from ..graph import Graph
from ..sample.sparsegraphpileup import touched_mask
from ..sparsediffs import SparseValues


class HolesCleaner:
    def __init__(self, graph, sparse_values, max_size, touched_nodes=None):
        self._graph = Graph.from_obg(graph)
        self._graph.require_ordered("HolesCleaner")
        self._sparse_values = sparse_values
        self._max_size = max_size
        self._touched_nodes = touched_nodes

    def run(self):
        torch = kernels.torch_cuda()
        pos, val = self._sparse_values.device_arrays()
        if val.dtype != torch.uint8:
            val = (val != 0).to(torch.uint8)
        dgraph = device_graph(self._graph)
        out_pos, out_val = kernels.clean_holes(
            dgraph, pos, val, int(self._max_size),
            touched_mask(self._touched_nodes, self._graph))
        cleaned = SparseValues(out_pos, out_val)
        cleaned.track_size = self._sparse_values.track_size
        return cleaned
