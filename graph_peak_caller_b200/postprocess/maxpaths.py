"""``SparseMaxPaths`` with the reference's contract
(graph_peak_caller/postprocess/maxpaths.py:11-65) over gpc_max_paths.

``run()`` returns ``(peaks, subgraphs)`` like the reference; the sub-graphs are a lazy
``SubGraphList`` over ``gpc_peak_subgraphs`` (postprocess/graphs.py): nothing is computed
for them unless somebody looks (the Reporter does when it writes files)."""
from .. import kernels, pipeline
from ..device import device_graph
from ..graph import Graph
from ..peakcollection import PeakList
from .graphs import SubGraphList

# maxpaths.py:34 subtracts (min_node - 1) from the 1-based node index of a
# single-node peak instead of adding it.  For min_node == 1 both agree; for
# other graphs the reference's ids point outside the graph (and its own
# scoring step then fails), so the default here is the real node id.
REPLICATE_INTERNAL_ID_QUIRK = False


class SparseMaxPaths:
    def __init__(self, sparse_values, graph, score_pileup, variant_maps=None):
        if variant_maps is not None:
            raise NotImplementedError("variant maps are outside this path")
        self._graph = Graph.from_obg(graph)
        self._graph.require_ordered("SparseMaxPaths")
        self._sparse_values = sparse_values
        self._score_pileup = score_pileup
        self._q_values = None

    def set_q_values(self, q_values):
        """Track used for the final score (max q over the path)."""
        self._q_values = q_values

    def _inputs(self):
        torch = kernels.torch_cuda()
        pos, val = self._sparse_values.device_arrays()
        if val.dtype != torch.uint8:
            val = (val != 0).to(torch.uint8)
        spos, sval = self._score_pileup.device_arrays()
        if sval.dtype not in (torch.float64, torch.int32):
            sval = sval.to(torch.float64)
        return pos, val, spos, sval

    def _sub_graph_arrays(self):
        pos, val, spos, sval = self._inputs()
        return kernels.peak_subgraphs(device_graph(self._graph), pos, val, spos, sval)

    def run_raw(self):
        torch = kernels.torch_cuda()
        pos, val, spos, sval = self._inputs()
        q = self._q_values if self._q_values is not None else self._score_pileup
        qpos, qval = q.device_arrays(torch.float64)
        shift = self._graph.min_node - 1
        return kernels.max_paths(device_graph(self._graph), pos, val, spos, sval, qpos, qval,
                                 -shift if REPLICATE_INTERNAL_ID_QUIRK else shift)

    def run(self):
        raw = self.run_raw()
        ps = pipeline.PeakSet(raw, 0)
        self.order = ps.order          # stable order by score, descending (device sort)
        self.peak_set = ps
        peaks = PeakList(ps, self._graph)      # Peak objects are made on access
        return peaks, SubGraphList(self._sub_graph_arrays, ps)
