"""``SamplePileupGenerator`` with the reference's contract
(graph_peak_caller/sample/sparsegraphpileup.py:214-250): ``run(reads,
reporter)`` returns a ``SparseDiffs`` carrying ``touched_nodes`` and reports
the direct pileup.  The work is three kernel calls (pipeline.sample_pileup)."""
import logging

import numpy as np

from .. import pipeline
from ..device import device_graph
from ..graph import Graph
from ..intervals import Intervals
from ..sparsediffs import SparseDiffs, SparseValues


class TouchedNodes:
    """Set-like view of the touched-node mask (node ids).  The reference builds
    a Python set (sparsegraphpileup.py:247-249); materialising millions of ids
    is only done when somebody iterates."""

    def __init__(self, mask_tensor, min_node):
        self.mask = mask_tensor            # uint8[n_nodes] on the device
        self.min_node = min_node
        self._ids = None

    def ids(self):
        if self._ids is None:
            self._ids = np.flatnonzero(self.mask.cpu().numpy()) + self.min_node
        return self._ids

    def __iter__(self):
        return iter(self.ids().tolist())

    def __len__(self):
        return int(self.ids().size)

    def __bool__(self):
        return len(self) > 0

    def __contains__(self, node_id):
        ids = self.ids()
        i = np.searchsorted(ids, node_id)
        return bool(i < ids.size and ids[i] == node_id)

    def __eq__(self, other):
        return set(self) == set(other)


def touched_mask(touched_nodes, graph):
    """uint8 device mask from whatever the caller holds (TouchedNodes, a set of
    ids, an array of ids, or None == every node)."""
    if touched_nodes is None:
        return None
    if isinstance(touched_nodes, TouchedNodes):
        return touched_nodes.mask
    from ..device import to_device
    mask = np.zeros(graph.n_nodes, dtype=np.uint8)
    if isinstance(touched_nodes, np.ndarray):          # e.g. np.load of the Reporter's file
        ids = touched_nodes.astype(np.int64).ravel()
    else:
        ids = np.fromiter((int(t) for t in touched_nodes), dtype=np.int64)
    mask[ids - graph.min_node] = 1
    return to_device(mask)


class SamplePileupGenerator:
    def __init__(self, graph, extension):
        logging.info("Using extension %d when extending reads. " % extension)
        if extension < 0:
            raise Exception("Invalid extension size %d used. Must be positive. "
                            "Is fragment length < read length?" % extension)
        self._graph = Graph.from_obg(graph)
        self._extension = extension
        self._direct = None

    def get_direct_pileup(self):
        return self._direct

    def run(self, reads, reporter=None):
        if not hasattr(reads, "consume"):
            reads = Intervals(reads)
        dreads, index, n = reads.consume()
        dgraph = device_graph(self._graph)
        frag, direct, touched, _ = pipeline.sample_pileup(dgraph, dreads, index, n,
                                                          self._extension)
        self._direct = SparseValues(direct[0], direct[1])
        self._direct.track_size = self._graph.size_bp
        if reporter is not None:
            reporter.add("direct_pileup", self._direct)
        sdiffs = SparseDiffs.from_device_runs(frag[0], frag[1])
        sdiffs.touched_nodes = TouchedNodes(touched, self._graph.min_node)
        sdiffs.direct_pileup = self._direct
        return sdiffs
