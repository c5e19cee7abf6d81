"""``SparseControl`` with the reference's contract
(graph_peak_caller/control/controlgenerator.py:9-44): ``create(reads)``
returns the control track as a ``SparseDiffs``.  Two kernel calls
(pipeline.background_track)."""
import logging

from .. import pipeline
from ..device import device_graph
from ..graph import Graph
from ..intervals import Intervals
from ..sample.sparsegraphpileup import touched_mask
from ..sparsediffs import SparseDiffs
from .linearmap import LinearMap


class SparseControl:
    def __init__(self, linear_map, graph, extension_sizes, fragment_length, touched_nodes):
        self._graph = Graph.from_obg(graph)
        self._graph.require_ordered("SparseControl")
        if isinstance(linear_map, LinearMap):
            self._linear_map = linear_map
        else:
            self._linear_map = LinearMap.from_file(linear_map, self._graph)
        self._extension_sizes = list(extension_sizes)
        self._fragment_length = fragment_length
        self._min_value = None
        self._touched_nodes = touched_nodes

    def set_min_value(self, value):
        self._min_value = value

    def create(self, reads):
        if not hasattr(reads, "consume"):
            reads = Intervals(reads)
        dreads, index, n = reads.consume()
        dgraph = device_graph(self._graph, self._linear_map.as_arrays())
        control, linear, min_value = pipeline.background_track(
            dgraph, dreads, index, n, self._extension_sizes, self._fragment_length,
            touched_mask(self._touched_nodes, self._graph), self._min_value)
        self._min_value = min_value
        logging.info("Using min value %s", min_value)
        track = SparseDiffs.from_device_runs(control[0], control[1])
        track.linear_track = linear
        return track
