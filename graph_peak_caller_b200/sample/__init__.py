"""Sample (fragment) pileup: ``get_fragment_pileup`` as ``CallPeaks`` calls it
(reference graph_peak_caller/sample/__init__.py:5-9).  Every read is extended
from its end by ``fragment_length - read_length`` along the graph; the work is
the kernel chain of SamplePileupGenerator.run (gpc_sample_events,
gpc_events_to_runs)."""
import logging

from .sparsegraphpileup import SamplePileupGenerator

__all__ = ["get_fragment_pileup", "SamplePileupGenerator"]


def get_fragment_pileup(graph, input_intervals, info, reporter=None):
    extension = info.fragment_length - info.read_length
    logging.info("Fragment pileup: reads of %d bp extended by %d bp",
                 info.read_length, extension)
    generator = SamplePileupGenerator(graph, extension)
    return generator.run(input_intervals, reporter)
