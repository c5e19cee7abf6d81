"""Sample (fragment) pileup -- reference graph_peak_caller/sample/."""
import logging

from .sparsegraphpileup import SamplePileupGenerator


def get_fragment_pileup(graph, input_intervals, info, reporter=None):
    """reference sample/__init__.py:5-9"""
    logging.info("Creating fragment pileup, using fragment length %d "
                 "and read length %d" % (info.fragment_length, info.read_length))
    spg = SamplePileupGenerator(graph, info.fragment_length - info.read_length)
    return spg.run(input_intervals, reporter)
