"""Seeded synthetic variation graphs and reads as the package's
``Graph`` / ``ReadBatch``.

The reference ships no runnable large data set (its MHC blobs are missing), so
every configuration in BASELINE.json is produced by one generator,
``synth_core.py`` (pure numpy on flat arrays, importable without this package so
that the CPU arm of bench.py gets the same inputs); this module wraps its
results.  Nothing here is on the measured path.
"""
from . import synth_core
from .graph import Graph
from .reads import ReadBatch
from .synth_core import (CONFIGS, HG19_BP, config_seed, global_min_of, scaled,  # noqa: F401
                         split_reads)


def make_graph(backbone_bp, site_rate=0.02, seed=0, min_node=1, **kw):
    """Returns a :class:`Graph` whose ids ascend topologically (see
    ``synth_core.graph_arrays`` for the parameters)."""
    node_len, src, dst = synth_core.graph_arrays(backbone_bp, site_rate, seed, **kw)
    return Graph.from_edges(node_len, src, dst, min_node=min_node)


def make_reads(graph, n_reads, read_length=36, fragment_length=200, seed=0, **kw):
    """Returns a :class:`ReadBatch` of ``n_reads`` reads (duplicates included)."""
    return ReadBatch(*synth_core.read_arrays(graph, n_reads, read_length, fragment_length,
                                             seed, **kw))


def chromosome_inputs(cfg, chrom_index, seed_index=None):
    """(Graph, sample ReadBatch, control ReadBatch or None) of one chromosome of a
    named workload (``CONFIGS[name]``, possibly ``scaled``)."""
    return synth_core.chromosome_inputs(cfg, chrom_index, make_graph, make_reads, seed_index)
