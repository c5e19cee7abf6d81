"""vg JSON on the way in -- the two ``pyvg.conversion`` functions the reference
imports on its callpeaks path, by the same names:

* ``vg_json_file_to_interval_collection(file_name, graph)``
  (callpeaks_interface.py:6,24,64; multiplegraphscallpeaks.py:4,97-98;
  preprocess_interface.py:2,44): alignments -> reads;
* ``json_file_to_obg_numpy_graph(file_name, n_nodes=0)``
  (preprocess_interface.py:2,53): graph -> the flat graph of ``graph.py``.

pyvg builds one Python object per alignment, mapping and edit; here both files
go through native host parsers (``gpc_parse_vg_alignments_host``,
``gpc_parse_vg_graph_host``) straight into the arrays that are uploaded.  Also
``split_vg_json_reads_into_chromosomes`` (preprocess_interface.py:83-145), the
step that shards one alignment file by chromosome before the per-chromosome
(here: per-GPU) runs.
"""
import ctypes
import logging

import numpy as np

from .graph import Graph
from .peakcollection import IntervalCollection
from .reads import ReadBatch


def vg_json_file_to_intervals(file_name, graph=None, n_threads=0):
    """The reads of a vg alignment file as a :class:`ReadBatch` (iterating it
    yields obg-style intervals, like the generator pyvg returns)."""
    return ReadBatch.from_vg_json_file(file_name, graph, n_threads=n_threads)


def vg_json_file_to_interval_collection(file_name, graph=None, n_threads=0):
    """``IntervalCollection`` whose ``intervals`` is the flat batch: CallPeaks
    and ``Intervals`` / ``UniqueIntervals`` take it as it is."""
    return IntervalCollection(vg_json_file_to_intervals(file_name, graph, n_threads))


def json_file_to_obg_numpy_graph(file_name, n_nodes=0):
    """vg graph JSON -> :class:`Graph`.  ``n_nodes`` is accepted and unused, as
    in pyvg.  Node ids missing between the smallest and the largest id become
    nodes of length 0 (pyvg's node array has the same holes).  Reversing edges
    (``from_start`` / ``to_end``) and edges that do not ascend in node id are
    rejected: the callpeaks path needs a topologically numbered DAG."""
    from . import _lib
    lib = _lib.load()
    from .reads import _map_file
    text = _map_file(file_name)
    tp = ctypes.c_void_p(text.ctypes.data) if text.size else None
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    # one parse: room for as many entries as the file could hold at all (a node object
    # takes 8 bytes or more, an edge 17; the arrays' untouched pages cost nothing); the
    # call says what it needs should that ever fall short
    n, m = ctypes.c_uint64(text.size // 8 + 1), ctypes.c_uint64(text.size // 17 + 1)
    while True:
        ids, lens = (np.empty(int(n.value), dtype=np.int32) for _ in range(2))
        src, dst = (np.empty(int(m.value), dtype=np.int32) for _ in range(2))
        status = lib.gpc_parse_vg_graph_host(tp, text.size, vp(ids), vp(lens), ctypes.byref(n),
                                             vp(src), vp(dst), ctypes.byref(m))
        if status != _lib.GPC_ERR_CAPACITY:
            break
    _lib.check(status)
    ids, lens = ids[:int(n.value)].copy(), lens[:int(n.value)].copy()
    src, dst = src[:int(m.value)].copy(), dst[:int(m.value)].copy()
    if ids.size == 0:
        raise ValueError("%s holds no nodes" % file_name)
    if src.size and (src.min() < 0 or dst.min() < 0):
        raise ValueError("reversing edges (from_start / to_end) are not supported on this path")
    min_node = int(ids.min())
    node_len = np.zeros(int(ids.max()) - min_node + 1, dtype=np.int64)
    node_len[ids - min_node] = lens
    return Graph.from_edges(node_len, src.astype(np.int64) - min_node,
                            dst.astype(np.int64) - min_node, min_node)


def load_graph(file_name):
    """A graph by file name: vg JSON (``*.json``) or the flat npz container that
    ``Graph.to_file`` / ``callpeaks_interface.create_ob_graph`` write (under any
    name, ``chr1.nobg`` included).  offsetbasedgraph's own binary ``.nobg`` needs
    that package and is refused with a message that says so."""
    import os
    from .custom_exceptions import GraphNotFoundException
    if isinstance(file_name, Graph):
        return file_name
    if not os.path.isfile(file_name):
        raise GraphNotFoundException("Graph file %s not found" % file_name)
    if file_name.endswith(".json"):
        return json_file_to_obg_numpy_graph(file_name)
    try:
        return Graph.from_file(file_name)
    except KeyError as e:
        raise GraphNotFoundException(
            "%s is not a flat graph file (%s): offsetbasedgraph's binary .nobg format is "
            "not read here; create the graph with create_ob_graph from the vg JSON file"
            % (file_name, e)) from None


def split_vg_json_reads_into_chromosomes(chromosomes, reads_file_name, range_files_base_name):
    """Shards an alignment file by chromosome (preprocess_interface.py:83-145):
    a line goes to ``<base>_<chrom>.json`` of the first chromosome whose
    ``node_range_<chrom>.txt`` (``first:last``) holds the first node id found in
    the line, else to ``<base>_unmapped.json`` (a node beyond every range makes the
    reference fail with KeyError on the limits of "unmapped", :102,113-114; here
    the line goes where its log message says).  Returns the number of lines that
    went to ``unmapped``.  One native pass over the mapped file
    (``gpc_split_alignments_host``) instead of a regular expression per line.
    Byte-identical files with the live reference: tests/test_ingest_vg.py."""
    if isinstance(chromosomes, str):
        chromosomes = chromosomes.split(",")
    base = ".".join(reads_file_name.split(".")[0:-1])
    limits = []
    for chrom in chromosomes:
        with open(range_files_base_name + "node_range_" + chrom + ".txt") as fh:
            first, last = fh.read().split(":")[:2]
        limits.append((int(first), int(last)))
    names = list(chromosomes) + ["unmapped"]
    from . import _lib
    from .reads import _map_file
    lib = _lib.load()
    text = _map_file(reads_file_name)
    lo = np.array([a for a, _ in limits], dtype=np.int64)
    hi = np.array([b for _, b in limits], dtype=np.int64)
    counts = np.zeros(len(names), dtype=np.uint64)
    outs = [open(base + "_" + c + ".json", "wb") for c in names]
    try:
        fds = np.array([o.fileno() for o in outs], dtype=np.int32)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.check(lib.gpc_split_alignments_host(vp(text), text.size, vp(lo), vp(hi), len(limits),
                                                 vp(fds), vp(counts)))
    finally:
        for o in outs:
            o.close()
    logging.info("Done. Found %d lines without node id or not matching into the given list of "
                 "chromosomes" % int(counts[-1]))
    return int(counts[-1])


def node_range_of(graph):
    """``first:last`` node id of a graph, the content of a node_range file."""
    return "%d:%d" % (graph.min_node, graph.min_node + graph.n_nodes - 1)


__all__ = ["vg_json_file_to_intervals", "vg_json_file_to_interval_collection",
           "json_file_to_obg_numpy_graph", "load_graph", "split_vg_json_reads_into_chromosomes",
           "node_range_of"]
