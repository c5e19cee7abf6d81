"""Fragment-length estimation when ``-f`` is not given: the
reference's ``shiftestimation`` package and ``linear_filter.py`` under their own
names --

* ``MultiGraphShiftEstimator(position_tuples, genome_size, min_fold, max_fold)``,
  ``.get_estimates()``, ``.from_files(graph_file_names, interval_json_file_names, ...)``
  (shiftestimation/shift_estimation_multigraph.py:6-44);
* ``Treatment``, ``Opt``, ``PeakModel``, ``NotEnoughPairsException``
  (shiftestimation/shiftestimation.py:17-160);
* ``LinearFilter`` with ``find_start_positions`` and ``from_vg_json_reads_and_graph``
  (linear_filter.py:12-60), over ``HaploTyper``'s most-covered path
  (haplotyping.py:18-58).

The whole-array steps -- visits per node over all read paths, the projection of every
read start onto the most-covered path with the split by strand, the sort of the projected
starts -- are kernels (``csrc/shiftdev.cu``: ``gpc_node_read_counts``,
``gpc_path_start_positions``; ``gpc_sort_u64_pairs``) over the reads as they sit in HBM
for the callpeaks pass that follows (one upload serves both).  The sweeps over the sorted
starts and along the graph are native host code (``csrc/shift.cu``:
``gpc_shift_model_host``, ``gpc_most_covered_path_host``): sequential by construction
(every window starts where the one before ended), milliseconds each.  Without a CUDA device
(``DEVICE = False``, or none visible) the whole-array steps are the numpy statements
below -- this is a one-off preprocessing estimate, like ``LinearMap.from_graph``, not the
callpeaks path.  What stays in numpy either way is the cross-correlation of the two
1211-entry model lines, in the same operations as the reference so that ``d`` is the same
float.

offsetbasedgraph's ``IndexedInterval.get_offset_at_position`` is not in the reference
tree; the reference's tests/test_linearfilter.py:33-42 pins it for forward starts
(``distance from the path's start to the node + offset``).  A reverse start
``(-node, offset)`` is taken as the same base pair counted on the forward strand,
``distance + node size - offset``: that test's one reverse start sits mid-node and so
agrees with this without telling it from ``distance + offset`` (unpinned)."""
import ctypes
import logging

import numpy as np

from . import _lib
from .graph import Graph


# None: use the device when one is visible; False: host arrays only (tests compare both)
DEVICE = None


def _use_device():
    if DEVICE is False:
        return False
    from .device import cuda_available
    return cuda_available()


def _device_starts(reads):
    """(start_node, start_off, path or None) of a ReadBatch / a bare namespace as CUDA
    tensors; a ReadBatch is uploaded once for this and for the callpeaks pass."""
    from .device import device_reads, to_device
    if hasattr(reads, "path_ptr") and hasattr(reads, "nbytes"):
        dr = device_reads(reads)
        return dr.start_node, dr.start_off, dr.path
    path = getattr(reads, "path", None)
    return (to_device(np.asarray(reads.start_node, dtype=np.int32)),
            to_device(np.asarray(reads.start_off, dtype=np.int32)),
            None if path is None else to_device(np.asarray(path, dtype=np.int32)))


class NotEnoughPairsException(Exception):
    def __init__(self, value):
        self.value = value

    def __str__(self):
        return repr(self.value)


class Treatment:
    """Sorted read starts per chromosome and strand: ``{"+": {chrom: positions},
    "-": {chrom: positions}}``."""

    def __init__(self, dicts):
        assert isinstance(dicts, dict)
        assert "+" in dicts and "-" in dicts and len(dicts) == 2
        self._chroms = list(set(dicts["+"]) | set(dicts["-"]))
        self._pos_dict = {c: np.sort(np.array(dicts["+"][c], dtype=np.int64)) for c in self._chroms}
        self._neg_dict = {c: np.sort(np.array(dicts["-"][c], dtype=np.int64)) for c in self._chroms}
        self.total = sum(v.size for v in self._pos_dict.values()) + \
            sum(v.size for v in self._neg_dict.values())

    def get_chr_names(self):
        return self._chroms

    def get_locations_by_chr(self, chrom):
        return self._pos_dict[chrom], self._neg_dict[chrom]


class Opt:
    def __init__(self, lmfold=5, umfold=50):
        self.umfold = umfold
        self.lmfold = lmfold
        self.bw = 300
        self.gsize = 300000000
        self.info, self.debug, self.warn = logging.info, logging.debug, logging.warning


def _smooth_flat(x, window_len=11):
    """Moving average over ``window_len`` with reflected ends (the reference's
    ``smooth(x, window="flat")``, shiftestimation.py:339-391)."""
    if x.size < window_len:
        raise ValueError("Input vector needs to be bigger than window size.")
    padded = np.r_[x[window_len - 1:0:-1], x, x[-1:-window_len:-1]]
    weights = np.ones(window_len, "d")
    y = np.convolve(weights / weights.sum(), padded, mode="valid")
    return y[window_len // 2:-(window_len // 2)]


class PeakModel:
    """``d``: the distance between the plus- and the minus-strand read pile of a
    binding site, from the cross-correlation of the two strands' model lines around
    the paired peaks.  Also kept, as in the reference: ``plus_line``, ``minus_line``,
    ``xcorr``, ``ycorr``, ``alternative_d``, ``scan_window``, ``min_tags``, ``max_tags``;
    and ``n_paired_peaks`` / ``paired_peakpos`` ({chrom: centres})."""

    def __init__(self, opt, treatment, max_pairnum=500):
        self.treatment = treatment
        self.gz = opt.gsize
        self.umfold = opt.umfold
        self.lmfold = opt.lmfold
        self.tag_expansion_size = 10
        self.bw = opt.bw
        self.max_pairnum = max_pairnum
        self.build()

    def build(self):
        self.peaksize = 2 * self.bw
        total = float(self.treatment.total)
        # unique hits on a single strand a window must / may hold (:87-91)
        self.min_tags = int(round(total * self.lmfold * self.peaksize / self.gz / 2))
        self.max_tags = int(round(total * self.umfold * self.peaksize / self.gz / 2))
        window = 1 + 2 * self.peaksize + self.tag_expansion_size
        lines = np.zeros((4, window), dtype=np.int32)       # plus start/end, minus start/end
        lib = _lib.load()
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        self.paired_peakpos = {}
        for chrom in sorted(self.treatment.get_chr_names()):
            plus, minus = self.treatment.get_locations_by_chr(chrom)
            n = ctypes.c_uint64(0)
            centers = np.empty(max(plus.size, minus.size, 1) * 4, dtype=np.int64)
            while True:
                n.value = centers.size
                scratch = np.zeros_like(lines)
                status = lib.gpc_shift_model_host(
                    vp(plus), plus.size, vp(minus), minus.size, self.peaksize,
                    self.tag_expansion_size, self.min_tags, self.max_tags, vp(scratch[0]),
                    vp(scratch[1]), vp(scratch[2]), vp(scratch[3]), vp(centers), ctypes.byref(n))
                if status != _lib.GPC_ERR_CAPACITY:
                    break
                centers = np.empty(int(n.value), dtype=np.int64)
            _lib.check(status)
            lines += scratch
            if n.value:
                self.paired_peakpos[chrom] = centers[:int(n.value)].copy()
        self.n_paired_peaks = sum(v.size for v in self.paired_peakpos.values())
        logging.info("#2 number of paired peaks: %d" % self.n_paired_peaks)
        if self.n_paired_peaks < 100:
            logging.warning("Too few paired peaks (%d) to get good shift estimate"
                            % self.n_paired_peaks)
            raise NotEnoughPairsException("Not enough pairs to build model. "
                                          "Try a broader range by setting -m lower and/or -M higher.")
        if self.n_paired_peaks < self.max_pairnum:
            logging.warning("Few paired peaks (%d) when estimating fragment length, which may lead "
                            "to bad estimate of fragment length. Set min fold enrichment (-m lower) "
                            "or max (-M) higher in order get more paired peaks." % self.n_paired_peaks)
        self.plus_line = np.cumsum(lines[0] + lines[1], dtype=np.int32)
        self.minus_line = np.cumsum(lines[2] + lines[3], dtype=np.int32)
        self._cross_correlate(window)

    def _cross_correlate(self, window):
        """shiftestimation.py:148-178, the same numpy operations in the same order."""
        plus, minus = self.plus_line, self.minus_line
        minus_data = (minus - minus.mean()) / (minus.std() * len(minus))
        plus_data = (plus - plus.mean()) / (plus.std() * len(plus))
        ycorr = np.correlate(minus_data, plus_data, mode="full")[
            window - self.peaksize:window + self.peaksize]
        xcorr = np.linspace(len(ycorr) // 2 * -1, len(ycorr) // 2, num=len(ycorr))
        ycorr = _smooth_flat(ycorr)
        rising = np.r_[False, ycorr[1:] > ycorr[:-1]]
        falling = np.r_[ycorr[:-1] > ycorr[1:], False]
        local_max = rising & falling
        positive = xcorr[local_max] > 0
        candidates = ycorr[local_max][positive]
        self.alternative_d = [int(v) for v in xcorr[local_max][positive]]
        assert len(self.alternative_d) > 0, "No proper d can be found! Tweak --mfold?"
        self.d = xcorr[np.where(ycorr == max(candidates))[0][0]]
        self.ycorr, self.xcorr = ycorr, xcorr
        self.scan_window = max(self.d, self.tag_expansion_size) * 2


def most_covered_path(graph, node_counts):
    """0-based node indexes of the path HaploTyper.get_maximum_interval_through_graph
    returns (haplotyping.py:28-58); ``node_counts[i]``: reads through node i."""
    g = Graph.from_obg(graph)
    counts = np.ascontiguousarray(node_counts, dtype=np.int64)
    assert counts.size == g.n_nodes
    lib = _lib.load()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    path = np.empty(g.n_nodes, dtype=np.int32)
    n = ctypes.c_uint64(path.size)
    _lib.check(lib.gpc_most_covered_path_host(vp(g.succ_ptr), vp(g.succ), vp(g.pred_ptr),
                                              vp(counts), g.n_nodes, vp(path), ctypes.byref(n)))
    return path[:int(n.value)]


class LinearFilter:
    """Read starts projected onto one path through the graph (linear_filter.py:12-46).
    ``reads``: a ``ReadBatch``; ``path``: 0-based node indexes in path order."""

    def __init__(self, reads, graph, path=None, start_offset=0):
        if path is None and hasattr(graph, "region_paths"):
            # the reference's form, LinearFilter(position_tuples, indexed_interval):
            # an iterable of obg Positions and an interval through its graph
            import types
            interval = graph
            graph = Graph.from_obg(interval.graph)
            positions = [(int(p.region_path_id), int(p.offset)) for p in reads]
            reads = types.SimpleNamespace(
                start_node=np.array([p[0] for p in positions], dtype=np.int32),
                start_off=np.array([p[1] for p in positions], dtype=np.int32))
            path = np.array(interval.region_paths, dtype=np.int64) - graph.min_node
            start_offset = interval.start_position.offset
        self._reads = reads
        self._graph = Graph.from_obg(graph)
        self._path = np.asarray(path, dtype=np.int64)
        self._start_offset = int(start_offset)      # of the path on its first node

    def length(self):
        return int(self._graph.node_len[self._path].sum())

    def find_start_positions(self, sort=False):
        """{"+": positions, "-": positions} in read order (linear_filter.py:22-46);
        ``sort=True``: ascending, which is what ``Treatment`` makes of them anyway."""
        if _use_device():
            return self._find_start_positions_device(sort)
        found = self._find_start_positions_host()
        return {k: np.sort(v) for k, v in found.items()} if sort else found

    def _find_start_positions_device(self, sort):
        from . import kernels
        from .device import device_graph, to_device
        start_node, start_off, _ = _device_starts(self._reads)
        path = to_device(self._path.astype(np.int32))
        plus, minus = kernels.path_start_positions(device_graph(self._graph), path,
                                                   self._start_offset, start_node, start_off)
        out = {}
        for strand, pos in (("+", plus), ("-", minus)):
            if sort and pos.numel() > 1 and self._start_offset <= 0:
                # (positions >= 0: int64 order is uint64 order; digits equal in all keys
                #  are skipped by the sort)
                pos, _ = kernels.sort_u64_pairs(pos)
                out[strand] = pos.cpu().numpy()
            else:
                out[strand] = np.sort(pos.cpu().numpy()) if sort else pos.cpu().numpy()
        logging.info("Found in total %d positions" % out["+"].size)
        return out

    def _find_start_positions_host(self):
        g, r = self._graph, self._reads
        distance = np.full(g.n_nodes, -1, dtype=np.int64)
        sizes = g.node_len[self._path]
        distance[self._path] = np.cumsum(sizes) - sizes
        on_the_path = distance >= 0
        distance -= self._start_offset
        node = np.abs(r.start_node.astype(np.int64)) - g.min_node
        inside = (node >= 0) & (node < g.n_nodes)
        on_path = np.zeros(node.size, dtype=bool)
        on_path[inside] = on_the_path[node[inside]]
        forward = r.start_node > 0
        offset = r.start_off.astype(np.int64)
        keep = on_path & forward
        plus = distance[node[keep]] + offset[keep]
        keep = on_path & ~forward
        minus = distance[node[keep]] + g.node_len[node[keep]] - offset[keep]
        logging.info("Found in total %d positions" % plus.size)
        return {"+": plus, "-": minus}

    @classmethod
    def from_vg_json_reads_and_graph(cls, json_file_name, graph_file_name):
        from .conversion import load_graph
        from .reads import ReadBatch
        graph = load_graph(graph_file_name)
        if isinstance(json_file_name, ReadBatch):
            reads = json_file_name
        elif json_file_name.endswith(".intervalcollection"):
            reads = ReadBatch.from_intervalcollection_file(json_file_name)
        else:
            reads = ReadBatch.from_vg_json_file(json_file_name)
        return cls.from_reads_and_graph(reads, graph)

    @classmethod
    def from_reads_and_graph(cls, reads, graph):
        """Node counts as HaploTyper.build keeps them (haplotyping.py:18-26: keyed by
        signed id, so only forward visits count on the forward path)."""
        g = Graph.from_obg(graph)
        if _use_device():
            from . import kernels
            counts = kernels.node_read_counts(_device_starts(reads)[2], g.min_node,
                                              g.n_nodes).cpu().numpy()
        else:
            visited = reads.path[reads.path > 0].astype(np.int64) - g.min_node
            visited = visited[(visited >= 0) & (visited < g.n_nodes)]
            counts = np.bincount(visited, minlength=g.n_nodes)
        return cls(reads, g, most_covered_path(g, counts))


class MultiGraphShiftEstimator(object):
    def __init__(self, position_tuples, genome_size, min_fold_enrichment=5,
                 max_fold_enrichment=50):
        self._position_tuples = position_tuples
        self.genome_size = genome_size
        self.min_fold_enrichment = min_fold_enrichment
        self.max_fold_enrichment = max_fold_enrichment

    def get_estimates(self):
        opt = Opt(self.min_fold_enrichment, self.max_fold_enrichment)
        opt.gsize = self.genome_size
        self.peakmodel = PeakModel(opt, Treatment(self._position_tuples))
        return round(self.peakmodel.d)

    @classmethod
    def from_files(cls, graph_file_names, interval_json_file_names, min_fold_enrichment=5,
                   max_fold_enrichment=50):
        """One most-covered path per graph; chromosomes are named "0", "1", ... and the
        genome size is the sum of the paths' lengths (:23-44)."""
        positions = {"+": {}, "-": {}}
        genome_size = 0
        for i, (graph, reads) in enumerate(zip(graph_file_names, interval_json_file_names)):
            linear_filter = LinearFilter.from_vg_json_reads_and_graph(reads, graph)
            found = linear_filter.find_start_positions(sort=True)
            positions["+"][str(i)] = found["+"]
            positions["-"][str(i)] = found["-"]
            genome_size += linear_filter.length()
        logging.info("Using genome size %d" % genome_size)
        return cls(positions, genome_size, min_fold_enrichment, max_fold_enrichment)
