"""``Peak`` / ``PeakCollection`` output types of the reference
(graph_peak_caller/peakcollection.py:10-66, 97-98): an interval with a score,
written one JSON object per line."""
import json
from collections.abc import Sequence

from .reads import Interval, Position


class Peak(Interval):
    # peaks built from the kernels' flat output keep the two offsets and create
    # the Position objects on first access (a thousand peaks per chromosome are
    # built per call; most callers only look at a few fields)
    _sp = _ep = None
    _start_off = _end_off = 0

    def __init__(self, start_position, end_position, region_paths=None, graph=None,
                 direction=None, score=0, unique_id=None, chromosome=None):
        super().__init__(start_position, end_position, region_paths, graph, direction)
        self.score = score
        self.unique_id = unique_id
        self.info = (0, 0)
        self.chromosome = chromosome
        self._length = None

    @property
    def start_position(self):
        if self._sp is None:
            p = Position.__new__(Position)
            p.region_path_id = self.region_paths[0]
            p.offset = self._start_off
            self._sp = p
        return self._sp

    @start_position.setter
    def start_position(self, value):
        self._sp = value

    @property
    def end_position(self):
        if self._ep is None:
            p = Position.__new__(Position)
            p.region_path_id = self.region_paths[-1]
            p.offset = self._end_off
            self._ep = p
        return self._ep

    @end_position.setter
    def end_position(self, value):
        self._ep = value

    @classmethod
    def from_flat(cls, start, end, nodes, graph, score, info, length):
        """Constructor for the kernels' output (plain ints already): skips the
        argument normalisation of ``Interval.__init__``."""
        self = object.__new__(cls)
        self.__dict__.update(_start_off=start, _end_off=end, region_paths=nodes, graph=graph,
                             direction=1, score=score, unique_id=None, info=info,
                             chromosome=None, _length=length)
        return self

    def __str__(self):
        return super().__repr__() + " [%s]" % self.score

    __repr__ = __str__

    def length(self):
        if self._length is not None:
            return self._length
        return super().length()

    @classmethod
    def from_interval_and_score(cls, interval, score):
        return cls(interval.start_position, interval.end_position, interval.region_paths,
                   interval.graph, interval.direction, score)

    def set_score(self, score):
        assert isinstance(score, float), "Score %s is invalid, type %s" % (score, type(score))
        self.score = score

    def to_file_line(self):
        return json.dumps({"start": int(self.start_position.offset),
                           "end": int(self.end_position.offset),
                           "region_paths": [int(r) for r in self.region_paths],
                           "direction": int(self.direction),
                           "average_q_value": float(self.score),
                           "unique_id": str(self.unique_id),
                           "is_ambigous": int(self.info[1]),
                           "is_diff": int(self.info[0]),
                           "chromosome": self.chromosome})

    @classmethod
    def from_file_line(cls, line, graph=None):
        o = json.loads(line)
        obj = cls(o["start"], o["end"], o["region_paths"], direction=o["direction"],
                  graph=graph, score=o["average_q_value"], unique_id=o.get("unique_id"))
        if "chromosome" in o:
            obj.chromosome = o["chromosome"]
        obj.info = (o["is_diff"], o["is_ambigous"])
        return obj

    def get_subinterval(self, start_offset, end_offset):
        sub = Interval.get_subinterval(self, start_offset, end_offset)
        peak = Peak(sub.start_position, sub.end_position, sub.region_paths, graph=self.graph,
                    score=self.score, unique_id=self.unique_id)
        peak.info = self.info
        return peak


class PeakList(Sequence):
    """The peaks of one call as a read-only sequence over the kernels' flat output
    (``pipeline.PeakSet``): ``Peak`` objects are made when an element is looked at and
    kept, so whoever holds the list sees the same objects.  A whole-genome run yields
    10^4-10^5 peaks per call; building them all eagerly cost more host time than the whole
    device pass."""

    def __init__(self, peak_set, graph, index=None, chromosome=None, cache=None):
        self._ps = peak_set
        self._graph = graph
        self._index = None if index is None else [int(i) for i in index]
        self._chromosome = chromosome
        self._cache = {} if cache is None else cache

    def __len__(self):
        return self._ps.n_all if self._index is None else len(self._index)

    def _peak(self, k):
        peak = self._cache.get(k)
        if peak is None:
            ps = self._ps
            lo, hi = int(ps.node_ptr[k]), int(ps.node_ptr[k + 1])
            info = int(ps.info[k])
            peak = Peak.from_flat(int(ps.start[k]), int(ps.end[k]), ps.nodes[lo:hi].tolist(),
                                  self._graph, float(ps.score[k]), (bool(info & 1), bool(info & 2)),
                                  int(ps.length[k]))
            if self._chromosome is not None and peak._length != 0:       # callpeaks.py:199-200
                peak.chromosome = self._chromosome
            self._cache[k] = peak
        return peak

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("peak index out of range")
        return self._peak(i if self._index is None else self._index[i])

    def subset(self, index, chromosome=None):
        """The peaks ``index`` (positions in the full output) in that order; shares the
        objects already made."""
        out = PeakList(self._ps, self._graph, index, self._chromosome, self._cache)
        return out

    def set_chromosome(self, chromosome):
        self._chromosome = chromosome
        for peak in self._cache.values():
            if chromosome is not None and peak.length() != 0:
                peak.chromosome = chromosome

    def __eq__(self, other):
        try:
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        except TypeError:
            return NotImplemented

    __hash__ = None

    def __add__(self, other):
        return list(self) + list(other)

    def __radd__(self, other):
        return list(other) + list(self)

    def __repr__(self):
        return "PeakList(%d peaks)" % len(self)


class IntervalCollection:
    """obg ``IntervalCollection``: a list of intervals with JSON-lines I/O."""
    interval_class = Interval

    def __init__(self, intervals):
        self.intervals = intervals

    def __iter__(self):
        return iter(self.intervals)

    def __len__(self):
        return len(self.intervals)

    def to_file(self, file_name, text_file=True):
        with open(file_name, "w") as f:
            for interval in self.intervals:
                f.write(interval.to_file_line() + "\n")

    to_text_file = to_file

    @classmethod
    def from_file(cls, file_name, text_file=True, graph=None):
        with open(file_name) as f:
            return cls([cls.interval_class.from_file_line(line, graph=graph)
                        for line in f if line.strip()])

    create_list_from_file = from_file


class PeakCollection(IntervalCollection):
    interval_class = Peak

    def to_fasta_file(self, file_name, sequence_graph):
        from .peakfasta import PeakFasta
        PeakFasta(sequence_graph).save_intervals(file_name, self)

    @classmethod
    def from_fasta_file(cls, file_name, graph=None):
        """Records as PeakFasta writes them (reference peakcollection.py:258-281): the
        id is the header's first word, the peak its JSON remainder; the sequence is kept
        on the peak."""
        peaks = []
        with open(file_name) as fh:
            while True:
                header, sequence = fh.readline(), fh.readline()
                if not sequence:
                    break
                name, line = header.split(maxsplit=1)
                peak = Peak.from_file_line(line, graph=graph)
                peak.unique_id = name.replace(">", "")
                peak.sequence = sequence.strip()
                peaks.append(peak)
        return cls(peaks)

    def summits(self, q_values, n_base_pairs_around=60, graph=None):
        """int32[n, 6] per peak: summit offset along the peak, peak length, and the
        cut around the summit as (first node index, last node index, start offset,
        end offset) -- kernel gpc_peak_summits over the run-length q-value track."""
        import numpy as np
        from . import kernels
        from .device import device_graph, to_device
        from .graph import Graph
        peaks = list(self.intervals)
        if not peaks:
            return np.zeros((0, 6), dtype=np.int32)
        graph = Graph.from_obg(graph if graph is not None else peaks[0].graph)
        torch = kernels.torch_cuda()
        start = np.array([p.start_position.offset for p in peaks], dtype=np.int32)
        end = np.array([p.end_position.offset for p in peaks], dtype=np.int32)
        ptr = np.zeros(len(peaks) + 1, dtype=np.int32)
        np.cumsum([len(p.region_paths) for p in peaks], out=ptr[1:])
        nodes = np.fromiter((r for p in peaks for r in p.region_paths), dtype=np.int32,
                            count=int(ptr[-1]))
        if nodes.size and nodes.min() < graph.min_node:
            raise ValueError("summits are defined for forward peaks on nodes of the graph")
        q_pos, q_val = q_values.device_arrays(torch.float64)
        out = kernels.peak_summits(device_graph(graph), to_device(start), to_device(end),
                                   to_device(ptr), to_device(nodes), q_pos, q_val,
                                   int(n_base_pairs_around))
        return out.cpu().numpy()

    def cut_around_summit(self, q_values, n_base_pairs_around=60, graph=None):
        """reference peakcollection.py:113-136: every peak is replaced by the part of
        it within ``n_base_pairs_around`` of its summit, the median of the positions
        with the highest q-value.  ``q_values`` is the ``SparseValues`` track of a run
        (``CallPeaks.q_values_pileup`` or ``SparseValues.from_sparse_files(base +
        "qvalues")``); no dense array is built."""
        peaks = list(self.intervals)
        cuts = self.summits(q_values, n_base_pairs_around, graph)
        out = []
        for peak, (summit, length, first, last, start_off, end_off) in zip(peaks, cuts.tolist()):
            cut = Peak(start_off, end_off, peak.region_paths[first:last + 1],
                       graph=peak.graph if peak.graph is not None else graph,
                       score=peak.score, unique_id=peak.unique_id)
            cut.info = peak.info
            cut.chromosome = peak.chromosome
            cut._length = min(summit + n_base_pairs_around, length) \
                - max(0, summit - n_base_pairs_around)
            assert cut._length <= n_base_pairs_around * 2
            out.append(cut)
        self.intervals = out
