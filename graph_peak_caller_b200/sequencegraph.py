"""``SequenceGraph``: the bases of every node, for turning peaks into FASTA records
-- what the reference gets from ``offsetbasedgraph.SequenceGraph`` (absent here):
``create_empty_from_ob_graph`` + ``set_sequences_using_vg_json_graph``
(preprocess_interface.py:60-66), ``from_file`` / ``to_file`` (the ``.sequences``
file next to every graph, callpeaks_interface.py:156-159,277-279) and
``get_interval_sequence`` (peakfasta.py:12,24).

Storage is flat: one byte per base pair of the graph in node-id order, so a node's
sequence is the slice ``[node_off[i], node_off[i + 1])`` -- the same coordinates as
every track of the path.  Bases are kept in lower case.  A node on the reverse
strand (negative id) reads as the reverse complement.  Host code: sequences are
only touched for the few thousand peaks a run reports."""
import numpy as np

from .graph import Graph

_COMPLEMENT = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"acgtn", b"tgcan"):
    _COMPLEMENT[_a] = _b


class SequenceGraph:
    def __init__(self, graph, sequences):
        self._graph = Graph.from_obg(graph)
        self._sequences = np.ascontiguousarray(sequences, dtype=np.uint8)
        if self._sequences.size != self._graph.size_bp:
            raise ValueError("sequence array does not match the graph: %d bases for %d bp"
                             % (self._sequences.size, self._graph.size_bp))

    @classmethod
    def create_empty_from_ob_graph(cls, graph):
        g = Graph.from_obg(graph)
        return cls(g, np.full(g.size_bp, ord("n"), dtype=np.uint8))

    def set_sequence(self, node_id, sequence):
        g = self._graph
        i = node_id - g.min_node
        if len(sequence) != g.node_len[i]:
            raise ValueError("node %d has %d bp, sequence has %d" % (node_id, g.node_len[i],
                                                                      len(sequence)))
        self._sequences[g.node_off[i]:g.node_off[i + 1]] = np.frombuffer(
            sequence.lower().encode("ascii"), dtype=np.uint8)

    def set_sequences_using_vg_json_graph(self, file_name):
        """Every node's sequence from a vg graph in JSON form (native parser
        gpc_parse_vg_sequences_host)."""
        import ctypes
        from . import _lib
        from .reads import _map_file
        lib = _lib.load()
        text = _map_file(file_name)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        # one parse: room for whatever the file could hold (a node object takes 8 bytes or
        # more, the bases are part of the text; untouched pages of the arrays cost nothing)
        n, m = ctypes.c_uint64(text.size // 8 + 1), ctypes.c_uint64(text.size)
        tp = vp(text) if text.size else None
        while True:
            ids = np.empty(int(n.value), dtype=np.int32)
            ptr = np.zeros(int(n.value) + 1, dtype=np.int64)
            seq = np.empty(max(int(m.value), 1), dtype=np.uint8)
            status = lib.gpc_parse_vg_sequences_host(tp, text.size, vp(ids), vp(ptr), vp(seq),
                                                     ctypes.byref(n), ctypes.byref(m))
            if status != _lib.GPC_ERR_CAPACITY:
                break
        _lib.check(status)
        ids, ptr = ids[:int(n.value)], ptr[:int(n.value) + 1]
        g = self._graph
        index = ids.astype(np.int64) - g.min_node
        if index.size and (index.min() < 0 or index.max() >= g.n_nodes):
            raise ValueError("%s names nodes outside the graph" % file_name)
        lens = np.diff(ptr)
        if not np.array_equal(lens, g.node_len[index]):
            raise ValueError("%s: sequence lengths differ from the graph's node sizes" % file_name)
        total = int(ptr[-1])
        if index.size == 0:
            return self
        first, past = int(g.node_off[index[0]]), int(g.node_off[index[-1] + 1])
        if np.all(index[1:] > index[:-1]) and past - first == total:
            # nodes in id order with nothing between them (what vg writes): the bases
            # already stand as the flat array wants them
            np.bitwise_or(seq[:total], 0x20, out=self._sequences[first:past])   # ASCII lower case
            return self
        # any other order: every node's bases to its slot (destination of base k of node j)
        dest = np.repeat(g.node_off[index] - ptr[:-1], lens) + np.arange(total)
        self._sequences[dest] = seq[:total] | 0x20
        return self

    @classmethod
    def from_vg_json_file(cls, file_name, graph):
        return cls.create_empty_from_ob_graph(graph).set_sequences_using_vg_json_graph(file_name)

    def to_file(self, file_name):
        with open(file_name, "wb") as fh:           # a file object: the name stays as given
            np.savez(fh, sequences=self._sequences)

    @classmethod
    def from_file(cls, file_name, graph=None):
        """``graph``: the graph the sequences belong to; by default the graph file whose
        name the sequence file extends (``chr1.nobg.sequences`` -> ``chr1.nobg``)."""
        sequences = np.load(file_name)["sequences"]      # FileNotFoundError before any graph work
        if graph is None:
            from .conversion import load_graph
            graph = load_graph(file_name.rsplit(".sequences", 1)[0])
        return cls(graph, sequences)

    def get_node_sequence(self, node_id):
        return self.get_sequence(node_id)

    def get_sequence(self, node_id, start=0, end=False):
        """Bases [start, end) of a node; for a negative id, of its reverse complement."""
        g = self._graph
        i = abs(node_id) - g.min_node
        bases = self._sequences[g.node_off[i]:g.node_off[i + 1]]
        if node_id < 0:
            bases = _COMPLEMENT[bases[::-1]]
        if end is False:
            end = bases.size
        return bases[start:end].tobytes().decode("ascii")

    def get_interval_sequence(self, interval):
        nodes = interval.region_paths
        parts = []
        for k, node in enumerate(nodes):
            start = interval.start_position.offset if k == 0 else 0
            end = interval.end_position.offset if k == len(nodes) - 1 else False
            parts.append(self.get_sequence(node, start, end))
        return "".join(parts)
