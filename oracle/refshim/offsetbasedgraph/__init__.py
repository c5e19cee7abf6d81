"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the third-party package
``offsetbasedgraph`` (unpinned in the reference's setup.py:16, source absent
from /root/reference, not installable offline).

It exists so that the *unmodified* reference modules under /root/reference can
be imported in the build container to (a) run the reference's own unit tests
and (b) generate the golden vectors committed under tests/golden/.  Nothing in
the product package imports this file.

Semantics are inferred from the reference's call sites
and pinned by the reference's own tests, which pass on top of this shim (see
oracle/run_reference_tests.sh).
"""
import json
from collections import defaultdict

import numpy as np


class Position:
    __slots__ = ("region_path_id", "offset")

    def __init__(self, region_path_id, offset):
        self.region_path_id = region_path_id
        self.offset = offset

    def __eq__(self, other):
        return (self.region_path_id == other.region_path_id
                and self.offset == other.offset)

    def __hash__(self):
        return hash((int(self.region_path_id), int(self.offset)))

    def __repr__(self):
        return "%s:%s" % (self.region_path_id, self.offset)


class Block:
    def __init__(self, length):
        self._length = length

    def length(self):
        return self._length


class BlockArray:
    """sizes[0] is a dummy 0; node ``node_id_offset + i`` has size sizes[i]."""

    def __init__(self, array, node_id_offset=0):
        self._array = np.asarray(array)
        self.node_id_offset = node_id_offset

    def __len__(self):
        return self._array.size - 1

    def keys(self):
        return range(self.node_id_offset + 1,
                     self.node_id_offset + self._array.size)

    def values(self):
        return [Block(int(s)) for s in self._array[1:]]

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, node_id):
        i = node_id - self.node_id_offset
        return 1 <= i < self._array.size

    def node_size(self, node_id):
        return int(self._array[abs(node_id) - self.node_id_offset])

    def __getitem__(self, node_id):
        return Block(self.node_size(node_id))


class Graph:
    uses_numpy_backend = True

    def __init__(self, blocks, adj_list, create_reverse_adj_list=True,
                 rev_adj_list=None):
        self.blocks = blocks
        self.adj_list = defaultdict(list)
        for k, v in adj_list.items():
            self.adj_list[k] = list(v)
        self.reverse_adj_list = defaultdict(list)
        for a, succs in list(self.adj_list.items()):
            for b in succs:
                self.reverse_adj_list[-b].append(-a)
        if isinstance(blocks, BlockArray):
            self.min_node = blocks.node_id_offset + 1
            self.node_indexes = np.cumsum(blocks._array).astype(np.int64)
            self._sizes = np.asarray(blocks._array[1:], dtype=np.int64)
        else:
            ids = sorted(blocks.keys())
            self.min_node = ids[0]
            n_slots = ids[-1] - ids[0] + 1
            sizes = np.zeros(n_slots, dtype=np.int64)
            for i in ids:
                sizes[i - self.min_node] = blocks[i].length()
            self._sizes = sizes
            self.node_indexes = np.r_[0, np.cumsum(sizes)].astype(np.int64)

    def node_size(self, node_id):
        return int(self._sizes[abs(node_id) - self.min_node])

    def get_sorted_node_ids(self, reverse=False):
        return sorted(self.blocks.keys(), reverse=reverse)

    def get_topological_sorted_node_ids(self):
        ids = list(self.blocks.keys())
        indeg = {i: 0 for i in ids}
        for a in ids:
            for b in self.adj_list[a]:
                indeg[b] += 1
        import heapq
        heap = [i for i in ids if indeg[i] == 0]
        heapq.heapify(heap)
        out = []
        while heap:
            a = heapq.heappop(heap)
            out.append(a)
            for b in self.adj_list[a]:
                indeg[b] -= 1
                if indeg[b] == 0:
                    heapq.heappush(heap, b)
        assert len(out) == len(ids), "graph has a cycle"
        return out

    def get_first_blocks(self):
        # nodes without any edge are not first blocks (reference
        # tests/test_haplotyping.py:10-25,50-66 expects one start with node 11 isolated)
        return [i for i in self.blocks.keys()
                if not self.reverse_adj_list[-i] and (self.adj_list[i] or len(self.blocks) == 1)]

    def convert_to_numpy_backend(self):
        return self

    def __str__(self):
        return "Graph(%d nodes)" % len(self.blocks)


GraphWithReversals = Graph


class Interval:
    def __init__(self, start_position, end_position, region_paths=None,
                 graph=None, direction=None):
        if region_paths is None:
            region_paths = [start_position.region_path_id]
            if end_position.region_path_id != region_paths[0]:
                region_paths.append(end_position.region_path_id)
        region_paths = [int(r) for r in region_paths]
        if not isinstance(start_position, Position):
            start_position = Position(region_paths[0], int(start_position))
        if not isinstance(end_position, Position):
            end_position = Position(region_paths[-1], int(end_position))
        self.start_position = start_position
        self.end_position = end_position
        self.region_paths = region_paths
        self.graph = graph
        self.direction = 1 if direction is None else direction

    def length(self):
        g = self.graph
        total = sum(g.node_size(rp) for rp in self.region_paths)
        return int(total - self.start_position.offset
                   - (g.node_size(self.region_paths[-1])
                      - self.end_position.offset))

    def get_reverse(self):
        g = self.graph
        rps = [-rp for rp in reversed(self.region_paths)]
        start = Position(-self.end_position.region_path_id,
                         g.node_size(self.end_position.region_path_id)
                         - self.end_position.offset)
        end = Position(-self.start_position.region_path_id,
                       g.node_size(self.start_position.region_path_id)
                       - self.start_position.offset)
        return self.__class__(start, end, rps, graph=g)

    def get_subinterval(self, start_offset, end_offset):
        g = self.graph
        # walk along the path; coordinates relative to the interval start
        pos = -self.start_position.offset
        start = end = None
        rps = []
        for rp in self.region_paths:
            size = g.node_size(rp)
            if start is None and pos + size > start_offset:
                start = Position(rp, start_offset - pos)
            if start is not None:
                rps.append(rp)
            if pos + size >= end_offset:
                end = Position(rp, end_offset - pos)
                break
            pos += size
        assert start is not None and end is not None
        return Interval(start, end, rps, graph=g)

    def hash(self, ignore_end_pos=False):
        if ignore_end_pos:
            return hash((int(self.start_position.region_path_id),
                         int(self.start_position.offset)))
        return hash((int(self.start_position.region_path_id),
                     int(self.start_position.offset),
                     int(self.end_position.region_path_id),
                     int(self.end_position.offset),
                     tuple(self.region_paths)))

    def __eq__(self, other):
        return (self.start_position == other.start_position
                and self.end_position == other.end_position
                and list(self.region_paths) == list(other.region_paths))

    def __hash__(self):
        return self.hash()

    def __repr__(self):
        return "Intv(%s, %s, %s)" % (self.start_position, self.end_position,
                                     self.region_paths)

    def to_file_line(self):
        return json.dumps({"start": int(self.start_position.offset),
                           "end": int(self.end_position.offset),
                           "region_paths": [int(r) for r in self.region_paths],
                           "direction": int(self.direction)})

    @classmethod
    def from_file_line(cls, line, graph=None):
        o = json.loads(line)
        return cls(o["start"], o["end"], o["region_paths"],
                   direction=o.get("direction"), graph=graph)


DirectedInterval = Interval


class IntervalCollection:
    interval_class = Interval

    def __init__(self, intervals):
        self.intervals = intervals

    def __iter__(self):
        return iter(self.intervals)

    def to_file(self, file_name, text_file=True):
        with open(file_name, "w") as f:
            for interval in self.intervals:
                f.write(interval.to_file_line() + "\n")

    to_text_file = to_file

    @classmethod
    def from_file(cls, file_name, text_file=True, graph=None):
        with open(file_name) as f:
            intervals = [cls.interval_class.from_file_line(line, graph=graph)
                         for line in f if line.strip()]
        return cls(intervals)

    @classmethod
    def create_list_from_file(cls, file_name, graph=None):
        return cls.from_file(file_name, graph=graph)


class SequenceGraph:
    pass


class NumpyIndexedInterval:
    pass


class IndexedInterval(Interval):
    """Enough of obg's IndexedInterval for haplotyping.py:58 and linear_filter.py:15-41.
    get_offset_at_position is obg's (source absent from the reference tree): stated
    here as the base pair's distance from the interval's start on the forward strand
    -- UNPINNED, see oracle/shift.py."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._distance_to_node = {}
        walked = -self.start_position.offset
        for rp in self.region_paths:
            self._distance_to_node[rp] = walked
            walked += self.graph.node_size(rp)

    def get_offset_at_position(self, position, direction="+"):
        rp = abs(position.region_path_id)
        if direction == "+":
            return int(self._distance_to_node[rp] + position.offset)
        return int(self._distance_to_node[rp] + self.graph.node_size(rp) - position.offset)
