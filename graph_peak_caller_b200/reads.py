"""Reads as a structure of arrays, plus the small obg-style value types
(``Position``, ``Interval``) that the reference API hands around.

The reference consumes an iterable of ``offsetbasedgraph.Interval`` objects
and looks at three things per read: ``region_paths`` (signed node ids, all
negative for a reverse-strand read), ``start_position`` and ``end_position``
(reference sample/sparsegraphpileup.py:47-92, control/linearmap.py:62-74).
``ReadBatch`` stores exactly that for all reads at once:

    start_node[R] int32 (signed id)   start_off[R] int32
    end_node[R]   int32 (signed id)   end_off[R]   int32
    path_ptr[R+1] int64               path[sum P]  int32 (signed ids)

Iterating a ``ReadBatch`` yields ``Interval`` views, so code written against
the reference's iterable-of-intervals contract keeps working, but the product
path never does that: it uploads the arrays as they are.
"""
import json

import os

import numpy as np


def _map_file(file_name):
    """The bytes of a file as a read-only uint8 array without copying them
    (memory-mapped; an empty file gives an empty array)."""
    import os
    if os.path.getsize(file_name) == 0:
        return np.empty(0, dtype=np.uint8)
    return np.memmap(file_name, dtype=np.uint8, mode="r")


class Position:
    __slots__ = ("region_path_id", "offset")

    def __init__(self, region_path_id, offset):
        self.region_path_id = int(region_path_id)
        self.offset = int(offset)

    def __eq__(self, other):
        return (self.region_path_id == other.region_path_id
                and self.offset == other.offset)

    def __hash__(self):
        return hash((self.region_path_id, self.offset))

    def __repr__(self):
        return "%d:%d" % (self.region_path_id, self.offset)


class Interval:
    """obg ``Interval``/``DirectedInterval`` (value type only)."""

    def __init__(self, start_position, end_position, region_paths=None,
                 graph=None, direction=None):
        if region_paths is None:
            region_paths = [start_position.region_path_id]
            if end_position.region_path_id != region_paths[0]:
                region_paths.append(end_position.region_path_id)
        region_paths = [int(r) for r in region_paths]
        if not isinstance(start_position, Position):
            start_position = Position(region_paths[0], start_position)
        if not isinstance(end_position, Position):
            end_position = Position(region_paths[-1], end_position)
        self.start_position = start_position
        self.end_position = end_position
        self.region_paths = region_paths
        self.graph = graph
        self.direction = 1 if direction is None else direction

    def length(self):
        g = self.graph
        total = sum(g.node_size(rp) for rp in self.region_paths)
        return int(total - self.start_position.offset
                   - (g.node_size(self.region_paths[-1]) - self.end_position.offset))

    def get_reverse(self):
        g = self.graph
        sp, ep = self.start_position, self.end_position
        return self.__class__(
            Position(-ep.region_path_id, g.node_size(ep.region_path_id) - ep.offset),
            Position(-sp.region_path_id, g.node_size(sp.region_path_id) - sp.offset),
            [-rp for rp in reversed(self.region_paths)], graph=g)

    def get_subinterval(self, start_offset, end_offset):
        g = self.graph
        walked = -self.start_position.offset
        start = end = None
        rps = []
        for rp in self.region_paths:
            size = g.node_size(rp)
            if start is None and walked + size > start_offset:
                start = Position(rp, start_offset - walked)
            if start is not None:
                rps.append(rp)
            if walked + size >= end_offset:
                end = Position(rp, end_offset - walked)
                break
            walked += size
        if start is None or end is None:
            raise ValueError("sub-interval outside interval")
        return Interval(start, end, rps, graph=g)

    def hash(self, ignore_end_pos=False):
        if ignore_end_pos:
            return hash(self.start_position)
        return hash((self.start_position, self.end_position,
                     tuple(self.region_paths)))

    def __eq__(self, other):
        return (self.start_position == other.start_position
                and self.end_position == other.end_position
                and list(self.region_paths) == list(other.region_paths))

    def __hash__(self):
        return self.hash()

    def __repr__(self):
        return "Interval(%r, %r, %r)" % (self.start_position, self.end_position,
                                         self.region_paths)

    def to_file_line(self):
        return json.dumps({"start": self.start_position.offset,
                           "end": self.end_position.offset,
                           "region_paths": self.region_paths,
                           "direction": int(self.direction)})

    @classmethod
    def from_file_line(cls, line, graph=None):
        o = json.loads(line)
        return cls(o["start"], o["end"], o["region_paths"],
                   direction=o.get("direction"), graph=graph)


DirectedInterval = Interval


class ReadBatch:
    """All reads of one chromosome as flat arrays (host side, numpy)."""

    def __init__(self, start_node, start_off, end_node, end_off, path_ptr, path):
        self.start_node = np.ascontiguousarray(start_node, dtype=np.int32)
        self.start_off = np.ascontiguousarray(start_off, dtype=np.int32)
        self.end_node = np.ascontiguousarray(end_node, dtype=np.int32)
        self.end_off = np.ascontiguousarray(end_off, dtype=np.int32)
        self.path_ptr = np.ascontiguousarray(path_ptr, dtype=np.int64)
        self.path = np.ascontiguousarray(path, dtype=np.int32)
        n = self.start_node.size
        if not (self.start_off.size == n == self.end_node.size == self.end_off.size
                and self.path_ptr.size == n + 1
                and (n == 0 or self.path_ptr[-1] == self.path.size)):
            raise ValueError("inconsistent read arrays")
        self._device = None

    def __len__(self):
        return int(self.start_node.size)

    @classmethod
    def from_intervals(cls, intervals):
        """Flatten any iterable of obg-style intervals (host loop: ingest,
        not the measured path)."""
        sn, so, en, eo, ptr, path = [], [], [], [], [0], []
        for iv in intervals:
            sn.append(iv.start_position.region_path_id)
            so.append(iv.start_position.offset)
            en.append(iv.end_position.region_path_id)
            eo.append(iv.end_position.offset)
            path.extend(int(r) for r in iv.region_paths)
            ptr.append(len(path))
        return cls(sn, so, en, eo, ptr, path)

    @classmethod
    def _from_text(cls, text, call):
        """Two calls of a native text parser: sizes, then fill.  ``call(arrays or
        Nones, n_reads, n_nodes)`` binds the entry point and its extra arguments."""
        import ctypes
        from . import _lib
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        n_reads, n_nodes = ctypes.c_uint64(0), ctypes.c_uint64(0)
        call(vp(text), [None] * 6, n_reads, n_nodes)
        r, m = int(n_reads.value), int(n_nodes.value)
        sn, so, en, eo = (np.empty(r, dtype=np.int32) for _ in range(4))
        ptr = np.zeros(r + 1, dtype=np.int64)
        path = np.empty(m, dtype=np.int32)
        if r:
            call(vp(text), [vp(a) for a in (sn, so, en, eo, ptr, path)], n_reads, n_nodes)
        return cls(sn, so, en, eo, ptr, path)

    @classmethod
    def from_intervalcollection_file(cls, file_name, n_threads=0):
        """Reads an ``.intervalcollection`` text file (one JSON interval per line,
        obg ``IntervalCollection.to_file(text_file=True)``) with the library's
        native parser (``gpc_parse_intervals_host``, host threads) -- same result
        as ``from_intervals(IntervalCollection.from_file(file_name))`` at a few
        hundred times the speed of the json module."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        text = _map_file(file_name)

        def call(tp, out, n_reads, n_nodes):
            _lib.check(lib.gpc_parse_intervals_host(tp, text.size, int(n_threads), *out,
                                                    ctypes.byref(n_reads), ctypes.byref(n_nodes)))
        return cls._from_text(text, call)

    @classmethod
    def from_vg_json_file(cls, file_name, graph=None, n_threads=0):
        """Reads vg alignments in JSON form (``vg view -aj``, one Alignment per
        line) with the native parser ``gpc_parse_vg_alignments_host``: what
        ``pyvg.conversion.vg_json_file_to_interval_collection(file_name, graph)``
        returns for the reference (callpeaks_interface.py:24,64), as flat arrays.
        Alignments without a path (unaligned reads) are dropped; their number is
        left in ``n_without_path``.  With a graph, a read naming a node outside it
        raises ``IntervalNotInGraphException`` as pyvg does."""
        import ctypes
        from . import _lib
        from .custom_exceptions import IntervalNotInGraphException, InvalidPileupInterval
        lib = _lib.load()
        text = _map_file(file_name)
        lo, hi = (0, 0) if graph is None else (graph.min_node, graph.min_node + graph.n_nodes - 1)
        skipped = ctypes.c_uint64(0)

        def call(tp, out, n_reads, n_nodes):
            try:
                _lib.check(lib.gpc_parse_vg_alignments_host(
                    tp, text.size, int(n_threads), int(lo), int(hi), *out,
                    ctypes.byref(n_reads), ctypes.byref(n_nodes), ctypes.byref(skipped)))
            except InvalidPileupInterval as e:
                raise IntervalNotInGraphException(str(e)) from None
        batch = cls._from_text(text, call)
        batch.n_without_path = int(skipped.value)
        return batch

    def __iter__(self):
        for i in range(len(self)):
            yield self.interval(i)

    def interval(self, i, graph=None):
        lo, hi = self.path_ptr[i], self.path_ptr[i + 1]
        return Interval(Position(self.start_node[i], self.start_off[i]),
                        Position(self.end_node[i], self.end_off[i]),
                        self.path[lo:hi].tolist(), graph=graph)

    def select(self, keep):
        """Sub-batch of the reads whose index is in / mask is set in ``keep``."""
        keep = np.asarray(keep)
        idx = np.flatnonzero(keep) if keep.dtype == bool else keep.astype(np.int64)
        lens = np.diff(self.path_ptr)[idx]
        ptr = np.r_[0, np.cumsum(lens)].astype(np.int64)
        gather = (np.repeat(self.path_ptr[:-1][idx] - ptr[:-1], lens)
                  + np.arange(ptr[-1]))
        return ReadBatch(self.start_node[idx], self.start_off[idx],
                         self.end_node[idx], self.end_off[idx], ptr,
                         self.path[gather])

    def lengths(self, graph):
        """Base pairs covered by every read (obg ``Interval.length()``): the
        sizes of its nodes minus what lies before the start and after the end."""
        size = np.asarray(graph.node_len, dtype=np.int64)
        index = np.abs(self.path.astype(np.int64)) - graph.min_node
        total = np.r_[0, np.cumsum(size[index])]
        covered = total[self.path_ptr[1:]] - total[self.path_ptr[:-1]]
        last = size[np.abs(self.end_node.astype(np.int64)) - graph.min_node]
        return covered - self.start_off - (last - self.end_off)

    def nbytes(self):
        return sum(a.nbytes for a in (self.start_node, self.start_off,
                                      self.end_node, self.end_off,
                                      self.path_ptr, self.path))

    # ---- the form that crosses the host link ------------------------------------
    def compact(self):
        """(offs uint32[R] = start offset | end offset << 16, path_len uint8[R], path) --
        15 bytes per 36-bp read instead of 30: start / end node repeat the ends of the path,
        path offsets are a prefix sum the device recomputes.  None when a read does not fit
        (offset >= 65536, more than 255 path nodes, empty path, or a start / end node that is
        not the end of its path): such a batch goes up as the six arrays."""
        n = len(self)
        if n == 0:
            return None
        lens = np.diff(self.path_ptr)
        if lens.min() < 1 or lens.max() > 255:
            return None
        if min(self.start_off.min(), self.end_off.min()) < 0 or \
                max(self.start_off.max(), self.end_off.max()) > 0xffff:
            return None
        if not (np.array_equal(self.path[self.path_ptr[:-1]], self.start_node)
                and np.array_equal(self.path[self.path_ptr[1:] - 1], self.end_node)):
            return None
        offs = self.start_off.astype(np.uint32) | (self.end_off.astype(np.uint32) << np.uint32(16))
        return offs, lens.astype(np.uint8), self.path

    def compact_delta(self):
        """The compact form with delta-coded paths: (offs uint32[R], path_len uint8[R],
        first_node int32[R], step int8[sum(path_len) - R]) -- 10.3 bytes per 36-bp read.
        None when the compact form does not apply or a step between neighbouring path nodes
        does not fit int8 (ids not sorted along the graph)."""
        c = self.compact()
        if c is None:
            return None
        offs, lens, path = c
        first = np.ascontiguousarray(path[self.path_ptr[:-1]], dtype=np.int32)
        inner = np.ones(path.size, dtype=bool)
        inner[self.path_ptr[:-1]] = False
        step = (path[1:].astype(np.int64) - path[:-1].astype(np.int64))[inner[1:]]
        if step.size and (step.min() < -128 or step.max() > 127):
            return None
        return offs, lens, first, step.astype(np.int8)

    def pin_memory(self, delta=None):
        """A batch over the same reads whose arrays -- and, when the reads allow it, their
        compact form -- sit in page-locked memory: uploads of such a batch are asynchronous
        copies of the compact form (ingest-side preparation, done once per read set);
        ``delta``: with delta-coded paths when the node ids allow it."""
        import torch
        if delta is None:                # GPC_WIRE_DELTA=0: plain compact form (measurements)
            delta = os.environ.get("GPC_WIRE_DELTA", "1") != "0"
        full = [torch.from_numpy(a).pin_memory() for a in (
            self.start_node, self.start_off, self.end_node, self.end_off,
            self.path_ptr.astype(np.int32) if self.path.size < 2 ** 31 else self.path_ptr, self.path)]
        out = ReadBatch.__new__(ReadBatch)
        (out.start_node, out.start_off, out.end_node, out.end_off, out.path_ptr, out.path) = \
            [t.numpy() for t in full]
        out._device = None
        out._pinned = full
        c = self.compact()
        out._compact = None if c is None else [torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                                               for a in c[:2]] + [full[5]]
        d = self.compact_delta() if c is not None and delta else None
        out._compact_delta = None if d is None else out._compact[:2] + [
            torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in d[2:]]
        return out

    def fresh_view(self):
        """The same host arrays without the cached device copy (every call uploads again)."""
        out = ReadBatch.__new__(ReadBatch)
        out.__dict__.update(self.__dict__)
        out._device = None
        return out

    def wire_form(self):
        """What an upload of this batch sends: "compact+delta", "compact", "pinned" or "pageable"."""
        if getattr(self, "_compact_delta", None) is not None:
            return "compact+delta"
        if getattr(self, "_compact", None) is not None:
            return "compact"
        return "pinned" if getattr(self, "_pinned", None) is not None else "pageable"

    def upload_bytes(self):
        c = getattr(self, "_compact_delta", None)
        if c is None:
            c = getattr(self, "_compact", None)
        if c is not None:
            return sum(t.numel() * t.element_size() for t in c)
        p = getattr(self, "_pinned", None)
        if p is not None:
            return sum(t.numel() * t.element_size() for t in p)
        return self.nbytes()

    def to_file(self, file_name):
        np.savez(file_name, start_node=self.start_node, start_off=self.start_off,
                 end_node=self.end_node, end_off=self.end_off,
                 path_ptr=self.path_ptr, path=self.path)

    @classmethod
    def from_file(cls, file_name):
        o = np.load(file_name)
        return cls(o["start_node"], o["start_off"], o["end_node"], o["end_off"],
                   o["path_ptr"], o["path"])
