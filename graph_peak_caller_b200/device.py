"""HBM residency: the graph and the reads as torch CUDA tensors plus the plain
pointer structs of include/gpc_b200.h.  PyTorch is only the allocator and the
stream here; every computation is a kernel of libgpc_b200.so."""
import ctypes
import os

import numpy as np

from . import _lib
from .custom_exceptions import DeviceLibraryMissing


_torch = None


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


def torch_cuda():
    """torch, after checking once that a CUDA device is there (the check reads the
    environment every time and this is called a few hundred times per pass)."""
    global _torch
    if _torch is None:
        import torch
        if not torch.cuda.is_available():
            raise DeviceLibraryMissing(
                "no CUDA device visible: the callpeaks path runs on the GPU only")
        _torch = torch
    return _torch


def cuda_available():
    """True when a CUDA device is visible.  Only ``LinearMap.from_graph`` asks -- the
    preprocessing step that also exists as a host sweep for machines without a device;
    nothing on the callpeaks path itself has a host form."""
    if _torch is not None:
        return True
    import torch
    return bool(torch.cuda.is_available())


_devices = {}


def device():
    torch = torch_cuda()
    index = torch.cuda.current_device()
    dev = _devices.get(index)
    if dev is None:
        dev = _devices[index] = torch.device("cuda", index)
    return dev


def stream_ptr():
    """cudaStream_t of torch's current stream (the raw accessor when this torch
    has it: the Stream object costs microseconds, and every stage asks)."""
    torch = torch_cuda()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return ctypes.c_void_p(raw(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def to_device(a, dtype=None):
    torch = torch_cuda()
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
    return torch.from_numpy(a).to(device(), non_blocking=False)


# Guard mode (GPC_GUARD=1, see csrc/common.cuh): the output arrays the kernels fill are
# poisoned and fenced by canaries too; check_guards() verifies the fences of every array
# handed out since the last check.
GUARD = os.environ.get("GPC_GUARD", "0") not in ("", "0")
_GUARD_PAD = 256
_guarded = []


def empty(n, dtype):
    torch = torch_cuda()
    if not GUARD:
        return torch.empty(int(n), dtype=dtype, device=device())
    size = torch.empty(0, dtype=dtype).element_size()
    nbytes = int(n) * size
    body = (nbytes + 15) & ~15
    raw = torch.full((body + 2 * _GUARD_PAD,), 0xA5, dtype=torch.uint8, device=device())
    _guarded.append((raw, nbytes))
    return raw[_GUARD_PAD:_GUARD_PAD + nbytes].view(dtype)


def check_guards():
    """Number of guarded arrays (Python-side outputs + the library's scratch) found written
    outside their bounds since the last call; 0 outside guard mode."""
    torch = torch_cuda()
    bad = 0
    torch.cuda.synchronize()
    for raw, nbytes in _guarded:
        head = raw[:_GUARD_PAD]
        tail = raw[_GUARD_PAD + nbytes:]
        if bool((head != 0xA5).any()) or bool((tail != 0xA5).any()):
            bad += 1
    _guarded.clear()
    lib = _lib.load()
    total = int(lib.gpc_guard_violations())
    bad += total - check_guards.seen
    check_guards.seen = total
    return bad


check_guards.seen = 0


class DeviceGraph:
    """Graph CSR in HBM (include/gpc_b200.h: gpc_graph_t)."""

    def __init__(self, graph, linear_map=None):
        torch = torch_cuda()
        n = graph.n_nodes
        rec = np.zeros((n + 1, 4), dtype=np.uint32)
        rec[:, 0] = graph.node_off
        rec[:, 1] = graph.succ_ptr
        rec[:, 2] = graph.pred_ptr
        rec[:n, 3] = graph.node_len
        # walk records: one 32-byte sector per node with the first two neighbours inline
        wrec = np.zeros((n, 8), dtype=np.uint32)
        deg_s = np.diff(graph.succ_ptr)
        deg_p = np.diff(graph.pred_ptr)
        wrec[:, 0] = graph.node_off[:n]
        wrec[:, 1] = graph.node_len
        wrec[:, 2] = np.minimum(deg_s, 0xffff) | (np.minimum(deg_p, 0xffff) << 16)
        for col, ptr, adj, deg, k in ((3, graph.succ_ptr, graph.succ, deg_s, 0),
                                      (4, graph.succ_ptr, graph.succ, deg_s, 1),
                                      (5, graph.pred_ptr, graph.pred, deg_p, 0),
                                      (6, graph.pred_ptr, graph.pred, deg_p, 1)):
            has = deg > k
            wrec[has, col] = adj[ptr[:n][has] + k]
        self.graph = graph
        self.walk_rec = to_device(wrec.view(np.int32))
        self.node_rec = to_device(rec.view(np.int32))
        self.succ = to_device(graph.succ.astype(np.int32))
        self.pred = to_device(graph.pred.astype(np.int32))
        self.lin_start = self.lin_end = None
        if linear_map is not None:
            self.set_linear_map(*linear_map)
        self._struct()

    def set_linear_map(self, starts, ends):
        starts = np.asarray(starts, dtype=np.float64)
        ends = np.asarray(ends, dtype=np.float64)
        self.lin_start = to_device(starts)
        self.lin_end = to_device(ends)
        self.linear_length = float(ends[-1])
        # every node's linear span a whole multiple of its length (scale of
        # control/linearmap.py:43-49 an integer everywhere): read starts project to integers
        span, size = ends - starts, self.graph.node_len.astype(np.float64)
        self.integer_linear = bool(size.size and np.all(size > 0) and np.all(starts == np.floor(starts))
                                   and np.all(span == np.floor(span)) and np.all(np.fmod(span, size) == 0)
                                   and float(ends.max()) <= self.graph.size_bp)
        self._struct()

    def _struct(self):
        g = self.graph
        self.c = _lib.GpcGraph(g.n_nodes, g.n_edges, g.size_bp, g.min_node,
                               ptr(self.node_rec), ptr(self.succ), ptr(self.pred),
                               ptr(self.lin_start), ptr(self.lin_end), ptr(self.walk_rec),
                               (0 if g.is_ordered else 1) | (2 if getattr(self, "integer_linear", False) else 0))

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.node_rec, self.succ, self.pred, self.lin_start, self.lin_end,
                    self.walk_rec)
                   if t is not None)


class DeviceReads:
    """Reads SoA in HBM (include/gpc_b200.h: gpc_reads_t)."""

    def __init__(self, batch, copy_stream=None):
        torch = torch_cuda()
        # (no back-reference to the batch: batch._device -> this object would make a
        #  cycle and the device arrays would wait for the cyclic collector)
        self.n = len(batch)
        self.host_bytes = batch.nbytes()
        self.ready_event = None
        pinned = getattr(batch, "_pinned", None)
        compact = getattr(batch, "_compact", None)
        null = ctypes.c_void_p(0)
        lib = _lib.load()

        def on_copy_stream(fn):
            """Run fn (copies + the kernels that finish the upload) on the copy stream when
            there is one; the compute stream later waits for ready_event."""
            if copy_stream is None:
                fn(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                return
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                fn(ctypes.c_void_p(copy_stream.cuda_stream))
                self.ready_event = torch.cuda.Event()
                self.ready_event.record(copy_stream)

        self.end_node = self.end_off = self.path_ptr = None
        delta = getattr(batch, "_compact_delta", None)
        if delta is not None:
            # page-locked compact form with delta-coded paths: four asynchronous copies, one
            # kernel rebuilds the path array and the per-read records
            dev = device()
            self.start_node = torch.empty(self.n, dtype=torch.int32, device=dev)
            self.start_off = torch.empty(self.n, dtype=torch.int32, device=dev)
            self.hdr = torch.empty(4 * self.n, dtype=torch.int32, device=dev)
            self.path = torch.empty(int(batch.path.size), dtype=torch.int32, device=dev)
            dst = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in delta]

            def go(stream):
                for a, b in zip(dst, delta):
                    a.copy_(b, non_blocking=True)
                _lib.check(lib.gpc_reads_from_compact_delta(
                    ptr(dst[0]), ptr(dst[1]), ptr(dst[2]), ptr(dst[3]), self.n, ptr(self.path),
                    ptr(self.start_node), ptr(self.start_off), ptr(self.hdr), stream))
            on_copy_stream(go)
            self._wire = dst                # (kept until the widening kernel has run)
            self.host_bytes = batch.upload_bytes()
            self.c = _lib.GpcReads(self.n, ptr(self.start_node), ptr(self.start_off), null, null,
                                   null, ptr(self.path), null, ptr(self.hdr))
            return
        if compact is not None:
            # page-locked compact form: three asynchronous copies, then one kernel widens
            # them into what the duplicate filter, the projection and the walk read
            dev = device()
            self.start_node = torch.empty(self.n, dtype=torch.int32, device=dev)
            self.start_off = torch.empty(self.n, dtype=torch.int32, device=dev)
            self.hdr = torch.empty(4 * self.n, dtype=torch.int32, device=dev)
            dst = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in compact]

            def go(stream):
                for a, b in zip(dst, compact):
                    a.copy_(b, non_blocking=True)
                _lib.check(lib.gpc_reads_from_compact(ptr(dst[0]), ptr(dst[1]), ptr(dst[2]), self.n,
                                                      ptr(self.start_node), ptr(self.start_off),
                                                      ptr(self.hdr), stream))
            on_copy_stream(go)
            self.path = dst[2]
            self._wire = dst[:2]            # (kept until the widening kernel has run)
            self.host_bytes = batch.upload_bytes()
            self.c = _lib.GpcReads(self.n, ptr(self.start_node), ptr(self.start_off), null, null,
                                   null, ptr(self.path), null, ptr(self.hdr))
            return
        if pinned is not None:      # page-locked host tensors: asynchronous copies
            dev = device()
            # destinations come from the current stream's cached pool; with a copy
            # stream only the transfers run there (it first waits for whatever the
            # current stream did with that memory before)
            dst = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in pinned]

            def go(stream):
                for a, b in zip(dst, pinned):
                    a.copy_(b, non_blocking=True)
            on_copy_stream(go)
            (self.start_node, self.start_off, self.end_node, self.end_off, self.path_ptr,
             self.path) = dst
        else:
            self.start_node = to_device(batch.start_node)
            self.start_off = to_device(batch.start_off)
            self.end_node = to_device(batch.end_node)
            self.end_off = to_device(batch.end_off)
            self.path_ptr = to_device(batch.path_ptr)
            self.path = to_device(batch.path)
        # 32-bit path offsets (a page-locked batch may carry them: 4 bytes less per
        # read on the link) go into path_ptr32, 64-bit ones into path_ptr
        narrow = self.path_ptr.element_size() == 4
        self.c = _lib.GpcReads(self.n, ptr(self.start_node), ptr(self.start_off),
                               ptr(self.end_node), ptr(self.end_off),
                               null if narrow else ptr(self.path_ptr), ptr(self.path),
                               ptr(self.path_ptr) if narrow else null, null)
        # the walk's view of a read: one 16-byte record (include/gpc_b200.h, gpc_reads_t.hdr),
        # built once per upload, in the order of the upload
        self.hdr = None
        if self.n and self.path.numel() < 2 ** 32:
            self.hdr = torch.empty(4 * self.n, dtype=torch.int32, device=device())
            on_copy_stream(lambda stream: _lib.check(lib.gpc_build_read_headers(
                ctypes.byref(self.c), ptr(self.hdr), stream)))
            self.c.hdr = self.hdr.data_ptr()


def device_graph(graph, linear_map=None):
    """Cached upload of a host Graph.  A linear map that differs from the one installed
    (another LinearMap object with other numbers) replaces it; the same arrays, or equal
    ones, are not uploaded again."""
    dg = getattr(graph, "_device", None)
    if dg is None:
        dg = DeviceGraph(graph, linear_map)
        graph._device = dg
        dg._lm_host = linear_map
    elif linear_map is not None:
        have = getattr(dg, "_lm_host", None)
        same = have is not None and all(
            a is b or (np.shape(a) == np.shape(b) and np.array_equal(a, b))
            for a, b in zip(have, linear_map))
        if dg.lin_start is None or not same:
            dg.set_linear_map(*linear_map)
            dg._lm_host = linear_map
    return dg


def device_reads(batch):
    dr = getattr(batch, "_device", None)
    if dr is None:
        dr = DeviceReads(batch)
        batch._device = dr
    ready = getattr(dr, "ready_event", None)
    if ready is not None:                 # transfers issued by prefetch_reads on the copy stream
        torch_cuda().cuda.current_stream().wait_event(ready)
        dr.ready_event = None
    return dr


_copy_streams = {}          # one per device index


def prefetch_reads(batch):
    """Upload a ReadBatch on a dedicated copy stream; device_reads() later makes
    the compute stream wait for it."""
    if getattr(batch, "_device", None) is not None:
        return
    torch = torch_cuda()
    if getattr(batch, "_pinned", None) is None:
        return                            # pageable memory: nothing to overlap
    index = torch.cuda.current_device()
    stream = _copy_streams.get(index)
    if stream is None:
        stream = _copy_streams[index] = torch.cuda.Stream()
    batch._device = DeviceReads(batch, copy_stream=stream)
