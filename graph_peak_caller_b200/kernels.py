"""Thin Python wrappers over the C ABI: allocate outputs with torch, call the
kernel entry point, trim to the reported size.  One function per entry point
of include/gpc_b200.h."""
import ctypes

from . import _lib
from .device import empty, ptr, stream_ptr, torch_cuda


def _u64():
    return ctypes.c_uint64(0)


def sort_u32(keys, begin_bit=0, end_bit=32):
    """keys: int32 CUDA tensor holding uint32 bit patterns.  Returns sorted tensor."""
    torch = torch_cuda()
    lib = _lib.load()
    tmp = torch.empty_like(keys)
    flag = ctypes.c_int32(0)
    _lib.check(lib.gpc_sort_u32(ptr(keys), ptr(tmp), keys.numel(), begin_bit, end_bit,
                                ctypes.byref(flag), stream_ptr()))
    return tmp if flag.value else keys


def sort_u64_pairs(keys, vals=None, begin_bit=0, end_bit=64):
    torch = torch_cuda()
    lib = _lib.load()
    ktmp = torch.empty_like(keys)
    vtmp = torch.empty_like(vals) if vals is not None else None
    flag = ctypes.c_int32(0)
    _lib.check(lib.gpc_sort_u64_pairs(ptr(keys), ptr(ktmp), ptr(vals), ptr(vtmp),
                                      keys.numel(), begin_bit, end_bit,
                                      ctypes.byref(flag), stream_ptr()))
    if flag.value:
        return ktmp, vtmp
    return keys, vals


def exclusive_scan_u32(x):
    torch = torch_cuda()
    lib = _lib.load()
    out = torch.empty_like(x)
    total = _u64()
    _lib.check(lib.gpc_exclusive_scan_u32(ptr(x), ptr(out), x.numel(),
                                          ctypes.byref(total), stream_ptr()))
    return out, total.value


def unique_reads(dreads):
    """S1.  Returns (index tensor int32[n_unique], n_unique)."""
    torch = torch_cuda()
    lib = _lib.load()
    out = empty(max(dreads.n, 1), torch.int32)
    count = _u64()
    _lib.check(lib.gpc_unique_reads(ptr(dreads.start_node), ptr(dreads.start_off),
                                    dreads.n, ptr(out), ctypes.byref(count),
                                    stream_ptr()))
    return out[:count.value], count.value


def sample_events(dgraph, dreads, read_index, n_reads, extension, events_per_read=6):
    """S2-S4.  Returns (events int32[n_events], touched uint8[n_nodes])."""
    torch = torch_cuda()
    lib = _lib.load()
    touched = empty(dgraph.graph.n_nodes, torch.uint8)
    cap = max(1024, int(n_reads) * events_per_read)
    while True:
        events = empty(cap, torch.int32)
        n_events = _u64()
        status = lib.gpc_sample_events(
            ctypes.byref(dgraph.c), ctypes.byref(dreads.c), ptr(read_index), n_reads,
            extension, ptr(events), cap, ctypes.byref(n_events), ptr(touched),
            stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            cap = int(n_events.value) + 1024
            continue
        _lib.check(status)
        return events[:n_events.value], touched


def events_to_runs(events, size_bp):
    """T1/T2.  Returns ((frag_idx, frag_val), (direct_idx, direct_val))."""
    torch = torch_cuda()
    lib = _lib.load()
    n = events.numel()
    tmp = torch.empty_like(events)
    fi, fv = empty(n + 1, torch.int32), empty(n + 1, torch.int32)
    di, dv = empty(n + 1, torch.int32), empty(n + 1, torch.int32)
    nf, nd = _u64(), _u64()
    _lib.check(lib.gpc_events_to_runs(ptr(events), ptr(tmp), n, size_bp, ptr(fi), ptr(fv),
                                      ctypes.byref(nf), ptr(di), ptr(dv),
                                      ctypes.byref(nd), stream_ptr()))
    return (fi[:nf.value], fv[:nf.value]), (di[:nd.value], dv[:nd.value])
