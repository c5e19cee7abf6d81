"""Thin Python wrappers over the C ABI: allocate outputs with torch, call the
kernel entry point, trim to the reported size.  One function per entry point
of include/gpc_b200.h."""
import ctypes

from . import _lib
from .device import empty, ptr, stream_ptr, to_device, torch_cuda


def _u64():
    return ctypes.c_uint64(0)


def sort_u32(keys, begin_bit=0, end_bit=32):
    """keys: int32 CUDA tensor holding uint32 bit patterns.  Returns sorted tensor."""
    torch = torch_cuda()
    lib = _lib.load()
    tmp = empty(keys.numel(), keys.dtype)
    flag = ctypes.c_int32(0)
    _lib.check(lib.gpc_sort_u32(ptr(keys), ptr(tmp), keys.numel(), begin_bit, end_bit,
                                ctypes.byref(flag), stream_ptr()))
    return tmp if flag.value else keys


def sort_u64_pairs(keys, vals=None, begin_bit=0, end_bit=64):
    torch = torch_cuda()
    lib = _lib.load()
    ktmp = empty(keys.numel(), keys.dtype)
    vtmp = empty(vals.numel(), vals.dtype) if vals is not None else None
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
    out = empty(x.numel(), x.dtype)
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


def sample_pileup(dgraph, dreads, read_index, n_reads, extension, mode=0):
    """S2-S5 + T1/T2 in one call.  Returns ((frag_idx, frag_val), (direct_idx, direct_val),
    touched uint8[n_nodes], n_events).  All four run arrays live in one allocation, sized
    from the runs per read the graph produced last time (a bubble-rich graph cuts a
    fragment into more runs); too small means one more call."""
    torch = torch_cuda()
    lib = _lib.load()
    touched = empty(dgraph.graph.n_nodes, torch.uint8)
    per_read = getattr(dgraph, "_runs_per_read", 2.6)
    cap = max(1024, int(int(n_reads) * per_read) + 16)
    while True:
        arena = empty(4 * cap, torch.int32)
        fi, fv, di, dv = (arena[k * cap:(k + 1) * cap] for k in range(4))
        nf, nd, ne = _u64(), _u64(), _u64()
        status = lib.gpc_sample_pileup(
            ctypes.byref(dgraph.c), ctypes.byref(dreads.c), ptr(read_index), n_reads, extension,
            ptr(fi), ptr(fv), ptr(di), ptr(dv), cap, ctypes.byref(nf), ctypes.byref(nd),
            ctypes.byref(ne), ptr(touched), mode, stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            cap = max(int(nf.value), int(nd.value)) + 1024
            continue
        _lib.check(status)
        if n_reads:
            dgraph._runs_per_read = 1.15 * max(int(nf.value), int(nd.value), 1) / int(n_reads) + 0.05
        return ((fi[:nf.value], fv[:nf.value]), (di[:nd.value], dv[:nd.value]), touched,
                int(ne.value))


def events_to_runs(events, size_bp):
    """T1/T2.  Returns ((frag_idx, frag_val), (direct_idx, direct_val))."""
    torch = torch_cuda()
    lib = _lib.load()
    n = events.numel()
    tmp = empty(events.numel(), events.dtype)
    fi, fv = empty(n + 1, torch.int32), empty(n + 1, torch.int32)
    di, dv = empty(n + 1, torch.int32), empty(n + 1, torch.int32)
    nf, nd = _u64(), _u64()
    _lib.check(lib.gpc_events_to_runs(ptr(events), ptr(tmp), n, size_bp, ptr(fi), ptr(fv),
                                      ctypes.byref(nf), ptr(di), ptr(dv),
                                      ctypes.byref(nd), stream_ptr()))
    return (fi[:nf.value], fv[:nf.value]), (di[:nd.value], dv[:nd.value])


TRACK_OPS = {"maximum": 0, "minimum": 1, "add": 2, "subtract": 3, "multiply": 4}


def _event_list(indices, diffs):
    """(int64 indices, float64 diffs) of an event list as CUDA tensors (host arrays are
    uploaded)."""
    import numpy as np
    torch = torch_cuda()
    if not type(indices).__module__.startswith("torch"):
        indices = to_device(np.asarray(indices).astype(np.int64, copy=False))
        diffs = to_device(np.asarray(diffs).astype(np.float64, copy=False))
    return indices.to(torch.int64), diffs.to(torch.float64)


def diffs_to_runs(indices, diffs):
    """T1/T2, generic: any (index, diff) list -> canonical runs (pos int32, value
    float64) on the device."""
    torch = torch_cuda()
    lib = _lib.load()
    idx, d = _event_list(indices, diffs)
    n = idx.numel()
    pos, val = empty(max(n, 1), torch.int32), empty(max(n, 1), torch.float64)
    n_out = _u64()
    _lib.check(lib.gpc_diffs_to_runs(ptr(idx), ptr(d), n, ptr(pos), ptr(val),
                                     ctypes.byref(n_out), stream_ptr()))
    return pos[:n_out.value], val[:n_out.value]


def diffs_merge(idx_a, diffs_a, idx_b, diffs_b):
    """Distinct indices of two event lists and both tracks' running sums there:
    (pos int32, a float64, b float64)."""
    torch = torch_cuda()
    lib = _lib.load()
    ia, da = _event_list(idx_a, diffs_a)
    ib, db = _event_list(idx_b, diffs_b)
    n = ia.numel() + ib.numel()
    pos = empty(max(n, 1), torch.int32)
    a, b = empty(max(n, 1), torch.float64), empty(max(n, 1), torch.float64)
    n_out = _u64()
    _lib.check(lib.gpc_diffs_merge(ptr(ia), ptr(da), ia.numel(), ptr(ib), ptr(db), ib.numel(),
                                   ptr(pos), ptr(a), ptr(b), ctypes.byref(n_out), stream_ptr()))
    m = n_out.value
    return pos[:m], a[:m], b[:m]


def diffs_combine(idx_a, diffs_a, idx_b, diffs_b, op, drop_leading_zero=False):
    """op(a, b) of two event-list tracks as runs without repeated values: (pos int32,
    value float64).  op: a key of TRACK_OPS."""
    torch = torch_cuda()
    lib = _lib.load()
    ia, da = _event_list(idx_a, diffs_a)
    ib, db = _event_list(idx_b, diffs_b)
    n = ia.numel() + ib.numel()
    pos, val = empty(max(n, 1), torch.int32), empty(max(n, 1), torch.float64)
    n_out = _u64()
    _lib.check(lib.gpc_diffs_combine(ptr(ia), ptr(da), ia.numel(), ptr(ib), ptr(db), ib.numel(),
                                     TRACK_OPS[op], 1 if drop_leading_zero else 0, ptr(pos),
                                     ptr(val), ctypes.byref(n_out), stream_ptr()))
    return pos[:n_out.value], val[:n_out.value]


def control_linear_track(dgraph, dreads, read_index, n_reads, extensions,
                         fragment_length, min_value, want_x=False):
    """C2/C3.  Returns (bp float64[U], val float64[U], x or None)."""
    torch = torch_cuda()
    lib = _lib.load()
    cap = max(2 * len(extensions) * int(n_reads), 1)
    bp, val = empty(cap, torch.float64), empty(cap, torch.float64)
    x = empty(max(int(n_reads), 1), torch.float64) if want_x else None
    ext = (ctypes.c_int32 * len(extensions))(*[int(e) for e in extensions])
    n_out = _u64()
    _lib.check(lib.gpc_control_linear_track(
        ctypes.byref(dgraph.c), ptr(dreads.start_node), ptr(dreads.start_off),
        ptr(read_index), n_reads, ext, len(extensions), int(fragment_length),
        float(min_value), ptr(bp), ptr(val), ctypes.byref(n_out), ptr(x), stream_ptr()))
    return bp[:n_out.value], val[:n_out.value], x


def control_backproject(dgraph, touched, bp, val, min_value, value_scale=1.0):
    """C4/C5.  Returns (pos int32[Uc], lambda float64[Uc])."""
    torch = torch_cuda()
    lib = _lib.load()
    cap = dgraph.graph.n_nodes + 2 * bp.numel() + 16
    while True:
        pos, lam = empty(cap, torch.int32), empty(cap, torch.float64)
        n_out = _u64()
        status = lib.gpc_control_backproject(
            ctypes.byref(dgraph.c), ptr(touched), ptr(bp), ptr(val), bp.numel(),
            float(min_value), float(value_scale), ptr(pos), ptr(lam), cap,
            ctypes.byref(n_out), stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            cap = int(n_out.value) + 16
            continue
        _lib.check(status)
        return pos[:n_out.value], lam[:n_out.value]


def pvalues(spos, sval, sample_scale, cpos, cval, control_scale=1.0):
    """P1.  Returns (pos int32[Up], p float64[Up])."""
    torch = torch_cuda()
    lib = _lib.load()
    cap = spos.numel() + cpos.numel()
    pos, p = empty(max(cap, 1), torch.int32), empty(max(cap, 1), torch.float64)
    n_out = _u64()
    _lib.check(lib.gpc_pvalues(ptr(spos), ptr(sval), spos.numel(), float(sample_scale),
                               ptr(cpos), ptr(cval), cpos.numel(), float(control_scale),
                               ptr(pos), ptr(p),
                               ctypes.byref(n_out), stream_ptr()))
    return pos[:n_out.value], p[:n_out.value]


def ptable_local(pos, p, track_size):
    """Q1 (local half).  Returns (p float64[D], bp int64[D]) unordered."""
    global _ptable_cap
    torch = torch_cuda()
    lib = _lib.load()
    cap = _ptable_cap          # sized from the last table: a table that does not fit costs a whole second pass
    while True:
        tp, tc = empty(cap, torch.float64), empty(cap, torch.int64)
        n_out = _u64()
        status = lib.gpc_ptable_local(ptr(pos), ptr(p), pos.numel(), int(track_size),
                                      ptr(tp), ptr(tc), cap, ctypes.byref(n_out),
                                      stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            cap = int(n_out.value) + 16
            continue
        _lib.check(status)
        _ptable_cap = max(1 << 16, int(n_out.value) * 5 // 4 + 16)
        return tp[:n_out.value], tc[:n_out.value]


_ptable_cap = 1 << 16


def qtable_build(table_p, table_count, total_bp):
    """Q1 (global half).  Returns (p descending float64[D], q float64[D])."""
    torch = torch_cuda()
    lib = _lib.load()
    n = table_p.numel()
    op, oq = empty(max(n, 1), torch.float64), empty(max(n, 1), torch.float64)
    n_out = _u64()
    _lib.check(lib.gpc_qtable_build(ptr(table_p), ptr(table_count), n, int(total_bp),
                                    ptr(op), ptr(oq), ctypes.byref(n_out), stream_ptr()))
    return op[:n_out.value], oq[:n_out.value]


def qvalues_apply(p, table_p, table_q):
    torch = torch_cuda()
    lib = _lib.load()
    q = empty(p.numel(), p.dtype)
    _lib.check(lib.gpc_qvalues_apply(ptr(p), p.numel(), ptr(table_p), ptr(table_q),
                                     table_p.numel(), ptr(q), stream_ptr()))
    return q


def threshold(pos, q, cutoff):
    """H1.  Returns (pos int32[Ut], value uint8[Ut])."""
    torch = torch_cuda()
    lib = _lib.load()
    n = pos.numel()
    op, ov = empty(max(n, 1), torch.int32), empty(max(n, 1), torch.uint8)
    n_out = _u64()
    _lib.check(lib.gpc_threshold(ptr(pos), ptr(q), n, float(cutoff), ptr(op), ptr(ov),
                                 ctypes.byref(n_out), stream_ptr()))
    return op[:n_out.value], ov[:n_out.value]


def clean_holes(dgraph, pos, val, max_size, touched):
    """H2-H4.  Returns (pos int32[Uh], value uint8[Uh])."""
    torch = torch_cuda()
    lib = _lib.load()
    cap = pos.numel() + 16
    while True:
        op, ov = empty(cap, torch.int32), empty(cap, torch.uint8)
        n_out = _u64()
        status = lib.gpc_clean_holes(ctypes.byref(dgraph.c), ptr(pos), ptr(val), pos.numel(),
                                     int(max_size), ptr(touched), ptr(op), ptr(ov), cap,
                                     ctypes.byref(n_out), stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            cap = int(n_out.value) + 16
            continue
        _lib.check(status)
        return op[:n_out.value], ov[:n_out.value]


def max_paths(dgraph, pos, val, score_pos, score_val, q_pos, q_val, internal_id_shift):
    """M1-M3.  Returns a dict of CUDA tensors (see include/gpc_b200.h)."""
    torch = torch_cuda()
    lib = _lib.load()
    pcap, ncap = max(1024, pos.numel()), max(1 << 16, 32 * pos.numel())
    fields = (("score", torch.float64, 8), ("start", torch.int32, 4), ("end", torch.int32, 4),
              ("length", torch.int32, 4), ("order", torch.int32, 4), ("node_ptr", torch.int32, 4),
              ("info", torch.uint8, 1), ("nodes", torch.int32, 4))   # ragged part last
    while True:
        # one arena for all outputs, so that the host reads them back in one copy
        counts = dict(score=pcap, start=pcap, end=pcap, length=pcap, order=pcap,
                      node_ptr=pcap + 1, nodes=ncap, info=pcap)
        layout, off = {}, 0
        for name, dtype, size in fields:
            layout[name] = (off, dtype, size)
            off += (counts[name] * size + 15) & ~15
        arena = empty(off, torch.uint8)
        t = {name: arena[o:o + counts[name] * size].view(dtype)
             for name, (o, dtype, size) in layout.items()}
        n_peaks, n_nodes = _u64(), _u64()
        status = lib.gpc_max_paths(
            ctypes.byref(dgraph.c), ptr(pos), ptr(val), pos.numel(), ptr(score_pos),
            ptr(score_val) if score_val.dtype == torch.float64 else None,
            ptr(score_val) if score_val.dtype == torch.int32 else None,
            score_pos.numel(), ptr(q_pos), ptr(q_val), q_pos.numel(),
            int(internal_id_shift), ptr(t["start"]), ptr(t["end"]), ptr(t["length"]),
            ptr(t["score"]), ptr(t["info"]), ptr(t["node_ptr"]), pcap, ptr(t["nodes"]), ncap,
            ptr(t["order"]), ctypes.byref(n_peaks), ctypes.byref(n_nodes), stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            pcap = max(pcap, int(n_peaks.value) + 16)
            ncap = max(ncap, int(n_nodes.value) + 16)
            continue
        _lib.check(status)
        n, m = int(n_peaks.value), int(n_nodes.value)
        out = {k: v[:n] for k, v in t.items() if k not in ("node_ptr", "nodes")}
        out["node_ptr"] = t["node_ptr"][:n + 1]
        out["nodes"] = t["nodes"][:m]
        out["n"] = n
        out["arena"] = (arena, {name: (o, dtype, size, out[name].numel())
                                for name, (o, dtype, size) in layout.items()})
        return out


def peak_subgraphs(dgraph, pos, val, score_pos, score_val):
    """Line-graph components of the peaks (the reference's sub-graphs).  Returns host
    numpy arrays: comp_ptr[n_comp + 1], member node ids, member weights, member exit flags,
    row_ptr[n_members + 1], col (local to the component)."""
    torch = torch_cuda()
    lib = _lib.load()
    ccap, mcap, ecap = 1024, max(1024, pos.numel() + 1), max(4096, 2 * pos.numel())
    while True:
        comp_ptr = empty(ccap, torch.int32)
        node, weight = empty(mcap, torch.int32), empty(mcap, torch.int32)
        ex, row_ptr = empty(mcap, torch.uint8), empty(mcap, torch.int32)
        col = empty(ecap, torch.int32)
        nc, nm, ne = _u64(), _u64(), _u64()
        status = lib.gpc_peak_subgraphs(
            ctypes.byref(dgraph.c), ptr(pos), ptr(val), pos.numel(), ptr(score_pos),
            ptr(score_val) if score_val.dtype == torch.float64 else None,
            ptr(score_val) if score_val.dtype == torch.int32 else None, score_pos.numel(),
            ptr(comp_ptr), ccap, ptr(node), ptr(weight), ptr(ex), ptr(row_ptr), mcap, ptr(col),
            ecap, ctypes.byref(nc), ctypes.byref(nm), ctypes.byref(ne), stream_ptr())
        if status == _lib.GPC_ERR_CAPACITY:
            ccap = max(ccap, int(nc.value) + 16)
            mcap = max(mcap, int(nm.value) + 16)
            ecap = max(ecap, int(ne.value) + 16)
            continue
        _lib.check(status)
        c, m, e = int(nc.value), int(nm.value), int(ne.value)
        if c == 0:
            import numpy as np
            z = np.zeros(0, dtype=np.int32)
            return np.zeros(1, dtype=np.int32), z, z, z.astype(np.uint8), np.zeros(1, dtype=np.int32), z
        return (comp_ptr[:c + 1].cpu().numpy(), node[:m].cpu().numpy(), weight[:m].cpu().numpy(),
                ex[:m].cpu().numpy(), row_ptr[:m + 1].cpu().numpy(), col[:e].cpu().numpy())


def peak_summits(dgraph, start, end, node_ptr, nodes, q_pos, q_val, window):
    """Summits.  start / end int32[n], node_ptr int32[n + 1], nodes
    int32[...] (node ids), q track (pos int32, val float64): all CUDA tensors.  Returns
    int32[n, 6] = summit, length, first, last, start_off, end_off per peak."""
    torch = torch_cuda()
    lib = _lib.load()
    n = start.numel()
    out = empty(6 * n, torch.int32)
    _lib.check(lib.gpc_peak_summits(ctypes.byref(dgraph.c), ptr(start), ptr(end), ptr(node_ptr),
                                    ptr(nodes), n, ptr(q_pos), ptr(q_val), q_pos.numel(),
                                    int(window), ptr(out), stream_ptr()))
    return out.view(n, 6)


def node_read_counts(path, min_node, n_nodes):
    """Forward visits per node over all read paths (fragment-length estimation,
    SURVEY.md §8f-4).  path: int32 CUDA tensor of signed node ids.  Returns int32[n_nodes]."""
    torch = torch_cuda()
    lib = _lib.load()
    counts = empty(max(int(n_nodes), 1), torch.int32)
    _lib.check(lib.gpc_node_read_counts(ptr(path), path.numel(), int(min_node), int(n_nodes),
                                        ptr(counts), stream_ptr()))
    return counts[:int(n_nodes)]


def path_start_positions(dgraph, path, start_offset, start_node, start_off):
    """Read starts on a path, per strand, in read order: (plus int64[...], minus int64[...])
    CUDA tensors."""
    torch = torch_cuda()
    lib = _lib.load()
    n = start_node.numel()
    plus, minus = empty(max(n, 1), torch.int64), empty(max(n, 1), torch.int64)
    n_plus, n_minus = _u64(), _u64()
    _lib.check(lib.gpc_path_start_positions(
        ctypes.byref(dgraph.c), ptr(path), path.numel(), int(start_offset), ptr(start_node),
        ptr(start_off), n, ptr(plus), ctypes.byref(n_plus), ptr(minus), ctypes.byref(n_minus),
        stream_ptr()))
    return plus[:n_plus.value], minus[:n_minus.value]
