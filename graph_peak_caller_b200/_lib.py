"""ctypes binding of include/gpc_b200.h.  The library is mandatory: there is
no CPU fallback, every failure to load it is an error."""
import ctypes
import os

from .custom_exceptions import DeviceLibraryMissing, InvalidPileupInterval

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPC_B200_LIB", os.path.join(HERE, "libgpc_b200.so"))

GPC_OK = 0
GPC_ERR_CUDA = 1
GPC_ERR_ARG = 2
GPC_ERR_CAPACITY = 3
GPC_ERR_INVALID_INTERVAL = 4
GPC_ERR_FRONTIER_OVERFLOW = 5
GPC_ERR_INTERNAL = 6

c_void_p = ctypes.c_void_p
u32, u64, i32, i64 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64
f64 = ctypes.c_double


class GpcGraph(ctypes.Structure):
    _fields_ = [("n_nodes", u32), ("n_edges", u32), ("size_bp", u32),
                ("min_node", i32), ("node_rec", c_void_p), ("succ", c_void_p),
                ("pred", c_void_p), ("lin_start", c_void_p), ("lin_end", c_void_p),
                ("walk_rec", c_void_p), ("flags", u32)]


class GpcReads(ctypes.Structure):
    _fields_ = [("n_reads", u64), ("start_node", c_void_p), ("start_off", c_void_p),
                ("end_node", c_void_p), ("end_off", c_void_p),
                ("path_ptr", c_void_p), ("path", c_void_p), ("path_ptr32", c_void_p),
                ("hdr", c_void_p)]


P = ctypes.POINTER
_SIGNATURES = {
    "gpc_unique_reads": [c_void_p, c_void_p, u64, c_void_p, P(u64), c_void_p],
    "gpc_sample_events": [P(GpcGraph), P(GpcReads), c_void_p, u64, i32, c_void_p,
                          u64, P(u64), c_void_p, c_void_p],
    "gpc_build_read_headers": [P(GpcReads), c_void_p, c_void_p],
    "gpc_reads_from_compact": [c_void_p, c_void_p, c_void_p, u64, c_void_p, c_void_p, c_void_p,
                               c_void_p],
    "gpc_reads_from_compact_delta": [c_void_p, c_void_p, c_void_p, c_void_p, u64, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p],
    "gpc_sample_pileup": [P(GpcGraph), P(GpcReads), c_void_p, u64, i32, c_void_p, c_void_p,
                          c_void_p, c_void_p, u64, P(u64), P(u64), P(u64), c_void_p, i32,
                          c_void_p],
    "gpc_diffs_to_runs": [c_void_p, c_void_p, u64, c_void_p, c_void_p, P(u64), c_void_p],
    "gpc_diffs_merge": [c_void_p, c_void_p, u64, c_void_p, c_void_p, u64, c_void_p, c_void_p,
                        c_void_p, P(u64), c_void_p],
    "gpc_diffs_combine": [c_void_p, c_void_p, u64, c_void_p, c_void_p, u64, i32, i32, c_void_p,
                          c_void_p, P(u64), c_void_p],
    "gpc_events_to_runs": [c_void_p, c_void_p, u64, u32, c_void_p, c_void_p, P(u64),
                           c_void_p, c_void_p, P(u64), c_void_p],
    "gpc_control_linear_track": [P(GpcGraph), c_void_p, c_void_p, c_void_p, u64,
                                 P(i32), i32, i32, f64, c_void_p, c_void_p, P(u64),
                                 c_void_p, c_void_p],
    "gpc_control_backproject": [P(GpcGraph), c_void_p, c_void_p, c_void_p, u64, f64, f64,
                                c_void_p, c_void_p, u64, P(u64), c_void_p],
    "gpc_pvalues": [c_void_p, c_void_p, u64, f64, c_void_p, c_void_p, u64, f64, c_void_p,
                    c_void_p, P(u64), c_void_p],
    "gpc_ptable_local": [c_void_p, c_void_p, u64, u32, c_void_p, c_void_p, u64, P(u64),
                         c_void_p],
    "gpc_qtable_build": [c_void_p, c_void_p, u64, u64, c_void_p, c_void_p, P(u64), c_void_p],
    "gpc_qvalues_apply": [c_void_p, u64, c_void_p, c_void_p, u64, c_void_p, c_void_p],
    "gpc_threshold": [c_void_p, c_void_p, u64, f64, c_void_p, c_void_p, P(u64), c_void_p],
    "gpc_clean_holes": [P(GpcGraph), c_void_p, c_void_p, u64, i32, c_void_p, c_void_p,
                        c_void_p, u64, P(u64), c_void_p],
    "gpc_max_paths": [P(GpcGraph), c_void_p, c_void_p, u64, c_void_p, c_void_p, c_void_p, u64,
                      c_void_p, c_void_p, u64, i32, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_void_p, c_void_p, u64, c_void_p, u64, c_void_p, P(u64),
                      P(u64), c_void_p],
    "gpc_peak_subgraphs": [P(GpcGraph), c_void_p, c_void_p, u64, c_void_p, c_void_p, c_void_p, u64,
                           c_void_p, u64, c_void_p, c_void_p, c_void_p, c_void_p, u64, c_void_p,
                           u64, P(u64), P(u64), P(u64), c_void_p],
    "gpc_peak_summits": [P(GpcGraph), c_void_p, c_void_p, c_void_p, c_void_p, u64, c_void_p,
                         c_void_p, u64, i32, c_void_p, c_void_p],
    "gpc_parse_intervals_host": [c_void_p, u64, i32, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, P(u64), P(u64)],
    "gpc_parse_vg_alignments_host": [c_void_p, u64, i32, i32, i32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, P(u64), P(u64), P(u64)],
    "gpc_parse_vg_graph_host": [c_void_p, u64, c_void_p, c_void_p, P(u64), c_void_p, c_void_p,
                                P(u64)],
    "gpc_split_alignments_host": [c_void_p, u64, c_void_p, c_void_p, u64, c_void_p, c_void_p],
    "gpc_parse_vg_sequences_host": [c_void_p, u64, c_void_p, c_void_p, c_void_p, P(u64), P(u64)],
    "gpc_shift_model_host": [c_void_p, u64, c_void_p, u64, i64, i64, i64, i64, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, P(u64)],
    "gpc_node_read_counts": [c_void_p, u64, i32, u32, c_void_p, c_void_p],
    "gpc_path_start_positions": [P(GpcGraph), c_void_p, u64, i64, c_void_p, c_void_p, u64, c_void_p,
                                 P(u64), c_void_p, P(u64), c_void_p],
    "gpc_most_covered_path_host": [c_void_p, c_void_p, c_void_p, c_void_p, u64, c_void_p, P(u64)],
    "gpc_linear_map_from_graph_host": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       u64, c_void_p, c_void_p],
    "gpc_linear_map_from_graph": [P(GpcGraph), c_void_p, c_void_p, P(u64), c_void_p],
    "gpc_sort_u32": [c_void_p, c_void_p, u64, i32, i32, P(i32), c_void_p],
    "gpc_sort_u64_pairs": [c_void_p, c_void_p, c_void_p, c_void_p, u64, i32, i32,
                           P(i32), c_void_p],
    "gpc_exclusive_scan_u32": [c_void_p, c_void_p, u64, P(u64), c_void_p],
}

_lib = None


def exported_symbols():
    """Every entry point include/gpc_b200.h declares."""
    return ["gpc_last_error_string", "gpc_version_string", "gpc_launch_count",
            "gpc_profile_enable", "gpc_profile_report", "gpc_guard_mode",
            "gpc_guard_violations"] + sorted(_SIGNATURES)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DeviceLibraryMissing(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.gpc_last_error_string.restype = ctypes.c_char_p
    lib.gpc_version_string.restype = ctypes.c_char_p
    lib.gpc_launch_count.restype = ctypes.c_uint64
    lib.gpc_guard_violations.restype = ctypes.c_uint64
    lib.gpc_guard_mode.restype = ctypes.c_int
    lib.gpc_profile_enable.argtypes = [ctypes.c_int]
    lib.gpc_profile_enable.restype = None
    lib.gpc_profile_report.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    lib.gpc_profile_report.restype = ctypes.c_uint64
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


class GpcError(RuntimeError):
    def __init__(self, status, text):
        super().__init__("gpc_b200 error %d: %s" % (status, text))
        self.status = status


def check(status):
    """Turn a non-zero status into the exception the reference would raise."""
    if status == GPC_OK:
        return
    text = load().gpc_last_error_string().decode("utf-8", "replace")
    if status == GPC_ERR_INVALID_INTERVAL:
        raise InvalidPileupInterval(text)
    if status == GPC_ERR_ARG:
        raise Exception(text)
    raise GpcError(status, text)


def launch_count():
    return int(load().gpc_launch_count())


def profile_enable(on=True):
    load().gpc_profile_enable(1 if on else 0)


def profile_report():
    """{kernel name: (calls, total milliseconds, algorithmic bytes)} since the
    last report."""
    lib = load()
    need = int(lib.gpc_profile_report(None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    lib.gpc_profile_report(buf, need + 1)
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms, nbytes = line.split()
        out[name] = (int(calls), float(ms), float(nbytes))
    return out
