"""The callpeaks path as a chain of kernel calls on device-resident arrays.

This is the body behind ``CallPeaks.run_to_p_values`` /
``get_p_to_q_values_mapping`` / ``get_q_values`` / ``call_peaks_from_q_values``
(reference graph_peak_caller/callpeaks.py:47-123, 126-218); the classes in
callpeaks.py only wrap these functions in the reference's object model, and
bench.py times them directly.  Everything stays in HBM between stages; the
only host traffic is the handful of element counts the C ABI reports.
"""
import numpy as np

from . import kernels
from .device import torch_cuda


class PValueState:
    """Device tensors produced by ``run_to_p_values``."""

    def __init__(self):
        self.sample = None        # (pos int32, count int32)   fragment pileup
        self.direct = None        # (pos int32, count int32)   reads-only pileup
        self.touched = None       # uint8[n_nodes]
        self.control = None       # (pos int32, lambda float64), scaled by ratio if < 1
        self.linear = None        # (breakpoint float64, lambda float64) linear background
        self.p_values = None      # (pos int32, p float64)
        self.n_sample = self.n_control = 0
        self.ratio = 1.0
        self.sample_scale = 1.0
        self.min_value = 0.0
        self.track_size = 0
        self.n_events = 0


def control_extensions(fragment_length, has_control):
    """control/__init__.py:19,27"""
    return [fragment_length, 1000, 10000] if has_control else [10000]


def sample_pileup(dgraph, dreads, read_index, n_reads, extension):
    """get_fragment_pileup (sample/__init__.py:5-9) incl. the direct pileup."""
    if extension < 0:
        raise Exception("Invalid extension size %d used. Must be positive. "
                        "Is fragment length < read length?" % extension)
    frag, direct, touched, n_events = kernels.sample_pileup(dgraph, dreads, read_index, n_reads,
                                                            extension)
    return frag, direct, touched, n_events


def linear_background(dgraph, dreads, read_index, n_reads, extensions, fragment_length,
                      global_min=None):
    """First half of SparseControl.create (controlgenerator.py:21-41): needs only
    the control reads, so it can run beside the sample pileup."""
    if global_min is not None:
        min_value = global_min
    else:
        # controlgenerator.py:24 -- same expression, same operand types
        min_value = n_reads * fragment_length / np.float64(dgraph.linear_length)
    bp, val, _ = kernels.control_linear_track(dgraph, dreads, read_index, n_reads,
                                              extensions, fragment_length, min_value)
    return bp, val, float(min_value)


def background_track(dgraph, dreads, read_index, n_reads, extensions, fragment_length,
                     touched, global_min=None, value_scale=1.0, linear=None):
    """get_background_track (control/__init__.py:7-14) -> SparseControl.create."""
    if linear is None:
        linear = linear_background(dgraph, dreads, read_index, n_reads, extensions,
                                   fragment_length, global_min)
    bp, val, min_value = linear
    cpos, clam = kernels.control_backproject(dgraph, touched, bp, val, min_value, value_scale)
    return (cpos, clam), (bp, val), float(min_value)


def run_to_p_values(dgraph, dsample, dcontrol, read_length, fragment_length,
                    has_control=True, global_min=None, unique=False,
                    sample_index=None, control_index=None):
    """CallPeaks.run_to_p_values (callpeaks.py:121-123).  ``unique`` applies the
    duplicate filter of UniqueIntervals (intervals.py:17-33) to both read sets;
    alternatively pass precomputed (index tensor, count) pairs."""
    assert fragment_length > read_length, \
        "Read length %d is larger than fragment size" % read_length
    st = PValueState()
    extensions = control_extensions(fragment_length, has_control)

    def control_chain():
        ci, nc = control_index if control_index is not None else (None, dcontrol.n)
        if unique:
            ci, nc = kernels.unique_reads(dcontrol)
        return ci, nc, linear_background(dgraph, dcontrol, ci, nc, extensions,
                                         fragment_length, global_min)

    # (running the control chain on a second stream beside the sample chain was measured
    # twice, the second time from a helper thread of its own so that neither chain's count
    # read-backs hold the other up: 5.39 against 5.22 ms per pass -- the kernels of the two
    # chains slow each other down by more than the covered gaps are worth; kept sequential)
    si, ns = sample_index if sample_index is not None else (None, dsample.n)
    if unique:
        si, ns = kernels.unique_reads(dsample)
    st.sample, st.direct, st.touched, st.n_events = sample_pileup(
        dgraph, dsample, si, ns, fragment_length - read_length)
    ci, nc, linear = control_chain()
    st.n_sample, st.n_control = ns, nc
    # scale_tracks (control/__init__.py:32-42)
    st.ratio = ns / nc
    st.sample_scale = 1 / st.ratio if st.ratio > 1 else 1.0
    value_scale = st.ratio if st.ratio < 1 else 1.0
    st.control, st.linear, st.min_value = background_track(
        dgraph, dcontrol, ci, nc, extensions, fragment_length, st.touched, global_min,
        value_scale, linear=linear)
    st.p_values = kernels.pvalues(st.sample[0], st.sample[1], st.sample_scale,
                                  st.control[0], st.control[1])
    st.track_size = dgraph.graph.size_bp
    return st


def local_p_table(p_values, track_size):
    """PToQValuesMapper.__get_sub_counts + first half of _from_subcounts."""
    return kernels.ptable_local(p_values[0], p_values[1], track_size)


def q_table(table_p, table_count, total_bp):
    """get_p_to_q_values (sparsepvalues.py:95-103) -> (p descending, q)."""
    return kernels.qtable_build(table_p, table_count, total_bp)


class PeakSet:
    """Final peaks of one graph, still as flat arrays (host numpy)."""

    def __init__(self, raw, fragment_length):
        if "arena" in raw:          # all outputs live in one buffer: one readback
            arena, layout = raw["arena"]
            last = max(o + size * cnt for o, _, size, cnt in layout.values())
            host = arena[:last].cpu()
            views = {name: host[o:o + size * cnt].view(dtype).numpy()
                     for name, (o, dtype, size, cnt) in layout.items()}
            cpu = lambda t: views[next(k for k in layout if raw[k] is t)]
        else:
            cpu = lambda t: t.cpu().numpy()
        self.n_all = raw["n"]
        self.start = cpu(raw["start"])
        self.end = cpu(raw["end"])
        self.length = cpu(raw["length"])
        self.score = cpu(raw["score"])
        self.info = cpu(raw["info"])
        self.node_ptr = cpu(raw["node_ptr"]).astype(np.int64)
        self.nodes = cpu(raw["nodes"])
        order = cpu(raw["order"]).astype(np.int64)
        # callpeaks.py:204-208: stable sort by score (done on the device), then
        # keep peaks at least one fragment long
        self.order = order[self.length[order] >= fragment_length] if order.size else order

    def __len__(self):
        return int(self.order.size)

    def peak(self, i):
        lo, hi = self.node_ptr[i], self.node_ptr[i + 1]
        return (int(self.start[i]), int(self.end[i]), self.nodes[lo:hi].tolist(),
                (bool(self.info[i] & 1), bool(self.info[i] & 2)), float(self.score[i]))

    def all_peaks(self):
        return [self.peak(i) for i in range(self.n_all)]

    def final_peaks(self):
        return [self.peak(int(i)) for i in self.order]


def run_from_p_values(dgraph, p_values, touched, direct, table_p, table_q, read_length,
                      fragment_length, q_threshold=0.05, q_values_max_path=False,
                      replicate_internal_id_quirk=False):
    """get_q_values + call_peaks_from_q_values (callpeaks.py:91-113, 214-218).
    Returns (q track, thresholded, hole_cleaned, raw peak tensors)."""
    ppos, p = p_values
    q = kernels.qvalues_apply(p, table_p, table_q)
    thr = kernels.threshold(ppos, q, -np.log10(q_threshold))
    cleaned = kernels.clean_holes(dgraph, thr[0], thr[1], read_length, touched)
    if q_values_max_path:
        score_pos, score_val = ppos, q
    else:
        score_pos, score_val = direct[0], direct[1]
    shift = dgraph.graph.min_node - 1
    raw = kernels.max_paths(dgraph, cleaned[0], cleaned[1], score_pos, score_val, ppos, q,
                            -shift if replicate_internal_id_quirk else shift)
    return (ppos, q), thr, cleaned, raw


def callpeaks(dgraph, dsample, dcontrol, read_length, fragment_length, has_control=True,
              global_min=None, unique=False, q_threshold=0.05):
    """CallPeaks.run (callpeaks.py:115-119) for one graph."""
    st = run_to_p_values(dgraph, dsample, dcontrol, read_length, fragment_length,
                         has_control, global_min, unique)
    tp, tc = local_p_table(st.p_values, st.track_size)
    qp, qq = q_table(tp, tc, st.track_size)
    q, thr, cleaned, raw = run_from_p_values(dgraph, st.p_values, st.touched, st.direct,
                                             qp, qq, read_length, fragment_length, q_threshold)
    return st, (qp, qq), q, thr, cleaned, PeakSet(raw, fragment_length)
