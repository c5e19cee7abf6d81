"""ORACLE (test infrastructure, CPU) -- the whole callpeaks path end to end.

Restates graph_peak_caller/callpeaks.py:47-123 (CallPeaks) and :126-218
(CallPeaksFromQvalues) of the reference by chaining the stage oracles.  Every
intermediate track is returned so the parity tests can check each CUDA stage
on its own.
"""
import time

import numpy as np

from . import control as ocontrol
from . import postprocess as opost
from . import sample as osample
from . import stats as ostats
from . import tracks


class _Sub:
    """A subset of reads as a flat view (mask or index array)."""

    def __init__(self, reads, keep):
        keep = np.asarray(keep)
        idx = np.flatnonzero(keep) if keep.dtype == bool else keep
        lens = np.diff(reads.path_ptr)[idx]
        ptr = np.r_[0, np.cumsum(lens)].astype(np.int64)
        gather = (np.repeat(reads.path_ptr[:-1][idx] - ptr[:-1], lens)
                  + np.arange(ptr[-1]))
        self.start_node = reads.start_node[idx]
        self.start_off = reads.start_off[idx]
        self.end_node = reads.end_node[idx]
        self.end_off = reads.end_off[idx]
        self.path_ptr = ptr
        self.path = reads.path[gather]


def unique_reads(reads):
    """UniqueIntervals (intervals.py:17-33) as a filtered flat batch."""
    return _Sub(reads, osample.unique_read_mask(reads))


def run_to_p_values(graph, linear_map, sample_reads, control_reads,
                    read_length, fragment_length, has_control=True,
                    global_min=None, timings=None):
    """CallPeaks.run_to_p_values (callpeaks.py:121-123 -> :47-83).  The reads
    are taken as they are (duplicate filtering is the caller's choice, as in
    the reference)."""
    assert fragment_length > read_length, \
        "Read length %d is larger than fragment size" % read_length
    t0 = time.perf_counter()
    out = osample.fragment_pileup(graph, sample_reads,
                                  fragment_length - read_length)
    t1 = time.perf_counter()
    extensions = [fragment_length, 1000, 10000] if has_control else [10000]
    lin_starts, lin_ends = linear_map
    bg = ocontrol.background_track(graph, lin_starts, lin_ends, control_reads,
                                   extensions, fragment_length,
                                   set(out["touched"].tolist()), global_min)
    t2 = time.perf_counter()
    ratio = len(sample_reads.start_node) / len(control_reads.start_node)
    s_idx, s_diffs = out["sample_diffs"]
    c_idx, c_diffs = bg["control_diffs"]
    s_diffs, c_diffs = ocontrol.scale_tracks(s_diffs, c_diffs, ratio)
    p_idx, p_val = ostats.p_value_runs(s_idx, s_diffs, c_idx, c_diffs)
    t3 = time.perf_counter()
    out.update(control_diffs=(c_idx, c_diffs), linear=bg["linear"],
               min_value=bg["min_value"], x=bg["x"], ratio=ratio,
               sample_scaled_diffs=(s_idx, s_diffs),
               control=tracks.diffs_to_runs(c_idx, c_diffs),
               p_values=(p_idx, p_val), track_size=int(graph.node_off[-1]))
    if timings is not None:
        timings.update(sample=t1 - t0, control=t2 - t1, pvalues=t3 - t2)
    return out


def p_table_of(p_idx, p_val, track_size):
    """PToQValuesMapper.from_p_values_pileup (sparsepvalues.py:61-68)."""
    return ostats.p_table(p_val, ostats.run_lengths(p_idx, track_size))


def joined_p_table(chromosomes):
    """PToQValuesMapper.from_files (sparsepvalues.py:76-93): concatenate all
    chromosomes' (p, run length) pairs.  ``chromosomes`` = iterable of
    (p_idx, p_val, track_size)."""
    ps, lens = [], []
    for p_idx, p_val, size in chromosomes:
        ps.append(np.asarray(p_val))
        lens.append(ostats.run_lengths(np.asarray(p_idx), size))
    return ostats.p_table(np.concatenate(ps), np.concatenate(lens))


def run_from_p_values(graph, state, table, read_length, fragment_length,
                      q_threshold=0.05, q_values_max_path=False,
                      replicate_internal_id_quirk=True, timings=None):
    """CallPeaks.get_q_values + call_peaks_from_q_values (callpeaks.py:91-113)
    -> CallPeaksFromQvalues.callpeaks (:214-218).  ``state`` is the dict from
    run_to_p_values; ``table`` the dict p -> q."""
    t0 = time.perf_counter()
    p_idx, p_val = state["p_values"]
    q_idx, q_val = ostats.q_value_runs(p_idx, p_val, table)
    t_idx, t_val = tracks.threshold_runs(q_idx, q_val, -np.log10(q_threshold))
    t1 = time.perf_counter()
    h_idx, h_val = opost.clean_holes(graph, t_idx, t_val, read_length,
                                     set(state["touched"].tolist()))
    t2 = time.perf_counter()
    if q_values_max_path:
        sc_idx, sc_val = q_idx, q_val
    else:
        sc_idx, sc_val = state["direct"]
    peaks = opost.max_paths(graph, h_idx, h_val, sc_idx, sc_val,
                            state["track_size"], replicate_internal_id_quirk)
    final = opost.score_and_filter_peaks(graph, peaks, q_idx, q_val,
                                         fragment_length)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(qvalues=t1 - t0, holes=t2 - t1, maxpaths=t3 - t2)
    return dict(q_values=(q_idx, q_val), thresholded=(t_idx, t_val),
                hole_cleaned=(h_idx, h_val), all_max_paths=peaks, peaks=final)


def callpeaks(graph, linear_map, sample_reads, control_reads, read_length,
              fragment_length, has_control=True, global_min=None,
              q_threshold=0.05, timings=None, replicate_internal_id_quirk=True):
    """CallPeaks.run (callpeaks.py:115-119) for one graph."""
    state = run_to_p_values(graph, linear_map, sample_reads, control_reads,
                            read_length, fragment_length, has_control,
                            global_min, timings)
    sp, cum = p_table_of(*state["p_values"], state["track_size"])
    table = ostats.q_table(sp, cum)
    state["p_table"] = (sp, cum)
    state.update(run_from_p_values(
        graph, state, table, read_length, fragment_length, q_threshold,
        False, replicate_internal_id_quirk, timings))
    return state
