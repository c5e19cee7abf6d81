"""GPU: the BASELINE workloads at their FULL size against digests of the oracle's output.

tests/golden/digest_<workload>.json were written once, in the build container, by
oracle/make_digests.py: the oracle port (pinned against the live reference) run over
every chromosome of the workload, SHA-256 of every integer result of the path and a
thinned sample of the float tracks.  Here the CUDA path runs the same seeded inputs and
must reproduce the digests: unique reads, direct pileup, fragment pileup, touched
nodes, thresholded and hole-cleaned tracks, peak intervals with node lists bit-exact;
lambda, p, q and the joined q table within 1e-6 relative (parity_util).

human_wg (24 chromosomes, 3.1 Gbp, 80 M reads: minutes of input generation alone) runs
at full size only with GPC_FULL_SIZE=1; by default its 1/8-scale digest is used.
"""
import json
import os

import numpy as np
import pytest

from graph_peak_caller_b200 import synthetic
from graph_peak_caller_b200.control.linearmap import LinearMap
from oracle.make_digests import peaks_digest, sha
from parity_util import ATOL, RTOL, step_at

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = os.environ.get("GPC_FULL_SIZE") == "1"

CASES = [("chr22_scale", 1), ("dense_variation", 1), ("arabidopsis_shaped", 1),
         ("human_wg", 1 if FULL else 8)]


def cpu(t):
    return t.cpu().numpy()


def upos(t):
    return cpu(t).view(np.uint32).astype(np.int64)


def check_thin(doc, pos, val, what):
    """The oracle's every-k-th run against the device track as a step function."""
    q = np.asarray(doc["positions"], dtype=np.int64)
    want = np.asarray(doc["values"], dtype=np.float64)
    got = step_at(pos, val, q)
    bad = np.flatnonzero(np.abs(got - want) > RTOL * np.abs(want) + ATOL)
    assert bad.size == 0, "%s: %d/%d sampled runs differ, first at %d: got %r want %r" % (
        what, bad.size, q.size, q[bad[0]], got[bad[0]], want[bad[0]])


@pytest.mark.gpu
@pytest.mark.parametrize("workload,scale", CASES)
def test_full_size_digests(workload, scale):
    import torch
    from graph_peak_caller_b200 import pipeline
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    path = os.path.join(GOLDEN, "digest_%s%s.json" % (workload, "" if scale == 1 else "_div%g" % scale))
    if not os.path.exists(path):
        pytest.skip("no digest file " + os.path.basename(path))
    doc = json.load(open(path))
    cfg = synthetic.CONFIGS[workload]
    if scale > 1:
        cfg = synthetic.scaled(cfg, scale)
    R, F, has_control = cfg["read_length"], cfg["fragment_length"], cfg["has_control"]
    assert (doc["read_length"], doc["fragment_length"], doc["has_control"]) == (R, F, has_control)
    gmin = synthetic.global_min_of(cfg)
    kept, tabs = [], []
    for ci, want in enumerate(doc["chromosomes"]):
        g, s, c = synthetic.chromosome_inputs(cfg, ci)
        assert (g.size_bp, g.n_nodes, g.n_edges) == (want["graph_bp"], want["nodes"], want["edges"])
        dg = DeviceGraph(g, LinearMap.from_graph(g).as_arrays())
        ds = DeviceReads(s)
        dc = DeviceReads(c) if c is not None else ds
        st = pipeline.run_to_p_values(dg, ds, dc, R, F, has_control, gmin, True)
        name = "%s/%s" % (workload, want["name"])
        assert [st.n_sample, st.n_control] == want["unique_reads"], name + " unique reads"
        assert sha(upos(st.direct[0]), cpu(st.direct[1])) == want["direct"], name + " direct pileup"
        assert sha(upos(st.sample[0]), cpu(st.sample[1])) == want["fragment"], name + " fragment pileup"
        touched = np.flatnonzero(cpu(st.touched)) + g.min_node
        assert sha(touched) == want["touched"], name + " touched nodes"
        check_thin(want["control"], upos(st.control[0]), cpu(st.control[1]), name + " control")
        check_thin(want["p_values"], upos(st.p_values[0]), cpu(st.p_values[1]), name + " p-values")
        tabs.append(pipeline.local_p_table(st.p_values, st.track_size))
        kept.append((dg, st.p_values, st.touched, st.direct, st.track_size, g.min_node))
        del st, ds, dc
    tp, tc = torch.cat([t[0] for t in tabs]), torch.cat([t[1] for t in tabs])
    total_bp = sum(k[4] for k in kept)
    assert total_bp == doc["q_table"]["total_bp"]
    qp, qq = pipeline.q_table(tp, tc, total_bp)
    # joined table: the oracle's sampled (p, q) pairs, looked up by value
    tab_p, tab_q = cpu(qp), cpu(qq)
    order = np.argsort(tab_p)
    tab_p, tab_q = tab_p[order], tab_q[order]
    for p, q in zip(doc["q_table"]["p"], doc["q_table"]["q"]):
        if p == 0:
            continue
        j = min(int(np.searchsorted(tab_p, p)), tab_p.size - 1)
        j = j if abs(tab_p[j] - p) <= abs(tab_p[max(j - 1, 0)] - p) else j - 1
        assert abs(tab_p[j] - p) <= RTOL * abs(p) + ATOL, "p %r missing from the joined table" % p
        assert abs(tab_q[j] - q) <= RTOL * abs(q) + 1e-9, "q(%r): got %r want %r" % (p, tab_q[j], q)
    n_peaks = 0
    for (dg, p_values, touched, direct, size, min_node), want in zip(kept, doc["chromosomes"]):
        name = "%s/%s" % (workload, want["name"])
        q, thr, cleaned, raw = pipeline.run_from_p_values(dg, p_values, touched, direct, qp, qq, R, F)
        check_thin(want["q_values"], upos(q[0]), cpu(q[1]), name + " q-values")
        assert sha(upos(thr[0]), cpu(thr[1]).astype(np.int64)) == want["thresholded"], \
            name + " thresholded track"
        assert sha(upos(cleaned[0]), cpu(cleaned[1]).astype(np.int64)) == want["hole_cleaned"], \
            name + " hole-cleaned track"
        peaks = pipeline.PeakSet(raw, F)
        assert peaks_digest(peaks.all_peaks()) == (want["all_max_paths"], want["n_all_max_paths"]), \
            name + " max paths"
        assert peaks_digest(peaks.final_peaks()) == (want["peaks"], want["n_peaks"]), name + " peaks"
        n_peaks += want["n_peaks"]
    assert n_peaks > 0
