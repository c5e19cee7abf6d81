"""GPU: degenerate inputs of every stage through the C ABI -- empty and
single-element tracks, read sets that are all duplicates, nothing / everything
above the threshold, no holes, no peaks, one-element sorts and scans."""
import numpy as np
import pytest

from graph_peak_caller_b200.reads import ReadBatch
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import pipeline as opipeline, postprocess as opost, sample as osample, stats as ostats

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.cpu().numpy()


def dev(a):
    from graph_peak_caller_b200.device import to_device
    return to_device(a)


def u(t):
    return cpu(t).view(np.uint32).astype(np.int64)


def empty_reads():
    z = np.zeros(0, dtype=np.int32)
    return ReadBatch(z, z, z, z, np.zeros(1, dtype=np.int64), z)


def test_sample_pileup_without_reads():
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    g = make_graph(3000, 0.05, seed=1)
    dg, dr = DeviceGraph(g), DeviceReads(empty_reads())
    idx, n = kernels.unique_reads(dr)
    assert n == 0 and idx.numel() == 0
    events, touched = kernels.sample_events(dg, dr, None, 0, 50)
    assert events.numel() == 0 and int(cpu(touched).sum()) == 0
    (fi, fv), (di, dv) = kernels.events_to_runs(events, g.size_bp)
    # the empty track is the single entry (0, 0), as SparseValues of an all-zero pileup
    assert u(fi).tolist() == [0] and cpu(fv).tolist() == [0]
    assert u(di).tolist() == [0] and cpu(dv).tolist() == [0]


def test_single_read_and_all_duplicates():
    from test_gpu_sample import check
    g = make_graph(2000, 0.1, seed=2)
    one = make_reads(g, 1, 15, 60, seed=3, dup_frac=0.0)
    check(g, one, 45, True)
    # the same read 257 times: one survives, pileup of a single read
    k = 257
    rep = ReadBatch(np.repeat(one.start_node, k), np.repeat(one.start_off, k),
                    np.repeat(one.end_node, k), np.repeat(one.end_off, k),
                    np.arange(k + 1, dtype=np.int64) * one.path.size, np.tile(one.path, k))
    got = check(g, rep, 45, True)
    assert got["n"] == 1 and got["index"].tolist() == [0]
    # without the filter the counts are k-fold
    both = check(g, rep, 45, False)
    assert both["sample"][1].max() == k


def test_pvalues_of_trivial_tracks():
    from graph_peak_caller_b200 import kernels
    # no coverage at all: p == 0 everywhere, one run
    spos, sval = dev(np.array([0], dtype=np.int32)), dev(np.array([0], dtype=np.int32))
    cpos, cval = dev(np.array([0], dtype=np.int32)), dev(np.array([0.25]))
    ppos, p = kernels.pvalues(spos, sval, 1.0, cpos, cval)
    assert u(ppos).tolist() == [0] and cpu(p).tolist() == [0.0]
    tp, tc = kernels.ptable_local(ppos, p, 1000)
    assert tp.numel() == 0                       # p == 0 never enters the table
    qp, qq = kernels.qtable_build(tp, tc, 1000)
    assert qp.numel() == 0
    # one covered stretch
    spos, sval = dev(np.array([0, 10, 20], dtype=np.int32)), dev(np.array([0, 3, 0], dtype=np.int32))
    ppos, p = kernels.pvalues(spos, sval, 1.0, cpos, cval)
    ref_p = ostats.clean_p_values(np.array([0.0, 3.0, 0.0]), np.array([0.25, 0.25, 0.25]))
    assert u(ppos).tolist() == [0, 10, 20]
    np.testing.assert_allclose(cpu(p), ref_p, rtol=1e-6, atol=1e-12)
    tp, tc = kernels.ptable_local(ppos, p, 1000)
    assert cpu(tc).tolist() == [10]
    qp, qq = kernels.qtable_build(tp, tc, 1000)
    q = kernels.qvalues_apply(p, qp, qq)
    sp, cum = ostats.p_table(ref_p, np.array([10.0, 10.0, 980.0]))
    want = ostats.q_table(sp, cum)
    np.testing.assert_allclose(cpu(q), [want[float(x)] for x in ref_p], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("values,cutoff,want_pos,want_val", [
    ([0.0, 0.5, 0.2], 1.0, [0], [0]),                 # nothing reaches the threshold
    ([2.0, 3.5, 1.2], 1.0, [0], [1]),                 # everything does
    ([0.0, 2.0, 0.0, 2.0], 2.0, [0, 5, 9, 30], [0, 1, 0, 1]),     # >= is inclusive
])
def test_threshold_degenerate(values, cutoff, want_pos, want_val):
    from graph_peak_caller_b200 import kernels
    pos = np.array([0, 5, 9, 30][:len(values)], dtype=np.int32)
    tpos, tval = kernels.threshold(dev(pos), dev(np.array(values)), cutoff)
    assert u(tpos).tolist() == want_pos and cpu(tval).astype(int).tolist() == want_val


def test_holes_and_peaks_of_constant_tracks():
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph
    g = make_graph(5000, 0.05, seed=4)
    dg = DeviceGraph(g)
    touched = dev(np.ones(g.n_nodes, dtype=np.uint8))
    score_pos, score_val = dev(np.array([0], dtype=np.int32)), dev(np.array([1], dtype=np.int32))
    q_pos, q_val = dev(np.array([0], dtype=np.int32)), dev(np.array([3.0]))
    for value in (0, 1):
        pos, val = dev(np.array([0], dtype=np.int32)), dev(np.array([value], dtype=np.uint8))
        cpos, cval = kernels.clean_holes(dg, pos, val, 36, touched)
        ref = opost.clean_holes(g, np.array([0]), np.array([bool(value)]), 36,
                                set(range(g.min_node, g.min_node + g.n_nodes)))
        assert u(cpos).tolist() == ref[0].tolist()
        assert cpu(cval).astype(bool).tolist() == ref[1].tolist()
    # an all-zero track has no peaks
    raw = kernels.max_paths(dg, dev(np.array([0], dtype=np.int32)), dev(np.array([0], dtype=np.uint8)),
                            score_pos, score_val, q_pos, q_val, g.min_node - 1)
    assert raw["n"] == 0 and raw["nodes"].numel() == 0


@pytest.mark.parametrize("n", [0, 1, 2])
def test_tiny_sorts_and_scans(n):
    from graph_peak_caller_b200 import kernels
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 31, size=n).astype(np.int32)
    assert np.array_equal(cpu(kernels.sort_u32(dev(keys))), np.sort(keys))
    k64 = rng.integers(0, 2 ** 62, size=n).astype(np.int64)
    vals = np.arange(n, dtype=np.int32)
    sk, sv = kernels.sort_u64_pairs(dev(k64), dev(vals))
    order = np.argsort(k64, kind="stable")
    assert np.array_equal(cpu(sk), k64[order]) and np.array_equal(cpu(sv), vals[order])
    x = rng.integers(0, 100, size=n).astype(np.int32)
    out, total = kernels.exclusive_scan_u32(dev(x))
    assert np.array_equal(cpu(out), np.cumsum(x) - x) and total == int(x.sum())


def test_whole_path_when_nothing_is_significant():
    """Uniform reads only: q-values never reach the cut-off, no peaks, and the
    classes return empty collections instead of failing."""
    from graph_peak_caller_b200 import CallPeaks, Configuration
    from graph_peak_caller_b200.control.linearmap import LinearMap
    from graph_peak_caller_b200.intervals import Intervals
    from graph_peak_caller_b200.reporter import Reporter
    g = make_graph(20000, 0.03, seed=5)
    s = make_reads(g, 300, 20, 80, seed=6, peak_frac=0.0, dup_frac=0.0)
    c = make_reads(g, 300, 20, 80, seed=7, peak_frac=0.0, dup_frac=0.0)
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = 20, 80, True
    config.linear_map_name = LinearMap.from_graph(g)
    caller = CallPeaks(g, config, Reporter(None))
    caller.run(Intervals(s), Intervals(c))
    ref = opipeline.callpeaks(g, config.linear_map_name.as_arrays(), s, c, 20, 80, True)
    assert len(caller.max_path_peaks) == len(ref["peaks"])
