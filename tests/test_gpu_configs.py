"""GPU: the other BASELINE.json configurations as parity / property cases.

* Arabidopsis-shaped: several chromosomes, the sample doubles as control
  (has_control False -> one 10 kb window), lambda floor given as global_min;
  small enough for the oracle, through MultipleGraphsCallpeaks.
* Dense variation: 10 % variant sites, long indels, unequal parallel branches
  (non-integer linear scales -> the merge path of the control kernel).
* Medium sizes the Python oracle cannot finish: size-independent properties
  (canonical tracks, conservation of read mass, idempotence, determinism).
"""
import numpy as np
import pytest

from graph_peak_caller_b200 import Configuration
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import control as ocontrol, pipeline as opipeline, stats as ostats
from parity_util import assert_exact, assert_step_close
from test_gpu_postprocess import assert_same_peaks

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.cpu().numpy()


def test_arabidopsis_shaped_sample_as_control_global_min():
    R, F = 12, 60
    lengths = [3040, 1970, 2350, 1860, 2700]            # Mbp x 100 of the five chromosomes
    names, graphs, reads, lms = [], [], [], []
    for k, bp in enumerate(lengths):
        g = make_graph(bp * 4, 0.01, seed=300 + k)
        names.append("chr%d" % (k + 1))
        graphs.append(g)
        reads.append(make_reads(g, bp // 2, R, F, seed=400 + k, peak_frac=0.6, peak_spacing=2500))
        lms.append(LinearMap.from_graph(g))
    total_unique = sum(len(opipeline.unique_reads(r).start_node) for r in reads)
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, False
    config.global_min = total_unique * F / float(sum(g.size_bp for g in graphs))
    caller = MultipleGraphsCallpeaks(names, graphs, reads, reads, lms, config, Reporter(None))
    caller.run()
    states = []
    for g, r in zip(graphs, reads):
        u = opipeline.unique_reads(r)
        states.append(opipeline.run_to_p_values(g, ocontrol.linear_map_from_graph(g), u, u, R, F,
                                                False, config.global_min))
    table = ostats.q_table(*opipeline.joined_p_table(
        [(s["p_values"][0], s["p_values"][1], s["track_size"]) for s in states]))
    n_peaks = 0
    for i, (name, g, st) in enumerate(zip(names, graphs, states)):
        ref = opipeline.run_from_p_values(g, st, table, R, F)
        mine = caller._callers[i]
        assert mine.config.global_min == config.global_min
        q = mine.q_values_pileup
        assert_step_close((q.indices, q.values), ref["q_values"], name, upto=g.size_bp)
        h = mine._reporter.memory["hole_cleaned"]
        assert_exact((h.indices, h.values.astype(bool)), ref["hole_cleaned"], name)
        got = [(p.start_position.offset, p.end_position.offset, list(p.region_paths),
                (bool(p.info[0]), bool(p.info[1])), float(p.score)) for p in caller.peaks()[name]]
        assert_same_peaks(got, ref["peaks"])
        n_peaks += len(got)
    assert n_peaks > 0


def test_dense_variation_graph_vs_oracle():
    from graph_peak_caller_b200 import pipeline
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    R, F = 14, 90
    g = make_graph(40000, 0.10, seed=55, mnp_frac=0.3, indel_mean=50.0, indel_max=500)
    s = make_reads(g, 5000, R, F, seed=56, peak_frac=0.6, peak_spacing=5000)
    c = make_reads(g, 7000, R, F, seed=57, peak_frac=0.0, dup_frac=0.0)
    lm = ocontrol.linear_map_from_graph(g)
    assert np.any(((lm[1] - lm[0]) / g.node_len) % 1 != 0), "needs non-integer scales"
    st, qt, q, thr, cleaned, peaks = pipeline.callpeaks(DeviceGraph(g, lm), DeviceReads(s),
                                                        DeviceReads(c), R, F, True, None, True)
    ref = opipeline.callpeaks(g, lm, opipeline.unique_reads(s), opipeline.unique_reads(c), R, F,
                              True)
    u = lambda t: cpu(t).view(np.uint32).astype(np.int64)
    assert_exact((u(st.sample[0]), cpu(st.sample[1]).astype(float)), ref["sample"], "sample")
    assert_exact((u(st.direct[0]), cpu(st.direct[1]).astype(float)), ref["direct"], "direct")
    assert_step_close((u(st.control[0]), cpu(st.control[1])), ref["control"], "lambda",
                      upto=g.size_bp)
    assert_step_close((u(q[0]), cpu(q[1])), ref["q_values"], "q", upto=g.size_bp)
    assert_exact((u(cleaned[0]), cpu(cleaned[1]).astype(bool)), ref["hole_cleaned"], "holes")
    assert_same_peaks(peaks.final_peaks(), ref["peaks"])


@pytest.fixture(scope="module", params=["medium", "chr22_scale"])
def medium(request):
    """Sizes the Python oracle cannot finish.  "chr22_scale" is BASELINE.json
    configs[1] at its full size (the workload bench.py times)."""
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    if request.param == "medium":
        g = make_graph(4_000_000, 0.05, seed=91, mnp_frac=0.05, indel_mean=10.0)
        s = make_reads(g, 400_000, 36, 200, seed=92)
        c = make_reads(g, 450_000, 36, 200, seed=93, peak_frac=0.0, dup_frac=0.0)
    else:
        from graph_peak_caller_b200.synthetic import CONFIGS
        cfg = CONFIGS["chr22_scale"]
        g = make_graph(cfg["chromosomes"][0][1], cfg["site_rate"], seed=20261001)
        s = make_reads(g, cfg["n_sample"], 36, 200, seed=1)
        c = make_reads(g, cfg["n_control"], 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
    lm = LinearMap.from_graph(g).as_arrays()
    return g, s, c, DeviceGraph(g, lm), DeviceReads(s), DeviceReads(c)


def test_medium_size_invariants(medium):
    from graph_peak_caller_b200 import kernels, pipeline
    g, s, c, dg, ds, dc = medium
    G = g.size_bp
    st, qt, q, thr, cleaned, peaks = pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
    u = lambda t: cpu(t).view(np.uint32).astype(np.int64)
    for pos, val in (st.sample, st.direct, st.p_values, thr, cleaned):
        p, v = u(pos), cpu(val)
        assert p[0] == 0 and np.all(np.diff(p) > 0), "positions strictly increasing from 0"
        assert np.all(v[1:] != v[:-1]), "neighbouring runs differ"
        assert p[-1] <= G
    # conservation: direct pileup mass == bases of the unique reads
    keep = cpu(kernels.unique_reads(ds)[0]).astype(np.int64)
    node_len = g.node_len[np.abs(s.path) - g.min_node]
    per_read = np.add.reduceat(node_len, s.path_ptr[:-1])
    e_node = np.abs(s.end_node) - g.min_node
    length = (per_read - s.start_off - (g.node_len[e_node] - s.end_off))[keep]
    dpos, dval = u(st.direct[0]), cpu(st.direct[1]).astype(np.int64)
    assert int(np.sum(np.diff(np.r_[dpos, G]) * dval)) == int(length.sum())
    # the fragment pileup covers at least the reads, at most reads + extension on every branch
    spos, sval = u(st.sample[0]), cpu(st.sample[1]).astype(np.int64)
    assert int(np.sum(np.diff(np.r_[spos, G]) * sval)) >= int(length.sum())
    assert sval.min() >= 0 and sval[-1] == 0
    # lambda never drops below the (scaled) genome-wide background
    lam = cpu(st.control[1])
    assert lam.min() >= st.min_value * min(1.0, st.ratio) * (1 - 1e-12)
    # q is a monotone function of p for p > 0 (running minimum); p == 0 is forced to
    # q == 0 by the reference (d[0] = 0) even though tiny p get negative q
    p, qv = cpu(st.p_values[1]), cpu(q[1])
    pos = p > 0
    order = np.argsort(p[pos], kind="stable")
    assert np.all(np.diff(qv[pos][order]) >= -1e-12)
    assert np.all(qv[p == 0] == 0)
    # cleaning only fills holes: every thresholded base stays set
    tpos, tval = u(thr[0]), cpu(thr[1]).astype(bool)
    cpos, cval = u(cleaned[0]), cpu(cleaned[1]).astype(bool)
    grid = np.union1d(tpos, cpos)
    tv = tval[np.searchsorted(tpos, grid, side="right") - 1]
    cv = cval[np.searchsorted(cpos, grid, side="right") - 1]
    assert np.all(cv[tv])
    # peaks: inside the graph, at least one fragment long, sorted by score
    final = peaks.final_peaks()
    assert len(final) > 10
    scores = [p[4] for p in final]
    assert all(a >= b for a, b in zip(scores, scores[1:]))
    for start, end, nodes, info, score in final:
        assert all(g.min_node <= v < g.min_node + g.n_nodes for v in nodes)
        assert score >= -np.log10(0.05) - 1e-9


def _add_step_functions(a, b, size):
    """Sum of two canonical step functions (positions, values)."""
    grid = np.union1d(a[0], b[0])
    va = a[1][np.searchsorted(a[0], grid, side="right") - 1]
    vb = b[1][np.searchsorted(b[0], grid, side="right") - 1]
    v = va + vb
    keep = np.r_[True, v[1:] != v[:-1]]
    return grid[keep], v[keep]


def test_pileup_is_additive_over_read_subsets(medium):
    """Linearity at full size: the pileups of two halves of the reads add up,
    bit for bit, to the pileup of all of them (walk, event sort and run
    extraction on 2 x 8 M events against 16 M events)."""
    from graph_peak_caller_b200 import pipeline
    from graph_peak_caller_b200.device import DeviceReads
    g, s, c, dg, ds, dc = medium
    n = len(s)
    u = lambda t: cpu(t).view(np.uint32).astype(np.int64)
    whole = pipeline.sample_pileup(dg, ds, None, n, 164)
    half = np.arange(n) % 2 == 0
    parts = []
    for mask in (half, ~half):
        sub = s.select(np.flatnonzero(mask))
        parts.append(pipeline.sample_pileup(dg, DeviceReads(sub), None, len(sub), 164))
    for k, name in ((0, "fragment"), (1, "direct")):
        a, b = ((u(p[k][0]), cpu(p[k][1]).astype(np.int64)) for p in parts)
        pos, val = _add_step_functions(a, b, g.size_bp)
        assert np.array_equal(pos, u(whole[k][0])), name
        assert np.array_equal(val, cpu(whole[k][1]).astype(np.int64)), name
    touched = cpu(parts[0][2]).astype(bool) | cpu(parts[1][2]).astype(bool)
    assert np.array_equal(touched, cpu(whole[2]).astype(bool))


def test_medium_size_determinism_and_idempotent_dedup(medium):
    from graph_peak_caller_b200 import kernels, pipeline
    g, s, c, dg, ds, dc = medium
    a = pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
    b = pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
    for x, y in ((a[0].sample, b[0].sample), (a[0].p_values, b[0].p_values), (a[4], b[4])):
        assert bool((x[0] == y[0]).all()) and bool((x[1] == y[1]).all())
    assert a[5].final_peaks() == b[5].final_peaks()
    # filtering the already filtered reads changes nothing
    from graph_peak_caller_b200.device import DeviceReads
    idx, n = kernels.unique_reads(ds)
    once = s.select(cpu(idx).astype(np.int64))
    idx2, n2 = kernels.unique_reads(DeviceReads(once))
    assert n2 == n
    assert np.array_equal(cpu(idx2), np.arange(n))


def test_mhc_shaped_config_full_size_vs_oracle():
    """BASELINE.json configs[0] at its real size (the MHC blobs of the reference are
    missing, SURVEY.md §8d replaces them by this generator): 5 Mbp, 200 k + 200 k
    reads, F = 135, R = 36, with control -- the largest case the CPU oracle still
    finishes in seconds.  Whole path through the public classes."""
    from graph_peak_caller_b200 import CallPeaks
    from graph_peak_caller_b200.intervals import UniqueIntervals
    from graph_peak_caller_b200.synthetic import CONFIGS, config_seed
    cfg = CONFIGS["mhc_shaped"]
    R, F = cfg["read_length"], cfg["fragment_length"]
    seed = config_seed(1)
    g = make_graph(cfg["chromosomes"][0][1], cfg["site_rate"], seed=seed)
    s = make_reads(g, cfg["n_sample"], R, F, seed=seed + 1)
    c = make_reads(g, cfg["n_control"], R, F, seed=seed + 2, peak_frac=0.0, dup_frac=0.0)
    lm = LinearMap.from_graph(g)
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, True
    config.linear_map_name = lm
    reporter = Reporter(None)
    caller = CallPeaks(g, config, reporter)
    caller.run(UniqueIntervals(s), UniqueIntervals(c))
    ref = opipeline.callpeaks(g, lm.as_arrays(), opipeline.unique_reads(s),
                              opipeline.unique_reads(c), R, F, True)
    direct = reporter.memory["direct_pileup"]
    assert_exact((direct.indices, direct.values.astype(float)), ref["direct"], "direct pileup")
    q = caller.q_values_pileup
    assert_step_close((q.indices, q.values), ref["q_values"], "q", upto=g.size_bp)
    h = reporter.memory["hole_cleaned"]
    assert_exact((h.indices, h.values.astype(bool)), ref["hole_cleaned"], "holes")
    got = [(p.start_position.offset, p.end_position.offset, list(p.region_paths),
            (bool(p.info[0]), bool(p.info[1])), float(p.score)) for p in caller.max_path_peaks]
    assert len(got) > 50
    assert_same_peaks(got, ref["peaks"])
