"""GPU parity: control (lambda) track, p-values, q-values and threshold through
the C ABI against the oracle."""
import numpy as np
import pytest

from golden_util import GoldenCase, case_names
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import control as ocontrol, pipeline, stats as ostats, tracks
from parity_util import assert_exact, assert_step_close

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.cpu().numpy()


def upos(t):
    return cpu(t).view(np.uint32).astype(np.int64)


def gpu_to_threshold(graph, lm, sample, control, R, F, has_control, unique, global_min=None,
                     q_threshold=0.05):
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    dg = DeviceGraph(graph, lm)
    ds, dc = DeviceReads(sample), DeviceReads(control)
    si, ns, ci, nc = None, len(sample), None, len(control)
    if unique:
        si, ns = kernels.unique_reads(ds)
        ci, nc = kernels.unique_reads(dc)
    events, touched = kernels.sample_events(dg, ds, si, ns, F - R)
    (fi, fv), (di, dv) = kernels.events_to_runs(events, graph.size_bp)
    ext = [F, 1000, 10000] if has_control else [10000]
    min_value = global_min if global_min is not None else nc * F / lm[1][-1]
    bp, val, x = kernels.control_linear_track(dg, dc, ci, nc, ext, F, min_value, want_x=True)
    ratio = ns / nc
    cpos, clam = kernels.control_backproject(dg, touched, bp, val, min_value,
                                             ratio if ratio < 1 else 1.0)
    ppos, p = kernels.pvalues(fi, fv, 1 / ratio if ratio > 1 else 1.0, cpos, clam)
    tp, tc = kernels.ptable_local(ppos, p, graph.size_bp)
    qp, qq = kernels.qtable_build(tp, tc, graph.size_bp)
    q = kernels.qvalues_apply(p, qp, qq)
    tpos, tval = kernels.threshold(ppos, q, -np.log10(q_threshold))
    return dict(x=cpu(x)[:nc], linear=(cpu(bp), cpu(val)), control=(upos(cpos), cpu(clam)),
                p_values=(upos(ppos), cpu(p)), q_values=(upos(ppos), cpu(q)),
                q_table=(cpu(qp), cpu(qq)), ptable=(cpu(tp), cpu(tc)),
                thresholded=(upos(tpos), cpu(tval).astype(bool)), ratio=ratio,
                min_value=min_value)


def oracle_to_threshold(graph, lm, sample, control, R, F, has_control, unique, global_min=None,
                        q_threshold=0.05):
    s, c = sample, control
    if unique:
        s, c = pipeline.unique_reads(s), pipeline.unique_reads(c)
    st = pipeline.run_to_p_values(graph, lm, s, c, R, F, has_control, global_min)
    sp, cum = pipeline.p_table_of(*st["p_values"], st["track_size"])
    table = ostats.q_table(sp, cum)
    q = ostats.q_value_runs(*st["p_values"], table)
    st["q_values"] = q
    st["thresholded"] = tracks.threshold_runs(q[0], q[1], -np.log10(q_threshold))
    st["table"] = table
    return st


def compare(graph, lm, sample, control, R, F, has_control, unique, **kw):
    got = gpu_to_threshold(graph, lm, sample, control, R, F, has_control, unique, **kw)
    ref = oracle_to_threshold(graph, lm, sample, control, R, F, has_control, unique, **kw)
    G = graph.size_bp
    assert np.array_equal(got["x"], ref["x"]), "linear projection must be bit-exact"
    assert got["min_value"] == ref["min_value"]
    assert_step_close(got["linear"], ref["linear"], "linear lambda")
    assert_step_close(got["control"], ref["control"], "graph lambda", upto=G)
    if got["ratio"] <= 1:
        assert_step_close(got["p_values"], ref["p_values"], "p-values", upto=G)
        assert_step_close(got["q_values"], ref["q_values"], "q-values", upto=G)
        assert_exact(got["thresholded"], ref["thresholded"], "thresholded")
    return got, ref


@pytest.mark.parametrize("name", case_names())
def test_golden_cases(name):
    case = GoldenCase(name)
    m = case.meta
    got, ref = compare(case.graph, case.linear_map, case.sample, case.control,
                       m["read_length"], m["fragment_length"], m["has_control"], m["unique"])
    # straight against the reference's own tracks
    G = case.graph.size_bp
    assert_step_close(got["control"], case.track("control"), "control vs reference", upto=G)
    if got["ratio"] <= 1:
        assert_step_close(got["p_values"], case.track("p_values"), "p vs reference", upto=G)
        assert_step_close(got["q_values"], case.track("q_values"), "q vs reference", upto=G)


@pytest.mark.parametrize("trial", range(10))
def test_random_graphs(trial):
    rng = np.random.default_rng(4242 + trial)
    g = make_graph(int(rng.integers(2000, 40000)),
                   site_rate=float(rng.choice([0.01, 0.03, 0.1])), seed=trial + 90,
                   mnp_frac=float(rng.choice([0, 0.3])),
                   indel_mean=float(rng.choice([2, 5, 20])))
    lm = ocontrol.linear_map_from_graph(g)
    R = int(rng.integers(6, 30))
    F = int(rng.integers(R + 5, 150))
    ns = int(rng.integers(100, 2500))
    nc = int(rng.integers(ns, 4000))
    s = make_reads(g, ns, R, F, seed=trial + 700, peak_frac=0.7, peak_spacing=2500)
    c = make_reads(g, nc, R, F, seed=trial + 800, peak_frac=0.0, dup_frac=0.0)
    compare(g, lm, s, c, R, F, bool(trial % 2), unique=True)


def test_global_min_and_more_sample_than_control():
    g = make_graph(20000, 0.03, seed=5, mnp_frac=0.2)
    lm = ocontrol.linear_map_from_graph(g)
    s = make_reads(g, 3000, 12, 60, seed=6, peak_frac=0.6, peak_spacing=3000)
    c = make_reads(g, 900, 12, 60, seed=7, peak_frac=0.0, dup_frac=0.0)
    got, ref = compare(g, lm, s, c, 12, 60, True, True, global_min=0.37)
    assert got["ratio"] > 1
    # with ratio > 1 the reference rescales the sample by float steps; compare
    # p-values where the sample is covered (count > 0), see DESIGN.md
    si, sv = ref["sample"]
    pi, pv = got["p_values"]
    ri, rv = ref["p_values"]
    q = np.union1d(pi, ri)
    q = q[q < g.size_bp]
    from parity_util import step_at
    covered = step_at(si, sv, q) > 0
    a, b = step_at(pi, pv, q)[covered], step_at(ri, rv, q)[covered]
    assert np.all(np.abs(a - b) <= 1e-6 * np.abs(b) + 1e-12)


def test_q_table_matches_dict():
    g = make_graph(30000, 0.03, seed=15)
    lm = ocontrol.linear_map_from_graph(g)
    s = make_reads(g, 2500, 15, 70, seed=16, peak_frac=0.7, peak_spacing=3000)
    c = make_reads(g, 3000, 15, 70, seed=17, peak_frac=0.0, dup_frac=0.0)
    got = gpu_to_threshold(g, lm, s, c, 15, 70, True, True)
    qp, qq = got["q_table"]
    assert np.all(np.diff(qp) < 0), "distinct p, descending"
    assert np.all(np.diff(qq) <= 0), "q is a running minimum"
    tp, tc = got["ptable"]
    # base pairs in the table + base pairs with p == 0 == graph size
    pi, pv = got["p_values"]
    lens = np.diff(np.r_[pi, g.size_bp])
    assert int(tc.sum()) + int(lens[pv == 0].sum()) == g.size_bp
    assert tp.size == np.unique(pv[pv != 0]).size


def test_dense_and_merge_paths_of_the_linear_track_agree(monkeypatch):
    """Integer linear coordinates take the dense shared-memory path, anything
    else the sample-partitioned merge; on integer input both must give the
    same breakpoints and values bit for bit."""
    import os
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    g = make_graph(300000, 0.02, seed=71)                 # SNPs / indels only: integer scales
    lm = ocontrol.linear_map_from_graph(g)
    c = make_reads(g, 40000, 36, 200, seed=72, peak_frac=0.3, peak_spacing=20000, dup_frac=0.1)
    dg, dc = DeviceGraph(g, lm), DeviceReads(c)
    out = {}
    monkeypatch.setenv("GPC_LINEAR_DENSE_MIN", "0")       # the dense path whatever the read density
    for mode in ("dense", "merge"):
        if mode == "merge":
            monkeypatch.setenv("GPC_FORCE_MERGE_PATH", "1")
        bp, val, x = kernels.control_linear_track(dg, dc, None, len(c), [200, 1000, 10000], 200,
                                                  0.9, want_x=True)
        out[mode] = (cpu(bp), cpu(val))
        assert np.all(cpu(x) == np.floor(cpu(x)))
    monkeypatch.delenv("GPC_FORCE_MERGE_PATH")
    assert np.array_equal(out["dense"][0], out["merge"][0])
    assert np.array_equal(out["dense"][1], out["merge"][1])
    assert out["dense"][0].size > 1000
