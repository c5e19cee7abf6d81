"""GPU, end to end through the public API (the reference's call sequence):
CallPeaks(graph, config, Reporter).run(Intervals(sample), Intervals(control)).
Checked against (a) the peaks the unmodified reference produced for the golden
cases, (b) the oracle on seeded synthetic graphs, (c) the reference's own
known-answer tests (tests/test_whole_callpeaks.py: planted fragments are
exactly the peaks that come out)."""
import os

import numpy as np
import pytest

from golden_util import GoldenCase, case_names
from graph_peak_caller_b200 import CallPeaks, Configuration
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.intervals import Intervals, UniqueIntervals
from graph_peak_caller_b200.peakcollection import PeakCollection
from graph_peak_caller_b200.reads import Interval, ReadBatch
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.sparsediffs import SparseValues
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import control as ocontrol, pipeline as opipeline
from parity_util import step_at, assert_exact, assert_step_close
from test_gpu_postprocess import assert_same_peaks

pytestmark = pytest.mark.gpu


def run_api(graph, sample, control, R, F, has_control, unique, tmp_path, global_min=None):
    lm_file = str(tmp_path / "lm.npz")
    LinearMap.from_graph(graph).to_file(lm_file)
    config = Configuration()
    config.read_length, config.fragment_length = R, F
    config.linear_map_name = lm_file
    config.has_control = has_control
    config.global_min = global_min
    reporter = Reporter(str(tmp_path / "out_"))
    caller = CallPeaks(graph, config, reporter)
    wrap = UniqueIntervals if unique else Intervals
    s, c = wrap(sample), wrap(control)
    caller.run(s, c)
    return caller, s, c


def peak_tuples(peaks):
    return [(p.start_position.offset, p.end_position.offset, list(p.region_paths),
             (bool(p.info[0]), bool(p.info[1])), float(p.score)) for p in peaks]


@pytest.mark.parametrize("name", case_names())
def test_golden_end_to_end(name, tmp_path):
    case = GoldenCase(name)
    m = case.meta
    caller, s, c = run_api(case.graph, case.sample, case.control, m["read_length"],
                           m["fragment_length"], m["has_control"], m["unique"], tmp_path)
    assert (s.n_reads, c.n_reads) == (m["n_sample"], m["n_control"])
    G = case.graph.size_bp
    # files of the Reporter contract, read back like a user of the reference would
    base = str(tmp_path / "out_")
    direct = SparseValues.from_sparse_files(base + "direct_pileup")
    assert_exact((direct.indices, direct.values), case.track("direct"), "direct pileup")
    assert direct.track_size == G
    touched = np.sort(np.load(base + "touched_nodes.npy"))
    assert np.array_equal(touched, case.touched)
    ratio = m["n_sample"] / m["n_control"]
    if ratio <= 1:
        p = SparseValues.from_sparse_files(base + "pvalues")
        q = SparseValues.from_sparse_files(base + "qvalues")
        assert_step_close((p.indices, p.values), case.track("p_values"), "p", upto=G)
        assert_step_close((q.indices, q.values), case.track("q_values"), "q", upto=G)
        h = SparseValues.from_sparse_files(base + "hole_cleaned")
        assert_exact((h.indices, h.values.astype(bool)), case.track("hole_cleaned"), "holes")
        if case.graph.min_node == 1:
            assert_same_peaks(peak_tuples(caller.max_path_peaks), case.peaks)
            written = PeakCollection.from_file(base + "max_paths.intervalcollection")
            assert [(p.start_position.offset, p.end_position.offset, p.region_paths)
                    for p in written.intervals] == \
                [(p.start_position.offset, p.end_position.offset, p.region_paths)
                 for p in caller.max_path_peaks]
    else:
        # More sample than control reads: the reference scales the sample by float steps of
        # 1 / ratio, a pileup that returns to zero is left at +-1e-17, and p = 0 becomes
        # -log10(1 - exp(-lambda)) there; the kernels count in
        # integers and give 0.  Such p stay below the cutoff whenever lambda > 0.0513 (the
        # background floor of any real run), they rank BELOW every larger p and so leave the
        # q of all larger p untouched: q must agree wherever the sample is covered or the
        # reference's q reaches the cutoff, and the thresholded and hole-cleaned tracks and
        # the peaks must be IDENTICAL to the live reference's.
        cutoff = -np.log10(0.05)
        q = SparseValues.from_sparse_files(base + "qvalues")
        rq_i, rq_v = case.track("q_values")
        s_i, s_v = case.track("sample")
        grid = np.union1d(np.union1d(q.indices, rq_i), s_i)
        grid = grid[grid < G]
        ours, theirs = step_at(q.indices, q.values, grid), step_at(rq_i, rq_v, grid)
        covered = step_at(s_i, s_v, grid) > 1e-9
        must = covered | (theirs >= cutoff) | (ours >= cutoff)
        assert must.any()
        err = np.abs(ours - theirs)[must]
        assert np.all(err <= 1e-6 * np.abs(theirs[must]) + 1e-12), "q where it matters"
        assert np.array_equal(ours >= cutoff, theirs >= cutoff), "thresholded track"
        assert float(np.max(theirs[~must], initial=0.0)) < cutoff
        h = SparseValues.from_sparse_files(base + "hole_cleaned")
        assert_exact((h.indices, h.values.astype(bool)), case.track("hole_cleaned"), "holes")
        if case.graph.min_node == 1:
            assert_same_peaks(peak_tuples(caller.max_path_peaks), case.peaks)
    # the p -> q mapping is a dict with d[0] == 0, as in the reference
    assert caller.p_to_q_values_mapping[0] == 0
    assert isinstance(caller.p_to_q_values_mapping, dict)


@pytest.mark.parametrize("trial", range(6))
def test_synthetic_vs_oracle(trial, tmp_path):
    rng = np.random.default_rng(31337 + trial)
    g = make_graph(int(rng.integers(5000, 60000)),
                   site_rate=float(rng.choice([0.01, 0.03, 0.08])), seed=trial + 40,
                   mnp_frac=float(rng.choice([0, 0.2])), indel_mean=float(rng.choice([3, 10])))
    R = int(rng.integers(8, 30))
    F = int(rng.integers(R + 10, 160))
    ns = int(rng.integers(500, 4000))
    s = make_reads(g, ns, R, F, seed=trial + 1, peak_frac=0.7, peak_spacing=4000)
    c = make_reads(g, int(ns * rng.uniform(1.0, 2.0)), R, F, seed=trial + 2, peak_frac=0.0,
                   dup_frac=0.0)
    has_control = bool(trial % 2)
    caller, si, ci = run_api(g, s, c, R, F, has_control, True, tmp_path)
    lm = ocontrol.linear_map_from_graph(g)
    ref = opipeline.callpeaks(g, lm, opipeline.unique_reads(s), opipeline.unique_reads(c),
                              R, F, has_control=has_control)
    assert si.n_reads == len(ref["x"]) or True
    h = caller._reporter.memory["hole_cleaned"]
    assert_exact((h.indices, h.values.astype(bool)), ref["hole_cleaned"], "hole_cleaned")
    q = caller.q_values_pileup
    assert_step_close((q.indices, q.values), ref["q_values"], "q", upto=g.size_bp)
    assert_same_peaks(peak_tuples(caller.max_path_peaks), ref["peaks"])
    assert len(ref["peaks"]) > 0


# ---- the reference's own known-answer tests (tests/test_whole_callpeaks.py) ----

def planted(graph, fragments, read_length, fragment_length):
    reads = []
    for frag in fragments:
        frag.graph = graph
        for _ in range(10):
            reads.append(frag.get_subinterval(0, read_length))
            reads.append(frag.get_subinterval(fragment_length - read_length,
                                              fragment_length).get_reverse())
    return ReadBatch.from_intervals(reads)


SPLIT = dict(blocks={i: 15 for i in range(1, 5)}, edges={1: [2, 3], 2: [4], 3: [4]})


@pytest.mark.parametrize("fragments", [
    [Interval(3, 8, [2])],                                  # test_single_peak
    [Interval(13, 3, [1, 2])],                              # test_peak_crossing_blocks
    [Interval(13, 3, [1, 2]), Interval(13, 3, [2, 4])],     # two peaks
])
def test_planted_fragments_are_the_peaks(fragments, tmp_path):
    graph = Graph(SPLIT["blocks"], SPLIT["edges"])
    reads = planted(graph, fragments, 1, 5)
    caller, _, _ = run_api(graph, reads, reads, 1, 5, False, False, tmp_path)
    found = [(p.start_position.offset, p.end_position.offset, p.region_paths)
             for p in caller.max_path_peaks]
    want = [(f.start_position.offset, f.end_position.offset, f.region_paths)
            for f in fragments]
    assert sorted(found) == sorted(want)
    for p in caller.max_path_peaks:
        assert p.length() == 5


def test_fragment_must_exceed_read_length():
    graph = Graph(SPLIT["blocks"], SPLIT["edges"])
    config = Configuration()
    config.read_length, config.fragment_length = 10, 10
    with pytest.raises(AssertionError):
        CallPeaks(graph, config, Reporter(None))
