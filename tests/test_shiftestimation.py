"""CPU: fragment-length estimation -- the native sweeps behind
``PeakModel`` / ``LinearFilter`` (csrc/shift.cu) against the oracle restatement
(oracle/shift.py), both against golden vectors made by the unmodified reference
(tests/golden/shift/shift_model.npz, oracle/make_golden_shift.py) and, in the build
container, against the live reference itself."""
import os
import types

import numpy as np
import pytest

from graph_peak_caller_b200 import callpeaks_interface as ci
from graph_peak_caller_b200 import shiftestimation as se
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import refbridge, shift as oshift
from test_callpeaks_interface import write_graph_json, write_reads_json

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "shift", "shift_model.npz"))


def golden_positions(k):
    return {"+": {c: GOLDEN["m%d_plus_%s" % (k, c)] for c in ("0", "1")},
            "-": {c: GOLDEN["m%d_minus_%s" % (k, c)] for c in ("0", "1")}}


def model_of(positions, genome, lmfold=5, umfold=50):
    opt = se.Opt(lmfold, umfold)
    opt.gsize = genome
    return se.PeakModel(opt, se.Treatment(positions))


def random_positions(rng, genome=600_000, sites=150, chroms=("a", "b", "c"), repeat=0.0):
    d = int(rng.integers(60, 300))
    pos = {"+": {}, "-": {}}
    for chrom in chroms:
        centers = rng.choice(genome - 4000, sites, replace=False) + 2000
        plus, minus = [rng.integers(0, genome, 1500)], [rng.integers(0, genome, 1500)]
        for c in centers:
            plus.append((c - d // 2 + rng.normal(0, 12, int(rng.integers(5, 40)))).astype(np.int64))
            minus.append((c + d // 2 + rng.normal(0, 12, int(rng.integers(5, 40)))).astype(np.int64))
        for s, parts in (("+", plus), ("-", minus)):
            v = np.concatenate(parts)
            if repeat:                                    # repeated tags count once per window summit
                v = np.r_[v, rng.choice(v, int(repeat * v.size))]
            pos[s][chrom] = v
    return pos, genome * len(chroms)


def same_model(got, want):
    assert got.d == want["d"]
    assert np.array_equal(got.plus_line, want["plus_line"])
    assert np.array_equal(got.minus_line, want["minus_line"])
    assert np.array_equal(got.ycorr, want["ycorr"])
    assert got.alternative_d == list(want["alternative_d"])
    assert (got.min_tags, got.max_tags) == (want["min_tags"], want["max_tags"])
    assert {c: v.tolist() for c, v in got.paired_peakpos.items()} == \
        {c: list(v) for c, v in want["centers"].items() if len(v)}


@pytest.mark.parametrize("k", [0, 1])
def test_golden_model(k):
    genome, planted, min_tags, max_tags = GOLDEN["m%d_meta" % k].tolist()
    got = model_of(golden_positions(k), genome)
    assert got.d == GOLDEN["m%d_d" % k][0] and abs(got.d - planted) < 3
    assert np.array_equal(got.plus_line, GOLDEN["m%d_plus_line" % k])
    assert np.array_equal(got.minus_line, GOLDEN["m%d_minus_line" % k])
    assert np.array_equal(got.ycorr, GOLDEN["m%d_ycorr" % k])
    assert got.alternative_d == GOLDEN["m%d_alternative_d" % k].tolist()
    assert (got.min_tags, got.max_tags) == (min_tags, max_tags)
    assert got.scan_window == max(got.d, 10) * 2
    want = oshift.peak_model(golden_positions(k), genome)          # the oracle on the same vectors
    same_model(got, want)
    estimator = se.MultiGraphShiftEstimator(golden_positions(k), genome)
    assert estimator.get_estimates() == round(GOLDEN["m%d_d" % k][0])


@pytest.mark.parametrize("seed,repeat,folds", [(1, 0.0, (5, 50)), (2, 0.5, (5, 50)), (3, 0.1, (3, 20)),
                                               (4, 0.0, (8, 30))])
def test_native_equals_oracle(seed, repeat, folds):
    rng = np.random.default_rng(seed)
    pos, genome = random_positions(rng, repeat=repeat)
    want = oshift.peak_model(pos, genome, *folds)
    if want["d"] is None:
        with pytest.raises(se.NotEnoughPairsException):
            model_of(pos, genome, *folds)
        return
    same_model(model_of(pos, genome, *folds), want)


def test_small_and_degenerate_inputs():
    rng = np.random.default_rng(9)
    pos, genome = random_positions(rng, sites=20, chroms=("x",))
    with pytest.raises(se.NotEnoughPairsException):                 # < 100 pairs
        model_of(pos, genome)
    assert oshift.peak_model(pos, genome)["d"] is None
    # a chromosome with no reads on one strand, one with a single read: skipped
    pos, genome = random_positions(rng, chroms=("a", "b", "c"))
    pos["+"]["e"], pos["-"]["e"] = np.array([], dtype=np.int64), pos["-"]["a"][:500]
    pos["+"]["f"], pos["-"]["f"] = np.array([5]), np.array([7])
    same_model(model_of(pos, genome), oshift.peak_model(pos, genome))
    # the library refuses unsorted tags (Treatment sorts; a direct caller must)
    import ctypes
    from graph_peak_caller_b200 import _lib
    lines = np.zeros((4, 1211), dtype=np.int32)
    tags = np.array([5, 3, 9], dtype=np.int64)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    n = ctypes.c_uint64(0)
    status = _lib.load().gpc_shift_model_host(vp(tags), 3, vp(tags), 3, 600, 10, 1, 5, vp(lines[0]),
                                              vp(lines[1]), vp(lines[2]), vp(lines[3]), None,
                                              ctypes.byref(n))
    assert status == _lib.GPC_ERR_ARG


def test_golden_linear_filter():
    n_bp, gseed, n_reads, rseed = GOLDEN["lf_graph_args"].tolist()
    g = make_graph(n_bp, 0.05, seed=gseed, mnp_frac=0.2)
    reads = make_reads(g, n_reads, 20, 80, seed=rseed, peak_frac=0.5, peak_spacing=2500)
    lf = se.LinearFilter.from_reads_and_graph(reads, g)
    assert (lf._path + g.min_node).tolist() == GOLDEN["lf_path"].tolist()
    found = lf.find_start_positions()
    assert found["+"].tolist() == GOLDEN["lf_plus"].tolist()
    assert found["-"].tolist() == GOLDEN["lf_minus"].tolist()
    assert lf.length() == GOLDEN["lf_length"][0]
    # and the oracle's loops on the same reads
    paths = [reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist() for i in range(len(reads))]
    path = oshift.most_covered_path(g.node_len, g.min_node, g.succ_ptr, g.succ, g.pred_ptr, paths)
    assert path == GOLDEN["lf_path"].tolist()
    want = oshift.start_positions(g.node_len, g.min_node, path,
                                  zip(reads.start_node.tolist(), reads.start_off.tolist()))
    assert want["+"] == GOLDEN["lf_plus"].tolist() and want["-"] == GOLDEN["lf_minus"].tolist()


def test_most_covered_path_rules():
    # ties go to the first successor in adjacency order; zero-count successors are still followed
    g = Graph({1: 3, 2: 2, 3: 2, 4: 5, 5: 1}, {1: [2, 3], 2: [4], 3: [4], 4: [5]})
    assert se.most_covered_path(g, [0, 0, 0, 0, 0]).tolist() == [0, 1, 3, 4]
    assert se.most_covered_path(g, [0, 1, 2, 0, 0]).tolist() == [0, 2, 3, 4]
    assert se.most_covered_path(g, [9, 2, 2, 0, 0]).tolist() == [0, 1, 3, 4]
    with pytest.raises(Exception, match="one start node"):         # haplotyping.py:33
        se.most_covered_path(Graph({1: 3, 2: 2, 3: 2}, {1: [3], 2: [3]}), [0, 0, 0])
    # a read start off the path is dropped; a reverse start counts from the node's far end
    reads = types.SimpleNamespace(start_node=np.array([1, 3, -2, -4, 2], dtype=np.int32),
                                  start_off=np.array([2, 1, 1, 5, 0], dtype=np.int32))
    found = se.LinearFilter(reads, g, [0, 1, 3, 4]).find_start_positions()
    assert found["+"].tolist() == [2, 3] and found["-"].tolist() == [3 + 2 - 1, 5 + 5 - 5]


def test_estimate_from_files(tmp_path):
    # reads with the plus pile F/2 left and the minus pile F/2 right of every site: the
    # estimate from files on disk is F within the jitter, and callpeaks2 uses it when -f is missing
    F, R = 180, 30
    g = make_graph(800000, 0.02, seed=70, mnp_frac=0.1)
    reads = make_reads(g, 50000, R, F, seed=71, peak_frac=0.8, peak_spacing=5000, dup_frac=0.0)
    ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write_graph_json(tmp_path / "1.json", g),
        out_file_name=str(tmp_path / "1.nobg")))
    sample = write_reads_json(tmp_path / "sample_1.json", g, reads)
    estimator = se.MultiGraphShiftEstimator.from_files([str(tmp_path / "1.nobg")], [sample])
    lf = se.LinearFilter.from_reads_and_graph(reads, g)
    assert estimator.genome_size == lf.length()
    d = estimator.get_estimates()
    assert isinstance(d, int) and abs(d - F) <= 12
    assert estimator.peakmodel.n_paired_peaks >= 100
    want = oshift.peak_model({s: {"0": v.tolist()} for s, v in lf.find_start_positions().items()},
                             lf.length())
    assert round(want["d"]) == d
    assert ci.estimate_fragment_length([str(tmp_path / "1.nobg")], [sample],
                                       types.SimpleNamespace(min_fold_enrichment=None,
                                                             max_fold_enrichment="50")) == d
    # run_callpeaks2 without -f and -r: both estimated before the run is set up
    # (callpeaks_interface.py:201-217); -G / -u then use the estimated length
    caller = ci.prepare_callpeaks2(types.SimpleNamespace(
        graph=[str(tmp_path / "1.nobg")], sample=[sample], control=None, fragment_length=None,
        read_length=None, genome_size="800000", unique_reads="40000",
        out_name=str(tmp_path / "est_")))
    assert caller._config.fragment_length == d and caller._config.read_length == R
    assert caller._config.global_min == 40000 * d / 800000
    assert ci.shift_estimation(types.SimpleNamespace(
        chromosomes="1", ob_graphs_location=str(tmp_path),
        sample_reads_base_name=str(tmp_path / "sample_"),
        min_fold_enrichment=None, max_fold_enrichment=None)) == d


@pytest.mark.skipif(not refbridge.available(), reason="/root/reference not present")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_equals_live_reference(seed):
    from oracle.make_golden_shift import reference_linear_filter, reference_model
    rng = np.random.default_rng(seed)
    pos, genome = random_positions(rng, repeat=0.2 * (seed % 2))
    ref = reference_model(pos, genome, 4, 40)
    got = model_of(pos, genome, 4, 40)
    assert got.d == ref.d and got.alternative_d == ref.alternative_d
    assert np.array_equal(got.plus_line, ref.plus_line) and np.array_equal(got.minus_line, ref.minus_line)
    assert np.array_equal(got.ycorr, ref.ycorr) and np.array_equal(got.xcorr, ref.xcorr)
    assert (got.min_tags, got.max_tags, got.scan_window) == (ref.min_tags, ref.max_tags, ref.scan_window)
    same_model(got, oshift.peak_model(pos, genome, 4, 40))
    g = make_graph(12000 + 1000 * seed, 0.06, seed=seed, mnp_frac=0.3)
    reads = make_reads(g, 1500, 20, 80, seed=seed + 1, peak_frac=0.4, peak_spacing=2000)
    path, found, length = reference_linear_filter(g, reads)
    lf = se.LinearFilter.from_reads_and_graph(reads, g)
    assert (lf._path + g.min_node).tolist() == path and lf.length() == length
    mine = lf.find_start_positions()
    assert mine["+"].tolist() == found["+"] and mine["-"].tolist() == found["-"]


# ---- the reference's own known answers ---------------------------------------------------

def test_reference_linearfilter_known_answer():
    # reference tests/test_linearfilter.py:7-42: path 10 (from offset 5), 11, 13 of six
    # 10-bp nodes; starts 10:7, 11:5, -11:5, 13:2 -> {"+": [2, 10, 17], "-": [10]}
    g = Graph({i: 10 for i in (9, 10, 11, 12, 13, 14)},
              {9: [10], 10: [11, 12], 11: [13], 12: [13], 13: [14]})
    reads = types.SimpleNamespace(start_node=np.array([10, 11, -11, 13, 12, -9], dtype=np.int32),
                                  start_off=np.array([7, 5, 5, 2, 1, 3], dtype=np.int32))
    found = se.LinearFilter(reads, g, [1, 2, 4], start_offset=5).find_start_positions()
    assert {k: v.tolist() for k, v in found.items()} == {"+": [2, 10, 17], "-": [10]}
    # the same through the reference's own constructor arguments (positions, indexed interval)
    from graph_peak_caller_b200.peakcollection import Interval, Position
    positions = [Position(10, 7), Position(11, 5), Position(-11, 5), Position(13, 2)]
    interval = Interval(Position(10, 5), Position(13, 5), [10, 11, 13], graph=g)
    found = se.LinearFilter(positions, interval).find_start_positions()
    assert {k: v.tolist() for k, v in found.items()} == {"+": [2, 10, 17], "-": [10]}
    want = oshift.start_positions(g.node_len, 9, [10, 11, 13], [(10, 7), (11, 5), (-11, 5), (13, 2)])
    assert {k: [x - 5 for x in v] for k, v in want.items()} == {"+": [2, 10, 17], "-": [10]}


def test_reference_haplotyper_known_answers():
    # reference tests/test_haplotyping.py:28-66
    from graph_peak_caller_b200.reads import ReadBatch
    from graph_peak_caller_b200.peakcollection import Interval

    def path_of(graph, intervals):
        reads = ReadBatch.from_intervals([Interval(s, e, n) for s, e, n in intervals])
        lf = se.LinearFilter.from_reads_and_graph(reads, graph)
        return (lf._path + graph.min_node).tolist()

    simple = Graph({i: 3 for i in range(1, 5)}, {1: [2, 3], 2: [4], 3: [4]})
    assert path_of(simple, [(0, 3, [1, 3])]) == [1, 3, 4]
    complex_graph = Graph({i: 3 for i in range(1, 13)},
                          {1: [2, 3], 2: [7, 8], 3: [4, 5], 4: [6], 5: [6], 6: [10], 7: [9],
                           8: [9], 9: [10], 10: [12]})
    intervals = [(0, 3, [1, 3, 4, 6, 10]), (1, 2, [2]), (2, 3, [2]), (0, 3, [7, 9])]
    # node 11 has no edges and is not a start node
    assert path_of(complex_graph, intervals) == [1, 2, 7, 9, 10, 12]


def test_reference_multigraph_shift_known_answer():
    # reference tests/test_multigraph_shiftestimation.py:8-33: 400 sites of 40 read pairs,
    # the minus read `shift` to the right of the plus read -> the estimate is the shift
    rng = np.random.default_rng(5)
    shift, genome_size = 150, 10000000
    positions = {"+": {1: []}, "-": {1: []}}
    for _ in range(400):
        pos = int(rng.integers(0, genome_size - shift))
        for _ in range(40):
            pos += int(rng.integers(-3, 3))
            positions["+"][1].append(pos)
            positions["-"][1].append(pos + shift)
    estimator = se.MultiGraphShiftEstimator(positions, genome_size)
    assert round(estimator.get_estimates()) == shift
    assert round(oshift.peak_model(positions, genome_size)["d"]) == shift


def test_sweeps_on_random_small_inputs():
    # the three sweeps of gpc_shift_model_host against the oracle's loops on adversarial small
    # inputs: tiny windows, repeated tags, negative positions, empty strands, any tag bounds
    import ctypes
    from hypothesis import HealthCheck, given, settings, strategies as st
    from graph_peak_caller_b200 import _lib
    lib = _lib.load()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    tags = st.lists(st.integers(-40, 400), max_size=120).map(sorted)

    @settings(derandomize=True, max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(tags, tags, st.integers(1, 40), st.integers(0, 11), st.integers(0, 4), st.integers(0, 30))
    def run(plus, minus, peaksize, expansion, min_tags, more):
        max_tags = min_tags + more
        window = 1 + 2 * peaksize + expansion
        pp = oshift.strand_peaks(plus, peaksize, expansion, min_tags, max_tags)
        mp = oshift.strand_peaks(minus, peaksize, expansion, min_tags, max_tags)
        centers = oshift.pair_centers(pp, mp, peaksize) if pp and mp else []
        want = [[0] * window for _ in range(4)]
        oshift.add_to_line(centers, plus, peaksize, expansion, want[0], want[1])
        oshift.add_to_line(centers, minus, peaksize, expansion, want[2], want[3])
        p, m = np.array(plus, dtype=np.int64), np.array(minus, dtype=np.int64)
        lines = np.zeros((4, window), dtype=np.int32)
        out = np.zeros(len(centers) + 1, dtype=np.int64)
        n = ctypes.c_uint64(out.size)
        status = lib.gpc_shift_model_host(vp(p), p.size, vp(m), m.size, peaksize, expansion, min_tags,
                                          max_tags, vp(lines[0]), vp(lines[1]), vp(lines[2]),
                                          vp(lines[3]), vp(out), ctypes.byref(n))
        assert status == _lib.GPC_OK and n.value == len(centers)
        assert out[:len(centers)].tolist() == centers
        assert lines.tolist() == want
        if len(centers) > 1:              # too little room: reported, lines left alone
            again = np.zeros((4, window), dtype=np.int32)
            n.value = len(centers) - 1
            assert lib.gpc_shift_model_host(vp(p), p.size, vp(m), m.size, peaksize, expansion,
                                            min_tags, max_tags, vp(again[0]), vp(again[1]),
                                            vp(again[2]), vp(again[3]), vp(out),
                                            ctypes.byref(n)) == _lib.GPC_ERR_CAPACITY
            assert n.value == len(centers) and not again.any()

    run()
