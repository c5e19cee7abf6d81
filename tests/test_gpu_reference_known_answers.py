"""GPU: the reference's own known-answer unit tests, run through this
package's classes (same names, same call sequences).  Inputs and expected
outputs are the literal values of
  tests/test_sparseholescleaning.py, tests/test_call_peaks_from_qvalues.py,
  tests/test_pvalues.py, tests/test_create_sample_pileup.py
of the reference."""
import numpy as np
import pytest

from graph_peak_caller_b200 import CallPeaksFromQvalues, Configuration
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.intervals import Intervals
from graph_peak_caller_b200.postprocess import HolesCleaner
from graph_peak_caller_b200.reads import Interval, ReadBatch
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.sample import get_fragment_pileup
from graph_peak_caller_b200.sparsediffs import SparseDiffs, SparseValues
from graph_peak_caller_b200.sparsepvalues import PToQValuesMapper, PValuesFinder

pytestmark = pytest.mark.gpu


def bubbles(first_id):
    i = first_id
    return Graph({i + k: 2 for k in range(10)},
                 {i: [i + 1, i + 2], i + 1: [i + 3], i + 2: [i + 3], i + 3: [i + 4, i + 5],
                  i + 4: [i + 6], i + 5: [i + 6], i + 6: [i + 7, i + 8], i + 8: [i + 9]})


def small_graph():
    return Graph({i: 10 for i in range(101, 107)},
                 {101: [102], 102: [103, 104], 103: [105], 104: [105], 105: [106]})


def sv(indices, values, dtype=bool):
    return SparseValues(np.array(indices, dtype="int"), np.array(values, dtype=dtype))


def same(a, b):
    return np.array_equal(a.indices, b.indices) and \
        np.array_equal(a.values.astype(float), np.asarray(b.values).astype(float))


# ---- tests/test_sparseholescleaning.py -----------------------------------------

def test_end_hole():
    pileup = sv([0, 5, 10], [False, True, False])
    graph = Graph({1: 12}, {})
    assert same(HolesCleaner(graph, pileup, 4).run(), pileup)


def test_long_hole():
    pileup = sv([0, 5, 95], [True, False, True])
    graph = Graph({i: 1 for i in range(1, 101)}, {i: [i + 1] for i in range(1, 100)})
    assert same(HolesCleaner(graph, pileup, 56).run(), pileup)


@pytest.mark.parametrize("first_id", [1, 101])
def test_complicated_and_offset(first_id):
    pileup = sv([0, 1, 2, 4, 6, 8, 10, 12, 14], [1, 0, 1, 0, 1, 0, 1, 0, 1])
    cleaned = HolesCleaner(bubbles(first_id), pileup, 3).run()
    assert same(cleaned, sv([0, 8, 10], [1, 0, 1]))


def test_internals():
    graph = Graph({101: 100}, {})
    pileup = sv([0, 10, 19, 30, 41, 50], [1, 0, 1, 0, 1, 0])
    cleaned = HolesCleaner(graph, pileup, 10).run()
    assert same(cleaned, sv([0, 30, 41, 50], [1, 0, 1, 0]))


def test_touched_edge():
    pileup = sv([0, 4, 22, 28, 56], [1, 0, 1, 0, 1])
    cleaned = HolesCleaner(small_graph(), pileup, 4, set([101, 103, 106])).run()
    assert same(cleaned, pileup)


def test_non_touched_mid():
    pileup = sv([0, 4, 22, 28, 56], [1, 0, 1, 0, 1])
    cleaned = HolesCleaner(small_graph(), pileup, 20, set(range(101, 107))).run()
    assert same(cleaned, sv([0, 30, 40], [1, 0, 1]))


# ---- tests/test_call_peaks_from_qvalues.py --------------------------------------

def q_track(graph, data):
    """tests/util.py convert_old_sparse: {node: (indexes, values, start value)}"""
    indices, values = [0], [0]
    for node in sorted(data):
        idx, val, start = data[node]
        base = int(graph.node_indexes[node - graph.min_node])
        indices.append(base)
        values.append(start)
        indices.extend(int(i) + base for i in idx)
        values.extend(val)
    track = SparseValues(np.array(indices, dtype="int"), np.array(values, dtype=float),
                         sanitize=True)
    track.track_size = int(graph.node_indexes[-1])
    return track


def run_caller(graph, data):
    config = Configuration()
    config.fragment_length, config.read_length = 6, 2
    caller = CallPeaksFromQvalues(graph, q_track(graph, data), config, Reporter(None),
                                  cutoff=0.1, q_values_max_path=True)
    caller.callpeaks()
    return caller.max_paths


def assert_finds(expected, graph, data):
    found = run_caller(graph, data)
    got = sorted((p.start_position.offset, p.end_position.offset, tuple(p.region_paths))
                 for p in found)
    want = sorted((s, e, tuple(rp)) for s, e, rp in expected)
    assert got == want


LINEAR = Graph({i: 10 for i in range(1, 5)}, {i: [i + 1] for i in range(1, 4)})
SPLIT = Graph({i: 10 for i in range(1, 5)}, {1: [2, 3], 2: [4], 3: [4]})
SINGLE = Graph({1: 20}, {})
MULTI = Graph({i: 10 for i in range(1, 6)}, {1: [3], 2: [3], 3: [4, 5]})


def test_finds_single_peak():
    assert_finds([(5, 3, [1, 2])], LINEAR, {1: ([5], [2], 0), 2: ([3], [0], 2)})


def test_finds_single_peak_with_hole():
    assert_finds([(5, 3, [1, 2])], LINEAR, {1: ([5, 8], [2, 0], 0), 2: ([3], [0], 2)})


def test_finds_no_peak_too_large_hole():
    assert_finds([], LINEAR, {1: ([5, 7], [2, 0], 0), 2: ([3], [0], 2)})


def test_find_max_path_on_split_graph():
    assert_finds([(0, 1, [1, 2, 4]), (4, 10, [4])], SPLIT,
                 {1: ([], [], 2), 2: ([], [], 3), 3: ([], [], 2), 4: ([1, 4], [0, 3], 2)})


def test_find_max_path_on_split_graph_with_hole_fill():
    assert_finds([(3, 10, [1, 2, 4])], SPLIT,
                 {1: ([3], [2], 0), 2: ([], [], 2.1), 3: ([], [], 2), 4: ([1, 3], [0, 3], 2)})


def test_removes_too_small_peak():
    assert_finds([], SINGLE, {1: ([5, 9], [2, 0], 0)})


def test_finds_internal_peak_with_internal_hole():
    assert_finds([(5, 15, [1])], SINGLE, {1: ([5, 8, 9, 15], [2, 0, 2, 0], 0)})


def test_finds_correct_max_path_among_many_paths():
    graph = Graph({i: 10 for i in range(1, 6)}, {1: [2, 3, 4], 2: [5], 4: [5], 3: [5]})
    assert_finds([(0, 10, [1, 4, 5])], graph,
                 {1: ([], [], 2), 2: ([1, 2, 7, 8], [0, 2.001, 0, 2.001], 2),
                  3: ([], [], 1.5), 4: ([], [], 2), 5: ([], [], 2)})


def test_multiple_start_and_end_nodes():
    assert_finds([(0, 10, [2, 3, 4]), (3, 10, [5])], MULTI,
                 {1: ([], [], 2), 2: ([], [], 2.2), 3: ([1, 9], [2, 0], 0), 4: ([], [], 2),
                  5: ([3], [3], 0)})


def test_multiple_start_and_end_nodes2():
    assert_finds([(0, 10, [2, 3, 5])], MULTI,
                 {1: ([], [], 2), 2: ([], [], 2.2), 3: ([1, 9], [2, 0], 0), 4: ([], [], 2),
                  5: ([1], [3], 0)})


# ---- tests/test_pvalues.py -------------------------------------------------------

def test_p_to_q_simple_conversion():
    mapping = PToQValuesMapper([2., 1., 0.5], [2, 3, 6]).get_p_to_q_values()
    assert mapping[2] == pytest.approx(2 + (np.log10(1) - np.log10(6)))
    assert mapping[1] == pytest.approx(1 + (np.log10(3) - np.log10(6)))
    assert mapping[0.5] == pytest.approx(0.5 + (np.log10(4) - np.log10(6)))
    assert mapping[0] == 0


def covered(graph, intervals):
    """tests/util.py from_intervals: closed pileup of the given intervals."""
    idx, dif = [], []
    for start, end, nodes in intervals:
        first, last = nodes[0] - graph.min_node, nodes[-1] - graph.min_node
        idx += [int(graph.node_indexes[first]) + start, int(graph.node_indexes[last]) + end]
        dif += [1.0, -1.0]
    order = np.argsort(idx, kind="stable")
    return SparseDiffs(np.array(idx)[order], np.array(dif)[order])


@pytest.mark.parametrize("sample,expected", [
    ([(0, 3, [1, 2])], ([0, 6], [-np.log10(0.26424), 0])),
    ([(0, 3, [2])], ([0, 3, 6], [0, -np.log10(0.26424), 0])),
    ([(0, 3, [2]), (0, 3, [2])], ([0, 3, 6], [0, -np.log10(0.08030), 0])),
])
def test_p_values_finder(sample, expected):
    graph = Graph({1: 3, 2: 3}, {1: [2]})
    finder = PValuesFinder(covered(graph, sample), covered(graph, [(0, 3, [1, 2])]))
    p = finder.get_p_values_pileup()
    assert np.array_equal(p.indices, expected[0])
    assert np.allclose(p.values, expected[1], rtol=1e-4)


# ---- tests/test_create_sample_pileup.py ------------------------------------------

def fragment_reads(graph, fragments, read_length, fragment_length):
    reads = []
    for f in fragments:
        f.graph = graph
        reads.append(f.get_subinterval(0, read_length))
        reads.append(f.get_subinterval(fragment_length - read_length,
                                       fragment_length).get_reverse())
    return ReadBatch.from_intervals(reads)


def pileup_of(graph, fragments, read_length, fragment_length):
    config = Configuration()
    config.fragment_length, config.read_length = fragment_length, read_length
    return get_fragment_pileup(graph, Intervals(fragment_reads(
        graph, fragments, read_length, fragment_length)), config).get_sparse_values()


def test_linear_graph_single_fragment():
    graph = Graph({1: 5, 2: 5}, {1: [2]})
    got = pileup_of(graph, [Interval(3, 3, [1, 2])], 2, 5)
    assert np.array_equal(got.indices, [0, 3, 8]) and np.array_equal(got.values, [0, 2, 0])


def test_linear_graph_two_fragments():
    graph = Graph({1: 5, 2: 5}, {1: [2]})
    got = pileup_of(graph, [Interval(0, 5, [1]), Interval(3, 3, [1, 2])], 2, 5)
    assert np.array_equal(got.indices, [0, 3, 5, 8]) and np.array_equal(got.values, [2, 4, 2, 0])


def test_split_graph_extension_counts_each_read_once():
    """TestSplitGraph: a read reaching a node over two branches counts once."""
    graph = Graph({1: 5, 2: 5, 3: 5, 4: 5}, {1: [2, 3], 2: [4], 3: [4]})
    reads = ReadBatch.from_intervals([Interval(3, 5, [1], graph=graph)])
    config = Configuration()
    config.fragment_length, config.read_length = 9, 2      # extension 7: 5 in 2|3, then 2 in 4
    got = get_fragment_pileup(graph, Intervals(reads), config).get_sparse_values()
    # covers 1:[3,5), all of 2 and 3, and 4:[0,2)
    assert np.array_equal(got.indices, [0, 3, 17]) and np.array_equal(got.values, [0, 1, 0])
