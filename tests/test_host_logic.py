"""CPU: host-side containers and the reference-facing value types."""
import json

import numpy as np
import pytest

from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.intervals import Intervals, UniqueIntervals
from graph_peak_caller_b200.peakcollection import Peak, PeakCollection
from graph_peak_caller_b200.reads import Interval, Position, ReadBatch
from graph_peak_caller_b200.sparsediffs import SparseDiffs, SparseValues
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from graph_peak_caller_b200 import multigraph
from oracle import control as ocontrol, sample as osample, tracks as otracks


def split_graph():
    return Graph({i: 15 for i in range(1, 5)}, {1: [2, 3], 2: [4], 3: [4]})


def test_graph_obg_surface():
    g = split_graph()
    assert list(g.node_indexes) == [0, 15, 30, 45, 60]
    assert g.min_node == 1 and g.node_size(-3) == 15 and len(g.blocks) == 4
    assert g.adj_list[1] == [2, 3] and g.adj_list[4] == []
    assert sorted(g.reverse_adj_list[-4]) == [-3, -2]
    assert list(g.get_topological_sorted_node_ids()) == [1, 2, 3, 4]
    assert 3 in g.blocks and 7 not in g.blocks and g.uses_numpy_backend


def test_graph_with_descending_edges_is_marked_and_refused_by_the_ordered_stages():
    """Flat edge lists (parsers, generator) must ascend; the obg-style constructor takes any
    graph and marks it (only the sample pileup walks such a graph, tests/test_cyclic.py)."""
    with pytest.raises(ValueError):
        Graph.from_edges([5, 5], [1], [0])
    g = Graph({1: 5, 2: 5}, {2: [1]})
    assert not g.is_ordered
    with pytest.raises(ValueError, match="LinearMap"):
        LinearMap.from_graph(g)
    assert Graph({1: 5, 2: 5}, {1: [2]}).is_ordered


def test_graph_file_round_trip(tmp_path):
    g = make_graph(5000, 0.05, seed=3, min_node=17)
    g.to_file(str(tmp_path / "g.npz"))
    h = Graph.from_file(str(tmp_path / "g.npz"))
    assert np.array_equal(g.node_off, h.node_off) and np.array_equal(g.succ, h.succ)
    assert np.array_equal(g.pred_ptr, h.pred_ptr) and h.min_node == 17


def test_interval_geometry_matches_reference_conventions():
    g = Graph({1: 5, 2: 5}, {1: [2]})
    frag = Interval(3, 3, [1, 2], graph=g)
    assert frag.length() == 5
    left = frag.get_subinterval(0, 2)
    assert (left.start_position, left.end_position, left.region_paths) == \
        (Position(1, 3), Position(1, 5), [1])
    right = frag.get_subinterval(3, 5).get_reverse()
    assert right.region_paths == [-2]
    assert (right.start_position, right.end_position) == (Position(-2, 2), Position(-2, 4))
    assert Interval.from_file_line(frag.to_file_line(), g) == frag


def test_read_batch_round_trip_and_select():
    g = make_graph(3000, 0.05, seed=1)
    r = make_reads(g, 200, 12, 40, seed=2)
    again = ReadBatch.from_intervals(list(r))
    for f in ("start_node", "start_off", "end_node", "end_off", "path_ptr", "path"):
        assert np.array_equal(getattr(r, f), getattr(again, f))
    keep = osample.unique_read_mask(r)
    sub = r.select(keep)
    assert len(sub) == int(keep.sum())
    assert [iv.region_paths for iv in sub] == [iv.region_paths for iv, k in zip(r, keep) if k]


def test_unique_intervals_host_view_counts_like_the_reference():
    # reference tests/test_filter_duplicates.py: same start position == duplicate
    reads = [Interval(Position(1, 4), Position(1, 10), [1]),
             Interval(Position(1, 4), Position(2, 3), [1, 2]),
             Interval(Position(1, 5), Position(1, 10), [1]),
             Interval(Position(-1, 4), Position(-1, 10), [-1])]
    u = UniqueIntervals(reads)
    assert len(list(u)) == 3 and u.n_reads == 3 and u.n_duplicates == 1
    i = Intervals(reads)
    assert len(list(i)) == 4 and i.n_reads == 4


def test_sparse_values_sanitize_threshold_dense():
    sv = SparseValues([0, 3, 3, 7, 9], [1.0, 2.0, 5.0, 5.0, 0.0], sanitize=True)
    assert list(sv.indices) == [0, 3, 9] and list(sv.values) == [1.0, 5.0, 0.0]
    sv.track_size = 12
    t = sv.threshold_copy(2.0)
    assert list(t.indices) == [0, 3, 9] and list(t.values) == [False, True, False]
    dense = sv.to_dense_pileup(12)
    assert list(dense) == [1, 1, 1, 5, 5, 5, 5, 5, 5, 0, 0, 0]
    back = SparseValues.from_dense_pileup(dense)
    assert back == sv


def test_sparse_values_file_contract(tmp_path):
    sv = SparseValues([0, 4, 8], [0.5, 1.5, 0.0])
    sv.track_size = 20
    base = str(tmp_path / "x_pvalues")
    sv.to_sparse_files(base)
    assert list(np.load(base + "_indexes.npy")) == [0, 4, 8, 20]     # size appended
    assert SparseValues.from_sparse_files(base) == sv


def test_sparse_diffs_host_ops_match_oracle():
    rng = np.random.default_rng(0)
    a = SparseDiffs(np.sort(rng.integers(0, 50, 20)), rng.normal(size=20))
    b = SparseDiffs(np.sort(rng.integers(0, 50, 15)), rng.normal(size=15))
    m = a.maximum(b)
    oi, od = otracks.maximum_diffs(a._indices, a._diffs, b._indices, b._diffs)
    assert np.array_equal(m._indices, oi) and np.array_equal(m._diffs, od)
    sv = a.get_sparse_values()
    ri, rv = otracks.diffs_to_runs(a._indices, a._diffs)
    assert np.array_equal(sv.indices, ri) and np.array_equal(sv.values, rv)
    a *= 2.0
    assert np.allclose(a.get_sparse_values().values, 2 * rv)


def test_linear_map_matches_oracle_and_file_round_trip(tmp_path):
    g = make_graph(20000, 0.05, seed=3, mnp_frac=0.3)
    lm = LinearMap.from_graph(g)
    os_, oe = ocontrol.linear_map_from_graph(g)
    assert np.array_equal(lm._node_starts, os_) and np.array_equal(lm._node_ends, oe)
    lm.to_file(str(tmp_path / "lm.npz"))
    assert LinearMap.from_file(str(tmp_path / "lm.npz"), g) == lm
    # reference tests/test_linearmap.py: scale of a short branch next to a long one
    g2 = Graph({1: 10, 2: 11, 3: 26, 4: 5}, {1: [2, 3], 2: [4], 3: [4]})
    lm2 = LinearMap.from_graph(g2)
    scale, offset = lm2.get_scale_and_offset(2)
    assert scale == 26 / 11 and offset == 10.0


def test_peak_file_line_contract():
    g = split_graph()
    p = Peak(3, 8, [2], graph=g, score=4.5)
    p.info = (True, False)
    p.chromosome = "chr1"
    o = json.loads(p.to_file_line())
    assert o == {"start": 3, "end": 8, "region_paths": [2], "direction": 1,
                 "average_q_value": 4.5, "unique_id": "None", "is_ambigous": 0,
                 "is_diff": 1, "chromosome": "chr1"}
    q = Peak.from_file_line(p.to_file_line(), g)
    assert q == p and q.score == 4.5 and q.length() == 5
    with pytest.raises(AssertionError):
        p.set_score(3)


def test_peak_collection_file(tmp_path):
    g = split_graph()
    peaks = [Peak(3, 8, [2], graph=g, score=1.0), Peak(13, 3, [1, 3], graph=g, score=2.0)]
    name = str(tmp_path / "p.intervalcollection")
    PeakCollection(peaks).to_file(name, text_file=True)
    back = PeakCollection.from_file(name, graph=g)
    assert back.intervals == peaks


def test_chromosome_bin_packing():
    hg19 = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90,
            81, 78, 59, 63, 48, 51, 155, 59]
    bins = multigraph.assign_chromosomes(hg19, 8)
    assert sorted(i for b in bins for i in b) == list(range(24))
    loads = [sum(hg19[i] for i in b) for b in bins]
    assert max(loads) <= 1.1 * sum(hg19) / 8


def test_peak_from_flat_builds_positions_on_demand():
    """Peaks made from the kernels' flat output behave like constructed ones."""
    from graph_peak_caller_b200.peakcollection import Peak
    from graph_peak_caller_b200.reads import Position
    g = make_graph(500, 0.05, seed=1)
    p = Peak.from_flat(3, 2, [1, 2, 4], g, 4.5, (True, False), 17)
    assert "_sp" not in p.__dict__                       # nothing built yet
    assert p.start_position == Position(1, 3) and p.end_position == Position(4, 2)
    assert p.length() == 17 and p.score == 4.5 and p.info == (True, False)
    assert p.region_paths == [1, 2, 4] and p.direction == 1 and p.chromosome is None
    same = Peak(Position(1, 3), Position(4, 2), [1, 2, 4], g, score=4.5)
    assert json.loads(p.to_file_line())["region_paths"] == json.loads(same.to_file_line())["region_paths"]
    q = Peak.from_file_line(p.to_file_line(), g)
    assert q.start_position == p.start_position and q.end_position == p.end_position
    p.start_position = Position(1, 0)                    # still assignable, as on an Interval
    assert p.start_position.offset == 0


def test_compact_wire_form_of_a_read_batch():
    """ReadBatch.compact(): offsets packed into 16 bits each, path lengths as bytes; batches
    that do not fit (long nodes, long paths, a start node that is not the first path node)
    report None and go up as the six arrays."""
    g = make_graph(20000, 0.05, seed=5)
    r = make_reads(g, 500, 20, 80, seed=6)
    offs, plen, path = r.compact()
    assert np.array_equal(offs & 0xffff, r.start_off) and np.array_equal(offs >> 16, r.end_off)
    assert np.array_equal(plen, np.diff(r.path_ptr)) and path is r.path
    far = ReadBatch(r.start_node, r.start_off + 70000, r.end_node, r.end_off, r.path_ptr, r.path)
    assert far.compact() is None
    other = r.start_node.copy()
    other[0] += 1
    assert ReadBatch(other, r.start_off, r.end_node, r.end_off, r.path_ptr, r.path).compact() is None
    long_path = ReadBatch([1], [0], [1], [1], [0, 300], np.ones(300, dtype=np.int32))
    assert long_path.compact() is None
    # delta-coded paths: first node + int8 steps decode back to the path array
    offs2, plen2, first, step = r.compact_delta()
    assert np.array_equal(offs2, offs) and np.array_equal(plen2, plen)
    assert first.dtype == np.int32 and step.dtype == np.int8
    assert step.size == r.path.size - len(r)
    rebuilt, k = [], 0
    for i in range(len(r)):
        v = int(first[i])
        rebuilt.append(v)
        for _ in range(int(plen[i]) - 1):
            v += int(step[k])
            k += 1
            rebuilt.append(v)
    assert rebuilt == r.path.tolist()
    # a step that does not fit int8: no delta form, the plain compact form still applies
    jump = ReadBatch([1, 5], [0, 0], [400, 6], [1, 1], [0, 2, 4], np.array([1, 400, 5, 6], dtype=np.int32))
    assert jump.compact() is not None and jump.compact_delta() is None
