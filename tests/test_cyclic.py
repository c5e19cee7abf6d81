"""Sample pileup on graphs with cycles (SURVEY.md §8a row S-cyc).

CPU: oracle/cyclic.py against the reference's own known answers
(tests/legacy/test_reversal.py:49-70 on the graphs of tests/cyclic_graph.py:6-22) and,
on acyclic graphs, against the pinned oracle of the live path (the two semantics agree
where no cycle lets an extension run back into its own read).
GPU: the worklist kernel (gpc_sample_pileup on a graph flagged GPC_GRAPH_UNORDERED)
against oracle/cyclic.py, both pileup routes, bit-exact."""
import numpy as np
import pytest

from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.reads import Interval, Position, ReadBatch
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import cyclic as ocyclic, sample as osample


def small_cyclic():         # tests/cyclic_graph.py:6-10
    return Graph({1: 100}, {1: [1]})


def large_cyclic():         # tests/cyclic_graph.py:13-16
    return Graph({1: 100, 2: 20}, {1: [2], 2: [1]})


def padded_cyclic():        # tests/cyclic_graph.py:19-23
    return Graph({i: 100 for i in range(1, 5)}, {1: [2], 2: [3, 4, 2]})


def batch(graph, intervals):
    for iv in intervals:
        iv.graph = graph
    return ReadBatch.from_intervals(intervals)


def covered(graph, reads, extension):
    """{node id: sorted covered offsets} of the first read, forward coordinates."""
    body, ext, fwd = ocyclic.read_areas(graph, reads, 0, extension)
    out = {}
    for v, a, b in body + ext:
        n = int(graph.node_len[v])
        span = range(a, b) if fwd else range(n - b, n - a)
        out.setdefault(v + graph.min_node, set()).update(span)
    return {k: sorted(v) for k, v in out.items()}


def test_graphs_with_cycles_are_marked_not_refused():
    assert not small_cyclic().is_ordered and not large_cyclic().is_ordered
    assert Graph({1: 5, 2: 5}, {1: [2]}).is_ordered
    with pytest.raises(ValueError, match="ascend"):
        Graph.from_edges([5, 5], [1], [0])
    from graph_peak_caller_b200.control.linearmap import LinearMap
    with pytest.raises(ValueError, match="LinearMap"):
        LinearMap.from_graph(large_cyclic())


def test_reference_known_answer_self_loop():
    """test_reversal.py:49-57: node of 100 bp with a self loop, read 70-90, fragment
    length 40 -> add_start(1, 10) and add_start(-1, 30)."""
    g = small_cyclic()
    got = covered(g, batch(g, [Interval(70, 90, [1])]), 40 - 20)
    assert got == {1: list(range(0, 10)) + list(range(70, 100))}


def test_reference_known_answer_two_node_cycle():
    """test_reversal.py:59-70: 1 (100 bp) <-> 2 (20 bp), read from 1:90 to 2:10, fragment
    length 40 -> add_start(1, 10), add_start(-1, 10), add_start(2, 20)."""
    g = large_cyclic()
    got = covered(g, batch(g, [Interval(90, 10, [1, 2])]), 40 - 20)
    assert got == {1: list(range(0, 10)) + list(range(90, 100)), 2: list(range(20))}


def test_extension_running_into_its_own_read_counts_once():
    g = small_cyclic()
    reads = batch(g, [Interval(10, 30, [1])])
    out = ocyclic.fragment_pileup(g, reads, 150)         # wraps once around and over the read
    idx, val = out["sample"]
    assert np.array_equal(idx, [0, 100]) and np.array_equal(val, [1, 0])     # every base once
    d_idx, d_val = out["direct"]
    assert np.array_equal(d_idx, [0, 10, 30]) and np.array_equal(d_val, [0, 1, 0])


@pytest.mark.parametrize("trial", range(6))
def test_agrees_with_the_live_path_oracle_on_acyclic_graphs(trial):
    rng = np.random.default_rng(500 + trial)
    g = make_graph(int(rng.integers(400, 3000)), float(rng.choice([0.03, 0.1])), seed=600 + trial,
                   mnp_frac=0.2, indel_mean=6.0)
    R = int(rng.integers(5, 20))
    ext = int(rng.integers(1, 60))
    r = make_reads(g, 150, R, R + ext, seed=700 + trial, dup_frac=0.0)
    a, b = ocyclic.fragment_pileup(g, r, ext), osample.fragment_pileup(g, r, ext)
    for key in ("sample", "direct"):
        assert np.array_equal(a[key][0], b[key][0]) and np.array_equal(a[key][1], b[key][1]), key
    assert np.array_equal(a["touched"], b["touched"])


def random_cyclic_case(seed):
    """A bubble graph with back edges, self loops and reads on both strands."""
    rng = np.random.default_rng(seed)
    g0 = make_graph(int(rng.integers(300, 1500)), 0.08, seed=seed, mnp_frac=0.3, indel_mean=5.0)
    n = g0.n_nodes
    src = np.repeat(np.arange(n), np.diff(g0.succ_ptr))
    dst = g0.succ.astype(np.int64)
    k = int(rng.integers(2, 8))
    back_src = rng.integers(1, n, size=k)
    back_dst = np.array([int(rng.integers(max(0, s - 6), s + 1)) for s in back_src])   # s -> s is a self loop
    g = Graph.from_edges(g0.node_len, np.r_[src, back_src], np.r_[dst, back_dst], g0.min_node,
                         allow_unordered=True)
    R = int(rng.integers(4, 12))
    ext = int(rng.integers(5, 50))
    r = make_reads(g0, 120, R, R + ext, seed=seed + 1, dup_frac=0.0)       # walks of the forward graph
    return g, r, ext


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["self_loop", "two_node", "padded"] + list(range(8)))
def test_worklist_kernel_vs_oracle(case):
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
    if case == "self_loop":
        g, ext = small_cyclic(), 20
        r = batch(g, [Interval(70, 90, [1]), Interval(10, 30, [1]),
                      Interval(Position(-1, 20), Position(-1, 45), [-1])])
    elif case == "two_node":
        g, ext = large_cyclic(), 20
        r = batch(g, [Interval(90, 10, [1, 2]), Interval(Position(-2, 5), Position(-1, 30), [-2, -1]),
                      Interval(5, 15, [2])])
    elif case == "padded":
        g, ext = padded_cyclic(), 170
        r = batch(g, [Interval(50, 20, [1, 2]), Interval(80, 95, [2]), Interval(10, 60, [4]),
                      Interval(Position(-3, 10), Position(-2, 40), [-3, -2])])
    else:
        g, r, ext = random_cyclic_case(900 + case)
    assert not g.is_ordered
    ref = ocyclic.fragment_pileup(g, r, ext)
    dg, dr = DeviceGraph(g), DeviceReads(r)
    cpu = lambda t: t.cpu().numpy()
    for mode in (1, 2):
        (fi, fv), (di, dv), touched, _ = kernels.sample_pileup(dg, dr, None, len(r), ext, mode=mode)
        assert np.array_equal(cpu(fi).view(np.uint32), ref["sample"][0]), mode
        assert np.array_equal(cpu(fv), ref["sample"][1]), mode
        assert np.array_equal(cpu(di).view(np.uint32), ref["direct"][0]), mode
        assert np.array_equal(cpu(dv), ref["direct"][1]), mode
        assert np.array_equal(np.flatnonzero(cpu(touched)) + g.min_node, ref["touched"]), mode
