"""GPU parity: LinearMap.from_graph as kernels (gpc_linear_map_from_graph, csrc/linearmap.cu)
against the oracle restatement of control/linearmap.py:102-154 (itself bit-equal to the
live reference, tests/test_oracle_vs_reference.py) and against the library's host sweep.
Bit-exact: the values are sums of integer node sizes."""
import numpy as np
import pytest

from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.synthetic import make_graph
from oracle import control as ocontrol

pytestmark = pytest.mark.gpu


def device_map(g):
    """gpc_linear_map_from_graph through the C ABI: (starts, ends, barriers)."""
    import ctypes
    import torch
    from graph_peak_caller_b200 import _lib
    from graph_peak_caller_b200.device import device_graph, empty, ptr, stream_ptr
    dg = device_graph(g)
    starts, ends = empty(g.n_nodes, torch.float64), empty(g.n_nodes, torch.float64)
    barriers = ctypes.c_uint64(0)
    _lib.check(_lib.load().gpc_linear_map_from_graph(
        ctypes.byref(dg.c), ptr(starts), ptr(ends), ctypes.byref(barriers), stream_ptr()))
    return starts.cpu().numpy(), ends.cpu().numpy(), int(barriers.value)


def check(g, min_barriers=None, valid=True):
    """``valid``: the graph has one source and one sink, so that the map is a LinearMap
    (starts < ends everywhere; the reference's constructor asserts it, linearmap.py:24-25).
    Graphs with sources / sinks in the middle only go through the arrays."""
    got_s, got_e, barriers = device_map(g)
    want_s, want_e = ocontrol.linear_map_from_graph(g)
    assert got_s.dtype == np.float64 and got_e.dtype == np.float64
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_e, want_e)
    assert 1 <= barriers <= g.n_nodes
    if min_barriers is not None:
        assert barriers >= min_barriers
    if not valid:
        return None
    lm = LinearMap.from_graph(g)
    assert lm.n_barriers == barriers, "from_graph did not run on the device"
    assert np.array_equal(lm.as_arrays()[0], want_s) and np.array_equal(lm.as_arrays()[1], want_e)
    assert LinearMap.from_graph_host(g) == lm
    return lm


def random_dag(rng, n, p_chain=0.85, max_jump=6, p_jump=0.4, max_len=9):
    node_len = rng.integers(1, max_len, n)
    edges = set()
    for u in range(n - 1):
        if rng.random() < p_chain:
            edges.add((u, u + 1))
        for _ in range(int(rng.integers(0, 3))):
            v = u + int(rng.integers(1, max_jump))
            if v < n and rng.random() < p_jump:
                edges.add((u, v))
    e = np.array(sorted(edges), dtype=np.int64).reshape(-1, 2)
    return Graph.from_edges(node_len, e[:, 0], e[:, 1], min_node=int(rng.integers(1, 50)))


def test_single_node_and_chain():
    check(Graph.from_edges([7], [], []))
    lm = check(Graph.from_edges([3, 4, 5, 6], [0, 1, 2], [1, 2, 3]), min_barriers=4)
    assert list(lm.as_arrays()[0]) == [0, 3, 7, 12] and list(lm.as_arrays()[1]) == [3, 7, 12, 18]


def test_bubbles_and_a_deletion_over_several_nodes():
    # 0 -> {1 | 2} -> 3 -> 4 -> 5 -> 6 and the deletion 3 -> 6 (jumps over nodes 4 and 5)
    g = Graph.from_edges([5, 1, 3, 4, 2, 6, 3], [0, 0, 1, 2, 3, 4, 5, 3], [1, 2, 3, 3, 4, 5, 6, 6])
    lm = check(g)
    assert lm.n_barriers == 3          # nodes 0, 3 and 6
    assert list(lm.as_arrays()[0]) == [0, 5, 5, 8, 12, 14, 20]


def test_sources_and_sinks_inside_segments():
    # missing chain edges give nodes without predecessors (max_dists stays 0) and nodes
    # without successors in the middle of the graph
    rng = np.random.default_rng(5)
    for _ in range(60):
        check(random_dag(rng, int(rng.integers(2, 60)), p_chain=0.6), valid=False)


def test_random_graphs_small_and_wide_jumps():
    rng = np.random.default_rng(6)
    for _ in range(40):
        check(random_dag(rng, int(rng.integers(1, 200))), valid=False)
    for _ in range(10):        # jumps over hundreds of nodes: long segments, few barriers
        check(random_dag(rng, int(rng.integers(500, 3000)), max_jump=400, p_jump=0.05), valid=False)


def test_more_segments_than_one_scan_block():
    # > 1024 segments per direction: block composites and their serial pass take part
    rng = np.random.default_rng(7)
    g = random_dag(rng, 20000, p_chain=0.97, max_jump=4, p_jump=0.3)
    check(g, min_barriers=2048, valid=False)


def test_variation_graph_of_the_generator():
    g = make_graph(400000, 0.03, seed=11, mnp_frac=0.1)
    lm = check(g, min_barriers=g.n_nodes // 8)
    # every read start projects to an integer on such a graph (device.py: integer_linear)
    s, e = lm.as_arrays()
    assert np.all(s == np.floor(s)) and np.all(e > s)


def test_unordered_graph_is_refused():
    g = Graph.from_edges([2, 2, 2], [0, 1, 2], [1, 2, 0], allow_unordered=True)
    with pytest.raises(ValueError):
        LinearMap.from_graph(g)
