"""GPU: the whole-array steps of the fragment-length estimate --
gpc_node_read_counts, gpc_path_start_positions and the device sort of the projected
starts -- against the golden vectors made by the unmodified reference
(tests/golden/shift/shift_model.npz: its HaploTyper path and LinearFilter starts), the
oracle's loops, and the host statements of the same module (bit-exact: integers)."""
import os
import types

import numpy as np
import pytest

from graph_peak_caller_b200 import shiftestimation as se
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import shift as oshift

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "shift", "shift_model.npz"))


@pytest.fixture
def host_form():
    """Runs a callable with the device switched off (the numpy statements)."""
    def run(fn):
        se.DEVICE = False
        try:
            return fn()
        finally:
            se.DEVICE = None
    return run


def test_golden_linear_filter_on_device():
    n_bp, gseed, n_reads, rseed = GOLDEN["lf_graph_args"].tolist()
    g = make_graph(n_bp, 0.05, seed=gseed, mnp_frac=0.2)
    reads = make_reads(g, n_reads, 20, 80, seed=rseed, peak_frac=0.5, peak_spacing=2500)
    from graph_peak_caller_b200 import _lib
    before = _lib.launch_count()
    lf = se.LinearFilter.from_reads_and_graph(reads, g)
    assert (lf._path + g.min_node).tolist() == GOLDEN["lf_path"].tolist()
    found = lf.find_start_positions()
    assert _lib.launch_count() > before                     # kernels ran, not numpy
    assert found["+"].tolist() == GOLDEN["lf_plus"].tolist()
    assert found["-"].tolist() == GOLDEN["lf_minus"].tolist()
    srt = lf.find_start_positions(sort=True)
    assert srt["+"].tolist() == sorted(GOLDEN["lf_plus"].tolist())
    assert srt["-"].tolist() == sorted(GOLDEN["lf_minus"].tolist())


@pytest.mark.parametrize("n_bp,n_reads,min_node,seed", [(3000, 50, 1, 1), (60_000, 20_000, 1, 2),
                                                        (400_000, 1_500_000, 37, 3)])
def test_device_equals_host_and_oracle(n_bp, n_reads, min_node, seed, host_form):
    g = make_graph(n_bp, 0.06, seed=seed, mnp_frac=0.2, min_node=min_node)
    reads = make_reads(g, n_reads, 25, 120, seed=seed + 10, peak_frac=0.6, peak_spacing=3000)
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import to_device
    counts = kernels.node_read_counts(to_device(reads.path), g.min_node, g.n_nodes).cpu().numpy()
    fwd = reads.path[reads.path > 0].astype(np.int64) - g.min_node
    assert np.array_equal(counts, np.bincount(fwd, minlength=g.n_nodes))
    lf = se.LinearFilter.from_reads_and_graph(reads, g)
    host_lf = host_form(lambda: se.LinearFilter.from_reads_and_graph(reads, g))
    assert np.array_equal(lf._path, host_lf._path)
    for sort in (False, True):
        got = lf.find_start_positions(sort=sort)
        want = host_form(lambda: host_lf.find_start_positions(sort=sort))
        assert np.array_equal(got["+"], want["+"]) and np.array_equal(got["-"], want["-"])
        assert got["+"].size + got["-"].size > 0
    if n_reads <= 20_000:
        path = (lf._path + g.min_node).tolist()
        want = oshift.start_positions(g.node_len, g.min_node, path,
                                      zip(reads.start_node.tolist(), reads.start_off.tolist()))
        got = lf.find_start_positions()
        assert got["+"].tolist() == want["+"] and got["-"].tolist() == want["-"]


def test_reference_form_and_start_offset(host_form):
    """LinearFilter(reads, graph, path, start_offset): a path that starts inside its first
    node gives negative positions before it (sorted on the host then); starts off the path,
    on node 0 and beyond the graph are dropped."""
    g = Graph({1: 3, 2: 2, 3: 2, 4: 5, 5: 1}, {1: [2, 3], 2: [4], 3: [4], 4: [5]})
    reads = types.SimpleNamespace(start_node=np.array([1, 3, -2, -4, 2, 0, 9, -9, 1], dtype=np.int32),
                                  start_off=np.array([2, 1, 1, 5, 0, 0, 1, 1, 0], dtype=np.int32))
    for start_offset in (0, 2):
        lf = se.LinearFilter(reads, g, [0, 1, 3, 4], start_offset)
        for sort in (False, True):
            got = lf.find_start_positions(sort=sort)
            want = host_form(lambda: lf.find_start_positions(sort=sort))
            assert got["+"].tolist() == want["+"].tolist() and got["-"].tolist() == want["-"].tolist()
    found = se.LinearFilter(reads, g, [0, 1, 3, 4]).find_start_positions()
    assert found["+"].tolist() == [2, 3, 0] and found["-"].tolist() == [3 + 2 - 1, 5 + 5 - 5]


def test_estimate_is_the_same_number(host_form):
    """The whole estimate, device steps against host steps: same d, same model lines."""
    F, R = 180, 30
    g = make_graph(800000, 0.02, seed=70, mnp_frac=0.1)
    reads = make_reads(g, 50000, R, F, seed=71, peak_frac=0.8, peak_spacing=5000, dup_frac=0.0)

    def estimate():
        lf = se.LinearFilter.from_reads_and_graph(reads, g)
        found = lf.find_start_positions(sort=True)
        est = se.MultiGraphShiftEstimator({"+": {"0": found["+"]}, "-": {"0": found["-"]}},
                                          lf.length())
        return est.get_estimates(), est.peakmodel
    d_dev, m_dev = estimate()
    d_host, m_host = host_form(estimate)
    assert d_dev == d_host and abs(d_dev - F) <= 12
    assert np.array_equal(m_dev.plus_line, m_host.plus_line)
    assert np.array_equal(m_dev.minus_line, m_host.minus_line)
