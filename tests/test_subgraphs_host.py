"""CPU: the host half of the sub-graphs -- ``SubGraphList`` wraps the flat arrays of
``gpc_peak_subgraphs`` as the scipy matrices of the reference's file format, and the
Reporter writes them under the reference's names (reporter.py:11-15).  The arrays are
derived here from the oracle's sub-graphs (pinned on the live reference); the kernel that
makes them on the device is held against the same oracle in tests/test_gpu_postprocess.py."""
import numpy as np

from graph_peak_caller_b200.postprocess.graphs import SubGraphList
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph
from oracle import postprocess as opost


class _Peaks:
    def __init__(self, peaks):
        self.n_all = len(peaks)
        self.node_ptr = np.r_[0, np.cumsum([len(p[2]) for p in peaks])].astype(np.int64)
        self.nodes = np.array([v for p in peaks for v in p[2]], dtype=np.int32)


def _flat(subs):
    """Oracle sub-graphs of the line-graph components -> the kernel's flat arrays."""
    comp_ptr, node, weight, is_exit, row_ptr, col = [0], [], [], [], [0], []
    for ids, m in subs:
        k = len(ids)
        if m.shape == (1, 1):
            break
        m = m.tocsr()
        m.sort_indices()
        for i in range(k):
            cols = m.indices[m.indptr[i]:m.indptr[i + 1]]
            data = m.data[m.indptr[i]:m.indptr[i + 1]]
            assert cols.size and np.all(data == data[0])
            node.append(int(ids[i]))
            weight.append(-int(data[0]))
            is_exit.append(int(cols[-1] == k))
            col.extend(int(c) for c in cols if c < k)
            row_ptr.append(len(col))
        comp_ptr.append(len(node))
    a = lambda x, t=np.int32: np.array(x, dtype=t)
    return a(comp_ptr), a(node), a(weight), a(is_exit, np.uint8), a(row_ptr), a(col)


def test_sub_graph_list_and_reporter_files(tmp_path):
    rng = np.random.default_rng(5)
    g = make_graph(2500, site_rate=0.1, seed=9, min_node=1, mnp_frac=0.2, indel_mean=4.0)
    G = g.size_bp
    idx = np.unique(np.r_[0, rng.integers(1, G, size=G // 9)])
    val = (np.arange(idx.size) % 2).astype(bool)
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=G // 5)])
    sval = rng.integers(0, 6, size=sidx.size).astype(np.float64)
    peaks, subs = opost.max_paths(g, idx, val, sidx, sval, G, with_subgraphs=True)
    n_comp = sum(1 for _, m in subs if m.shape != (1, 1))
    assert 0 < n_comp < len(subs)
    calls = []

    def fetch():
        calls.append(1)
        return _flat(subs)

    lst = SubGraphList(fetch, _Peaks(peaks))
    assert len(lst) == len(subs) and not calls          # nothing fetched until looked at
    for k, (ids, m) in enumerate(subs):
        assert list(lst[k]._node_ids) == list(ids)
        assert lst[k]._graph.shape == m.shape
        assert np.array_equal(lst[k]._graph.toarray(), m.toarray())
    assert len(calls) == 1
    pick = np.array([len(subs) - 1, 2, 0])
    sub = lst.subset(pick)
    assert len(sub) == 3 and list(sub[1]._node_ids) == list(subs[2][0]) and len(calls) == 1
    # the Reporter's files (reporter.py:11-15)
    base = str(tmp_path / "chr1_")
    Reporter(base).add("sub_graphs", sub)
    graphs = np.load(base + "sub_graphs.graphs.npz", allow_pickle=True)
    ids = np.load(base + "sub_graphs.nodeids.npz", allow_pickle=True)
    assert sorted(graphs.files) == ["peak0", "peak1", "peak2"]
    for i, k in enumerate(pick):
        assert np.array_equal(graphs["peak%d" % i][()].toarray(), subs[k][1].toarray())
        assert list(ids["peak%d" % i]) == list(subs[k][0])
