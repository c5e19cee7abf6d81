"""GPU parity: hole cleaning, max paths and peak scoring through the C ABI
against the oracle (bit-exact: integer positions, node lists, info flags; the
score is a max over q-values that are given as input, so exact as well)."""
import numpy as np
import pytest

from graph_peak_caller_b200.synthetic import make_graph
from oracle import postprocess as opost

pytestmark = pytest.mark.gpu


def cpu(t):
    return t.cpu().numpy()


def gpu_clean(graph, idx, val, max_size, touched_ids):
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, to_device
    dg = DeviceGraph(graph)
    touched = None
    if touched_ids is not None:
        t = np.zeros(graph.n_nodes, dtype=np.uint8)
        t[np.asarray(sorted(touched_ids), dtype=np.int64) - graph.min_node] = 1
        touched = to_device(t)
    op, ov = kernels.clean_holes(dg, to_device(np.asarray(idx, dtype=np.int32)),
                                 to_device(np.asarray(val, dtype=np.uint8)), max_size, touched)
    return cpu(op).astype(np.int64), cpu(ov).astype(bool)


def gpu_max_paths(graph, idx, val, sidx, sval, qidx, qval, shift):
    from graph_peak_caller_b200 import kernels
    from graph_peak_caller_b200.device import DeviceGraph, to_device
    dg = DeviceGraph(graph)
    out = kernels.max_paths(dg, to_device(np.asarray(idx, dtype=np.int32)),
                            to_device(np.asarray(val, dtype=np.uint8)),
                            to_device(np.asarray(sidx, dtype=np.int32)),
                            to_device(np.asarray(sval, dtype=np.float64)),
                            to_device(np.asarray(qidx, dtype=np.int32)),
                            to_device(np.asarray(qval, dtype=np.float64)), shift)
    o = {k: (cpu(v) if k not in ("n", "arena") else v) for k, v in out.items()}
    peaks = []
    for i in range(o["n"]):
        nodes = o["nodes"][o["node_ptr"][i]:o["node_ptr"][i + 1]].tolist()
        info = (bool(o["info"][i] & 1), bool(o["info"][i] & 2))
        peaks.append((int(o["start"][i]), int(o["end"][i]), nodes, info,
                      float(o["score"][i]), int(o["length"][i])))
    return peaks, o["order"].astype(np.int64)


def assert_same_peaks(got, ref):
    """Same peaks (interval, nodes, flags: exact); score within 1e-6 relative --
    the reference scores a peak by the max over a dense track rebuilt with a
    float cumsum (sparsediffs.py:64-76), so its scores carry ulp noise and the
    order of peaks whose scores tie up to that noise is not defined."""
    assert len(got) == len(ref)
    key = lambda p: (p[0], p[1], tuple(p[2]))
    g = sorted(got, key=key)
    r = sorted(ref, key=key)
    for a, b in zip(g, r):
        assert key(a) == key(b) and tuple(a[3]) == tuple(b[3])
        assert abs(a[4] - b[4]) <= 1e-6 * abs(b[4]) + 1e-12
    scores = [p[4] for p in got]
    assert all(x >= y for x, y in zip(scores, scores[1:])), "sorted by score, descending"
    # where scores are clearly different the order must agree with the reference
    strict = lambda ps: [key(p) for i, p in enumerate(ps)
                         if all(abs(p[4] - o[4]) > 1e-9 for j, o in enumerate(ps) if j != i)]
    sg, sr = strict(got), strict(ref)
    assert [k for k in sg if k in set(sr)] == [k for k in sr if k in set(sg)]


def random_track(rng, G):
    k = int(rng.integers(1, max(2, G // 8)))
    idx = np.unique(np.r_[0, rng.integers(1, G, size=k)])
    first = int(rng.integers(0, 2))
    val = ((np.arange(idx.size) + first) % 2).astype(bool)
    return idx, val


@pytest.mark.parametrize("trial", range(60))
def test_hole_cleaner_random_tracks(trial):
    rng = np.random.default_rng(900 + trial)
    g = make_graph(int(rng.integers(200, 4000)),
                   site_rate=float(rng.choice([0.03, 0.1, 0.2])), seed=trial,
                   min_node=int(rng.choice([1, 1, 50])), mnp_frac=0.2,
                   indel_mean=float(rng.choice([2, 6])))
    idx, val = random_track(rng, g.size_bp)
    max_size = int(rng.integers(1, 30))
    touched = None
    if rng.random() < 0.7:
        t = np.flatnonzero(rng.random(g.n_nodes) < 0.8) + g.min_node
        touched = set(t.tolist()) or None
    try:
        ri, rv = opost.clean_holes(g, idx, val, max_size, touched)
    except Exception:
        pytest.skip("reference algorithm raises on this input")
    gi, gv = gpu_clean(g, idx, val, max_size, touched)
    assert np.array_equal(gi, ri), (gi, ri)
    assert np.array_equal(gv, rv)


@pytest.mark.parametrize("trial", range(60))
def test_max_paths_random_tracks(trial):
    rng = np.random.default_rng(300 + trial)
    g = make_graph(int(rng.integers(200, 4000)),
                   site_rate=float(rng.choice([0.03, 0.1, 0.2])), seed=trial + 1000,
                   min_node=1, mnp_frac=0.2, indel_mean=float(rng.choice([2, 6])))
    G = g.size_bp
    idx, val = random_track(rng, G)
    sk = int(rng.integers(1, max(2, G // 5)))
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=sk)])
    sval = rng.integers(0, 6, size=sidx.size).astype(np.float64)   # many ties
    qk = int(rng.integers(1, max(2, G // 6)))
    qidx = np.unique(np.r_[0, rng.integers(1, G, size=qk)])
    qval = np.round(rng.random(qidx.size) * 5, 1)
    F = int(rng.integers(1, 60))
    ref_all = opost.max_paths(g, idx, val, sidx, sval, G)
    ref = opost.score_and_filter_peaks(g, ref_all, qidx, qval, F)
    got_all, order = gpu_max_paths(g, idx, val, sidx, sval, qidx, qval, g.min_node - 1)
    assert [(p[0], p[1], p[2], p[3]) for p in got_all] == \
        [(p[0], p[1], p[2], tuple(p[3])) for p in ref_all]
    final = [got_all[i] for i in order if got_all[i][5] >= F]
    assert_same_peaks(final, ref)


def test_offset_ids_sane_internal_peaks():
    """min_node != 1: the reference mis-numbers single-node peaks
    (maxpaths.py:34); with the sane shift ids are real node ids."""
    g = make_graph(3000, 0.05, seed=4, min_node=101)
    G = g.size_bp
    rng = np.random.default_rng(1)
    idx, val = random_track(rng, G)
    sidx, sval = np.array([0]), np.array([1.0])
    ref_all = opost.max_paths(g, idx, val, sidx, sval, G, replicate_internal_id_quirk=False)
    got_all, _ = gpu_max_paths(g, idx, val, sidx, sval, sidx, sval, g.min_node - 1)
    assert [(p[0], p[1], p[2]) for p in got_all] == [(p[0], p[1], p[2]) for p in ref_all]
    quirk_all = opost.max_paths(g, idx, val, sidx, sval, G, replicate_internal_id_quirk=True)
    got_q, _ = gpu_max_paths(g, idx, val, sidx, sval, sidx, sval, -(g.min_node - 1))
    assert [(p[0], p[1], p[2]) for p in got_q] == [(p[0], p[1], p[2]) for p in quirk_all]


@pytest.mark.parametrize("backbone,site_rate", [(2500, 0.05), (1500, 0.2)])
def test_max_paths_one_giant_component(backbone, site_rate):
    """A track that is set everywhere: the whole graph is one component of
    several hundred vertices (the replay's global-memory path; small components
    keep their working arrays in shared memory)."""
    g = make_graph(backbone, site_rate=site_rate, seed=77, min_node=1, mnp_frac=0.2,
                   indel_mean=4.0)
    G = g.size_bp
    assert g.n_nodes > 300
    rng = np.random.default_rng(3)
    idx, val = np.array([0]), np.array([1], dtype=np.uint8)
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=G // 7)])
    sval = rng.integers(0, 5, size=sidx.size).astype(np.float64)
    ref_all = opost.max_paths(g, idx, val.astype(bool), sidx, sval, G)
    got_all, _ = gpu_max_paths(g, idx, val, sidx, sval, sidx, sval, 0)
    assert len(got_all) == 1 and len(got_all[0][2]) > 128
    assert [(p[0], p[1], p[2], p[3]) for p in got_all] == \
        [(p[0], p[1], p[2], tuple(p[3])) for p in ref_all]


def _sub_graphs_of(graph, idx, val, sidx, sval, min_node_shift=None):
    """SparseMaxPaths.run() of the product on host-made tracks -> (peaks, sub-graphs)."""
    from graph_peak_caller_b200.postprocess.maxpaths import SparseMaxPaths
    from graph_peak_caller_b200.sparsediffs import SparseValues
    track = SparseValues(np.asarray(idx), np.asarray(val, dtype=bool))
    track.track_size = graph.size_bp
    score = SparseValues(np.asarray(sidx), np.asarray(sval, dtype=np.float64))
    score.track_size = graph.size_bp
    return SparseMaxPaths(track, graph, score).run()


@pytest.mark.parametrize("trial", range(30))
def test_sub_graphs_random_tracks(trial):
    """The second result of SparseMaxPaths.run(): node ids and matrix of every peak's
    line-graph component against the oracle (pinned on the live reference's SubGraph
    objects, tests/test_oracle_vs_reference.py)."""
    rng = np.random.default_rng(900 + trial)
    g = make_graph(int(rng.integers(200, 4000)),
                   site_rate=float(rng.choice([0.03, 0.1, 0.2])), seed=trial + 2000,
                   min_node=int(rng.choice([1, 1, 57])), mnp_frac=0.2,
                   indel_mean=float(rng.choice([2, 6])))
    G = g.size_bp
    idx, val = random_track(rng, G)
    sk = int(rng.integers(1, max(2, G // 5)))
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=sk)])
    sval = rng.integers(0, 6, size=sidx.size).astype(np.float64)
    ref_peaks, ref_subs = opost.max_paths(g, idx, val, sidx, sval, G,
                                          replicate_internal_id_quirk=False, with_subgraphs=True)
    peaks, subs = _sub_graphs_of(g, idx, val, sidx, sval)
    assert len(subs) == len(peaks) == len(ref_subs)
    for k, (rn, rm) in enumerate(ref_subs):
        s = subs[k]
        assert np.array_equal(np.asarray(s._node_ids), np.asarray(rn)), k
        assert s._graph.shape == rm.shape, k
        assert np.array_equal(s._graph.toarray(), rm.toarray()), k
    # a subset follows the peaks it is taken with
    if len(subs) > 2:
        pick = [len(subs) - 1, 0, 1]
        sub = subs.subset(pick)
        assert [list(np.asarray(x._node_ids)) for x in sub] == \
            [list(np.asarray(ref_subs[i][0])) for i in pick]


def test_sub_graphs_one_giant_component_and_no_component():
    g = make_graph(1500, site_rate=0.2, seed=77, min_node=1, mnp_frac=0.2, indel_mean=4.0)
    G = g.size_bp
    rng = np.random.default_rng(3)
    sidx = np.unique(np.r_[0, rng.integers(1, G, size=G // 7)])
    sval = rng.integers(0, 5, size=sidx.size).astype(np.float64)
    idx, val = np.array([0]), np.array([True])
    _, ref_subs = opost.max_paths(g, idx, val, sidx, sval, G, with_subgraphs=True)
    _, subs = _sub_graphs_of(g, idx, val, sidx, sval)
    assert len(subs) == 1 and subs[0]._graph.shape[0] > 128
    assert np.array_equal(subs[0]._graph.toarray(), ref_subs[0][1].toarray())
    assert np.array_equal(np.asarray(subs[0]._node_ids), ref_subs[0][0])
    # one peak strictly inside one node: no line graph at all
    off = int(g.node_off[3])
    idx, val = np.array([0, off + 1, off + 2]), np.array([False, True, False])
    if g.node_len[3] < 4:
        pytest.skip("node too short for an internal peak")
    ref_peaks, ref_subs = opost.max_paths(g, idx, val, sidx, sval, G,
                                          replicate_internal_id_quirk=False, with_subgraphs=True)
    peaks, subs = _sub_graphs_of(g, idx, val, sidx, sval)
    assert len(subs) == len(ref_subs) == 1
    assert list(subs[0]._node_ids) == list(ref_subs[0][0]) and subs[0]._graph.shape == (1, 1)
