"""CPU: peak summits.  The per-peak code of the summit kernel
(csrc/summits.cuh, compiled for the host) against the oracle's restatement of
PeakCollection.cut_around_summit, the oracle against the live reference where it
exists, and the reference's own known answer."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.synthetic import make_graph
from oracle import refbridge, summits as osummits

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def summits_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostsim") / "summits_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "hostsim", "summits_host.cpp")])
    lib = ctypes.CDLL(so)

    def run(graph, peaks, q_pos, q_val, window):
        rec = np.zeros((graph.n_nodes + 1, 4), dtype=np.uint32)
        rec[:, 0] = graph.node_off
        rec[:-1, 3] = graph.node_len
        start = np.array([p[0] for p in peaks], dtype=np.int32)
        end = np.array([p[1] for p in peaks], dtype=np.int32)
        ptr = np.r_[0, np.cumsum([len(p[2]) for p in peaks])].astype(np.uint32)
        nodes = np.array([n for p in peaks for n in p[2]], dtype=np.int32)
        q_pos = np.ascontiguousarray(q_pos, dtype=np.uint32)
        q_val = np.ascontiguousarray(q_val, dtype=np.float64)
        out = np.zeros((len(peaks), 6), dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.peak_summits_host(vp(rec), ctypes.c_int32(graph.min_node), vp(start), vp(end), vp(ptr),
                              vp(nodes), ctypes.c_long(len(peaks)), vp(q_pos), vp(q_val),
                              ctypes.c_long(q_pos.size), ctypes.c_int32(window), vp(out))
        return out
    return run


def random_peaks(graph, n, rng, max_nodes=6):
    peaks = []
    while len(peaks) < n:
        v = int(rng.integers(0, graph.n_nodes))
        nodes = [v]
        for _ in range(int(rng.integers(0, max_nodes))):
            succ = graph.succ[graph.succ_ptr[v]:graph.succ_ptr[v + 1]]
            if succ.size == 0:
                break
            v = int(rng.choice(succ))
            nodes.append(v)
        first, last = int(graph.node_len[nodes[0]]), int(graph.node_len[nodes[-1]])
        if first == 0 or last == 0:
            continue
        start = int(rng.integers(0, first))
        end = int(rng.integers(1, last + 1))
        if len(nodes) == 1:
            if last < 2:
                continue
            start = int(rng.integers(0, last - 1))
            end = int(rng.integers(start + 1, last + 1))
        peaks.append((start, end, [x + graph.min_node for x in nodes]))
    return peaks


def random_track(size, rng, n_runs, levels):
    pos = np.r_[0, np.sort(rng.choice(np.arange(1, size), size=n_runs - 1, replace=False))]
    val = rng.choice(levels, size=n_runs)
    keep = np.r_[True, val[1:] != val[:-1]]
    return pos[keep], val[keep]


def unpack(row, peak):
    summit, length, first, last, a, b = (int(x) for x in row)
    return summit, length, (a, b, peak[2][first:last + 1])


@pytest.mark.parametrize("trial", range(6))
def test_kernel_code_equals_oracle(summits_host, trial):
    rng = np.random.default_rng(500 + trial)
    g = make_graph(int(rng.integers(3000, 20000)), float(rng.choice([0.02, 0.1])), seed=trial,
                   mnp_frac=0.2, indel_mean=float(rng.choice([3, 15])))
    if trial == 5:
        g = Graph.from_edges(g.node_len, np.repeat(np.arange(g.n_nodes), np.diff(g.succ_ptr)),
                             g.succ, min_node=1000)
    peaks = random_peaks(g, 400, rng)
    # few levels and long runs: plateaus of the maximum, ties between pieces
    levels = np.array([0.0, 0.5, 1.25, 3.0, 3.0000000000000004, 7.5])[:int(rng.integers(2, 7))]
    q_pos, q_val = random_track(g.size_bp, rng, int(rng.integers(5, g.size_bp // 20)), levels)
    if trial == 3:                      # a track that starts after position 0: zeros before it
        q_pos, q_val = q_pos[1:], q_val[1:]
    dense = osummits.dense_track(q_pos, q_val, g.size_bp)
    for window in (1, 7, 60):
        got = summits_host(g, peaks, q_pos, q_val, window)
        for row, peak in zip(got, peaks):
            want = osummits.cut_around_summit(g.node_off, g.node_len, g.min_node, dense,
                                              peak[0], peak[1], peak[2], window)
            assert unpack(row, peak) == want, (peak, window)


def test_reference_known_answer(summits_host):
    # reference tests/test_command_line_interface.py:121-137: q = 3 everywhere,
    # Peak(0, 2, [1, 2]) on the graph of tests/vg_test_graph.json, window 2 -> Peak(2, 6, [1])
    g = Graph({1: 7, 2: 4, 3: 7, 4: 4}, {1: [2, 3], 2: [4], 3: [4]})
    peak = (0, 2, [1, 2])
    got = unpack(summits_host(g, [peak], [0], [3.0], 2)[0], peak)
    assert got == (4, 9, (2, 6, [1]))
    dense = osummits.dense_track([0], [3.0], g.size_bp)
    assert osummits.cut_around_summit(g.node_off, g.node_len, 1, dense, 0, 2, [1, 2], 2) == got
    # one sharp maximum on the second node: the cut is centred on it and spans both nodes
    got = unpack(summits_host(g, [peak], [0, 8, 9], [1.0, 5.0, 1.0], 3)[0], peak)
    assert got == (8, 9, (5, 2, [1, 2]))


@pytest.mark.skipif(not refbridge.available(), reason="/root/reference not present")
@pytest.mark.parametrize("trial", range(3))
def test_oracle_equals_live_reference(trial):
    refbridge._ref()
    from graph_peak_caller.mindense import DensePileup
    from graph_peak_caller.peakcollection import Peak, PeakCollection
    rng = np.random.default_rng(900 + trial)
    g = make_graph(int(rng.integers(3000, 9000)), 0.05, seed=40 + trial, mnp_frac=0.2)
    rg = refbridge.to_ref_graph(g)
    peaks = random_peaks(g, 150, rng)
    q_pos, q_val = random_track(g.size_bp, rng, 200, np.array([0.0, 1.0, 2.5, 4.0]))
    dense = osummits.dense_track(q_pos, q_val, g.size_bp)
    for window in (5, 60):
        collection = PeakCollection([Peak(s, e, list(n), graph=rg) for s, e, n in peaks])
        collection.cut_around_summit(DensePileup(rg, dense), n_base_pairs_around=window)
        for cut, peak in zip(collection.intervals, peaks):
            _, _, want = osummits.cut_around_summit(g.node_off, g.node_len, g.min_node, dense,
                                                    peak[0], peak[1], peak[2], window)
            assert (cut.start_position.offset, cut.end_position.offset,
                    list(cut.region_paths)) == want


# ---- the Python side of the device path, with the kernel's host build in the kernel's place ----

class _Host:
    """Stands in for a CUDA tensor holding the kernel's output."""

    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


@pytest.fixture
def host_kernel(summits_host, monkeypatch):
    """PeakCollection.summits / cut_around_summit / analysis_interface.get_summits run as they
    do on the GPU box (tests/test_gpu_zz_interface.py), except that uploads are no-ops and
    gpc_peak_summits is replaced by the same per-peak code compiled for the host."""
    import types
    from graph_peak_caller_b200 import device, kernels
    from graph_peak_caller_b200.sparsediffs import SparseValues

    def peak_summits(graph, start, end, ptr, nodes, q_pos, q_val, window):
        peaks = [(int(start[i]), int(end[i]), nodes[ptr[i]:ptr[i + 1]].tolist())
                 for i in range(start.size)]
        return _Host(summits_host(graph, peaks, q_pos, q_val, window))

    monkeypatch.setattr(kernels, "torch_cuda", lambda: types.SimpleNamespace(float64=np.float64))
    monkeypatch.setattr(kernels, "peak_summits", peak_summits)
    monkeypatch.setattr(device, "device_graph", lambda g: g)
    monkeypatch.setattr(device, "to_device", lambda a, dtype=None: np.asarray(a))
    monkeypatch.setattr(SparseValues, "device_arrays",
                        lambda self, value_dtype=None: (np.asarray(self.indices, dtype=np.uint32),
                                                        np.asarray(self.values, dtype=np.float64)))


def test_cut_around_summit_python_side(host_kernel):
    from graph_peak_caller_b200.peakcollection import Peak, PeakCollection
    from graph_peak_caller_b200.sparsediffs import SparseValues
    rng = np.random.default_rng(77)
    g = make_graph(9000, 0.05, seed=8, mnp_frac=0.2)
    peaks = random_peaks(g, 120, rng)
    q_pos, q_val = random_track(g.size_bp, rng, 300, np.array([0.0, 1.0, 2.5, 4.0]))
    track = SparseValues(q_pos, q_val)
    track.track_size = g.size_bp
    dense = osummits.dense_track(q_pos, q_val, g.size_bp)
    collection = PeakCollection([Peak(s, e, list(n), graph=g, score=1.5, unique_id="peak%d" % i)
                                 for i, (s, e, n) in enumerate(peaks)])
    collection.cut_around_summit(track, 25)
    for cut, peak in zip(collection.intervals, peaks):
        _, _, want = osummits.cut_around_summit(g.node_off, g.node_len, g.min_node, dense,
                                                peak[0], peak[1], peak[2], 25)
        assert (cut.start_position.offset, cut.end_position.offset, list(cut.region_paths)) == want
        assert cut.score == 1.5 and cut.unique_id.startswith("peak") and cut.length() <= 50
    assert PeakCollection([]).summits(track).shape == (0, 6)


def test_get_summits_known_answer_python_side(host_kernel, tmp_path):
    # reference tests/test_command_line_interface.py:121-138 (the device run of the same
    # steps: tests/test_gpu_zz_interface.py::test_get_summits_known_answer)
    import types
    from graph_peak_caller_b200 import analysis_interface as ai, callpeaks_interface as ci
    from graph_peak_caller_b200.peakcollection import Peak, PeakCollection
    from graph_peak_caller_b200.peakfasta import PeakFasta
    from graph_peak_caller_b200.sequencegraph import SequenceGraph
    from graph_peak_caller_b200.sparsediffs import SparseValues
    from test_sequences import VG_TEST_GRAPH, write
    d = str(tmp_path) + "/"
    ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write(tmp_path / "vg_test_graph.json", VG_TEST_GRAPH),
        out_file_name=d + "testgraph.obg"))
    qvalues = SparseValues(np.array([0]), np.array([3.0]))
    qvalues.track_size = 22
    qvalues.to_sparse_files(d + "test_qvalues")
    PeakFasta(SequenceGraph.from_file(d + "testgraph.obg.sequences")).write_max_path_sequences(
        d + "test_max_paths.fasta", PeakCollection([Peak(0, 2, [1, 2], score=3)]))
    ai.get_summits(types.SimpleNamespace(graph=d + "testgraph.obg", window_size="2",
                                         peaks_fasta_file=d + "test_max_paths.fasta",
                                         q_values_base_name=d + "test_qvalues"))
    result = PeakCollection.from_fasta_file(d + "test_max_paths_summits.fasta")
    assert result.intervals[0] == Peak(2, 6, [1])
    assert result.intervals[0].sequence.lower() == "tccc"
