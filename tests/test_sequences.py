"""CPU: node sequences and FASTA output (second half):
``SequenceGraph`` (what the reference takes from offsetbasedgraph, absent here),
``gpc_parse_vg_sequences_host``, ``PeakFasta`` (reference peakfasta.py:4-31),
``PeakCollection.to_fasta_file / from_fasta_file`` (peakcollection.py:254-281) and
the ``sequences.fasta`` file names of the entry points.  Known answers are the
reference's own: tests/test_command_line_interface.py:18-27,102-119,121-138 on
tests/vg_test_graph.json."""
import json
import types

import numpy as np
import pytest

from graph_peak_caller_b200 import callpeaks_interface as ci
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
from graph_peak_caller_b200.peakcollection import Peak, PeakCollection
from graph_peak_caller_b200.peakfasta import PeakFasta
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.sequencegraph import SequenceGraph
from graph_peak_caller_b200.synthetic import make_graph

VG_TEST_GRAPH = ('{"node": [     {"id": 1, "sequence": "TTTCCCC"},     {"id": 2, "sequence": "TTTT"},'
                 '     {"id": 3, "sequence": "CCCCTTT"},     {"id": 4, "sequence": "ACTG"} ],'
                 '"edge": [     {"from": 1, "to": 2},     {"from": 1, "to": 3},     {"from": 2, "to": 4},'
                 '    {"from": 3, "to": 4} ] }')


def reference_graph():
    return Graph({1: 7, 2: 4, 3: 7, 4: 4}, {1: [2, 3], 2: [4], 3: [4]})


def write(path, text):
    with open(path, "w") as fh:
        fh.write(text)
    return str(path)


def complement(s):
    return s[::-1].translate(str.maketrans("acgtn", "tgcan"))


def random_sequence_json(g, rng, chunks=3, min_node=None, shuffle=True):
    """vg graph JSON lines with random bases; returns (lines, {node id: sequence})."""
    min_node = g.min_node if min_node is None else min_node
    seqs = {int(i + min_node): "".join(rng.choice(list("ACGTN"), size=int(n)))
            for i, n in enumerate(g.node_len)}
    ids = list(seqs)
    if shuffle:
        rng.shuffle(ids)
    lines = []
    for part in np.array_split(np.array(ids), chunks):
        lines.append(json.dumps({"edge": [], "node": [{"sequence": seqs[int(i)], "name": "x\\\"y",
                                                        "id": (str(int(i)) if i % 2 else int(i))}
                                                       for i in part]}))
    return lines, seqs


def test_reference_known_answer(tmp_path):
    # test_command_line_interface.py:102-119: Peak(0, 2, [1, 2]) reads "tttcccctt"
    g = reference_graph()
    sg = SequenceGraph.create_empty_from_ob_graph(g)
    assert sg.get_interval_sequence(Peak(0, 2, [1, 2])) == "n" * 9
    sg.set_sequences_using_vg_json_graph(write(tmp_path / "g.json", VG_TEST_GRAPH))
    peak = Peak(0, 2, [1, 2], score=3)
    assert sg.get_interval_sequence(peak) == "tttcccctt"
    assert sg.get_interval_sequence(Peak(2, 6, [1])) == "tccc"        # :137-138, the summit cut
    assert sg.get_interval_sequence(Peak(3, 4, [1, 3, 4])) == "cccc" + "ccccttt" + "actg"
    assert sg.get_node_sequence(4) == "actg" and sg.get_sequence(3, 4) == "ttt"
    # reverse strand: reverse complement of the node, offsets counted on that strand
    assert sg.get_sequence(-4) == "cagt" and sg.get_sequence(-1, 0, 4) == "gggg"
    assert sg.get_interval_sequence(Peak(1, 3, [-4, -2])) == "agt" + "aaa"
    # PeakFasta + from_fasta_file
    name = str(tmp_path / "peaks.fasta")
    PeakFasta(sg).write_max_path_sequences(name, PeakCollection([peak]))
    with open(name) as fh:
        lines = fh.read().split("\n")
    assert lines[0] == ">peak0 " + peak.to_file_line() and lines[1:] == ["tttcccctt", ""]
    back = PeakCollection.from_fasta_file(name)
    assert len(back.intervals) == 1 and back.intervals[0] == peak
    assert back.intervals[0].sequence == "tttcccctt" and back.intervals[0].unique_id == "peak0"
    assert back.intervals[0].score == 3


def test_files_round_trip(tmp_path):
    g = reference_graph()
    graph_name = str(tmp_path / "testgraph.obg")
    out = ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write(tmp_path / "vg_test_graph.json", VG_TEST_GRAPH),
        out_file_name=graph_name))
    assert out == graph_name
    sg = SequenceGraph.from_file(graph_name + ".sequences")        # finds the graph by name
    assert sg.get_interval_sequence(Peak(0, 2, [1, 2])) == "tttcccctt"
    sg2 = SequenceGraph.from_file(graph_name + ".sequences", graph=g)
    assert sg2.get_node_sequence(3) == "ccccttt"
    with pytest.raises(FileNotFoundError):
        SequenceGraph.from_file(str(tmp_path / "nothing.nobg.sequences"))
    with pytest.raises(ValueError):
        SequenceGraph.from_file(graph_name + ".sequences", graph=Graph({1: 5}, {}))
    # PeakCollection.to_fasta_file == PeakFasta.save_intervals; several peaks, numbered in order
    peaks = PeakCollection([Peak(0, 2, [1, 2], score=3), Peak(1, 4, [3, 4], score=1.5), Peak(0, 7, [1])])
    name = str(tmp_path / "many.fasta")
    peaks.to_fasta_file(name, sg)
    back = PeakCollection.from_fasta_file(name, graph=g)
    assert [p.unique_id for p in back.intervals] == ["peak0", "peak1", "peak2"]
    assert [p.sequence for p in back.intervals] == ["tttcccctt", "cccttt" + "actg", "tttcccc"]
    assert back.intervals == peaks.intervals and back.intervals[1].graph is g


@pytest.mark.parametrize("min_node,chunks,shuffle", [(1, 1, True), (1, 4, True), (500, 3, True),
                                                     (1, 1, False), (500, 4, False)])
def test_random_sequences_against_json_module(tmp_path, min_node, chunks, shuffle):
    # nodes in id order (what vg writes: one slice) and in any order (scattered node by node)
    rng = np.random.default_rng(7 + chunks)
    g0 = make_graph(6000, 0.06, seed=5, mnp_frac=0.2)
    src = np.repeat(np.arange(g0.n_nodes), np.diff(g0.succ_ptr))
    g = Graph.from_edges(g0.node_len, src, g0.succ, min_node)
    lines, seqs = random_sequence_json(g, rng, chunks, min_node, shuffle)
    name = write(tmp_path / "g.json", "\n".join(lines) + "\n")
    sg = SequenceGraph.from_vg_json_file(name, g)
    for node in rng.choice(list(seqs), size=200):
        node = int(node)
        assert sg.get_node_sequence(node) == seqs[node].lower()
        assert sg.get_sequence(-node) == complement(seqs[node].lower())
    # a path through the graph: first successor until the sink
    node, path = min_node, []
    while len(path) < 40:
        path.append(node)
        nxt = g.adj_list[node]
        if len(nxt) == 0:
            break
        node = int(nxt[-1])
    whole = "".join(seqs[n].lower() for n in path)
    end = int(g.node_len[path[-1] - min_node])
    assert sg.get_interval_sequence(Peak(0, end, path)) == whole
    start = int(g.node_len[path[0] - min_node]) - 1
    assert sg.get_interval_sequence(Peak(start, 1, path)) == \
        whole[start:len(whole) - end + 1]
    # file round trip keeps every base
    sg.to_file(str(tmp_path / "g.nobg.sequences"))
    again = SequenceGraph.from_file(str(tmp_path / "g.nobg.sequences"), graph=g)
    assert np.array_equal(again._sequences, sg._sequences)


def test_parser_errors(tmp_path):
    g = reference_graph()
    sg = SequenceGraph.create_empty_from_ob_graph(g)
    bad_len = VG_TEST_GRAPH.replace('"TTTT"', '"TTT"')
    with pytest.raises(ValueError, match="lengths differ"):
        sg.set_sequences_using_vg_json_graph(write(tmp_path / "a.json", bad_len))
    outside = VG_TEST_GRAPH.replace('"id": 4', '"id": 9')
    with pytest.raises(ValueError, match="outside the graph"):
        sg.set_sequences_using_vg_json_graph(write(tmp_path / "b.json", outside))
    escaped = VG_TEST_GRAPH.replace('"ACTG"', '"AC\\u0054G"')
    with pytest.raises(Exception, match="escaped sequence"):
        sg.set_sequences_using_vg_json_graph(write(tmp_path / "c.json", escaped))
    with pytest.raises(Exception, match="malformed"):
        sg.set_sequences_using_vg_json_graph(write(tmp_path / "d.json", VG_TEST_GRAPH[:-30]))
    # no nodes at all: nothing set, nothing raised
    sg.set_sequences_using_vg_json_graph(write(tmp_path / "e.json", '{"edge": []}\n\n'))
    assert sg.get_node_sequence(1) == "n" * 7
    with pytest.raises(ValueError):
        sg.set_sequence(2, "ACG")
    sg.set_sequence(2, "ACGT")
    assert sg.get_node_sequence(2) == "acgt" and sg.get_node_sequence(1) == "n" * 7


class _Caller:
    def __init__(self, peaks):
        self.max_path_peaks = peaks


def test_multigraph_sequence_files(tmp_path):
    # <base><name>_sequences.fasta per chromosome (multiplegraphscallpeaks.py:157-166);
    # a missing .sequences file is a warning
    g = reference_graph()
    ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write(tmp_path / "1.json", VG_TEST_GRAPH),
        out_file_name=str(tmp_path / "1.nobg")))
    base = str(tmp_path / "run_")
    peaks = PeakCollection([Peak(0, 2, [1, 2], score=3)])
    retrievers = ci._SequenceGraphFiles([str(tmp_path / "1.nobg.sequences"),
                                         str(tmp_path / "2.nobg.sequences")])
    m = MultipleGraphsCallpeaks(["1", "2"], [g, g], None, None, None, None, Reporter(base),
                                sequence_retrievers=retrievers)
    m._write_sequences(0, _Caller(peaks))
    m._write_sequences(1, _Caller(peaks))                      # no such file: skipped
    back = PeakCollection.from_fasta_file(base + "1_sequences.fasta")
    assert back.intervals[0].sequence == "tttcccctt"
    assert not (tmp_path / "run_2_sequences.fasta").exists()
    # the reference's form: an iterator consumed in chromosome order; and a plain list
    sg = SequenceGraph.from_file(str(tmp_path / "1.nobg.sequences"))
    for source, name in ((iter([sg]), "it_"), ([sg], "list_")):
        m = MultipleGraphsCallpeaks(["1"], [g], None, None, None, None,
                                    Reporter(str(tmp_path / name)), sequence_retrievers=source)
        m._write_sequences(0, _Caller(peaks))
        assert PeakCollection.from_fasta_file(
            str(tmp_path / name) + "1_sequences.fasta").intervals[0].sequence == "tttcccctt"
    # nothing asked for, or a Reporter that keeps everything in memory: no file
    m = MultipleGraphsCallpeaks(["1"], [g], None, None, None, None, Reporter(None),
                                sequence_retrievers=[sg])
    m._write_sequences(0, _Caller(peaks))


# ---- the args-driven callers (analysis/analysis_interface.py) ---------------------------

def test_peaks_to_fasta_known_answer(tmp_path):
    # reference tests/test_command_line_interface.py:102-119
    from graph_peak_caller_b200 import analysis_interface as ai
    graph_name = str(tmp_path / "testgraph.obg")
    ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write(tmp_path / "vg_test_graph.json", VG_TEST_GRAPH),
        out_file_name=graph_name))
    PeakCollection([Peak(0, 2, [1, 2], score=3)]).to_file(
        str(tmp_path / "testintervals.intervalcollection"), text_file=True)
    ai.peaks_to_fasta(types.SimpleNamespace(
        sequence_graph=graph_name + ".sequences",
        intervals_file_name=str(tmp_path / "testintervals.intervalcollection"),
        out_file_name=str(tmp_path / "testsequences.fasta")))
    collection = PeakCollection.from_fasta_file(str(tmp_path / "testsequences.fasta"))
    assert len(collection.intervals) == 1
    assert collection.intervals[0].sequence.lower() == "tttcccctt"


def test_concatenate_fasta_files_known_answer(tmp_path, monkeypatch):
    # reference tests/test_command_line_interface.py:80-106: two chromosomes' files, by score
    from graph_peak_caller_b200 import analysis_interface as ai
    monkeypatch.chdir(tmp_path)
    peak1, peak2 = Peak(3, 5, [1], score=3), Peak(2, 5, [1], score=7)
    write("1_sequences.fasta", ">peak1 " + peak1.to_file_line() + "\nAACC\n")
    write("2_sequences.fasta", ">peak2 " + peak2.to_file_line() + "\nCCAA\n")
    ai.concatenate_sequence_files(types.SimpleNamespace(
        chromosomes="1,2", out_file_name="out.fasta", is_summits=None, use_input_file_pattern=None))
    with open("out.fasta") as fh:
        lines = fh.readlines()
    assert Peak.from_file_line(lines[0].split(maxsplit=1)[1]) == peak2
    assert Peak.from_file_line(lines[2].split(maxsplit=1)[1]) == peak1
    assert lines[0].startswith(">peak0 ") and lines[2].startswith(">peak1 ")
    assert [lines[1], lines[3]] == ["CCAA\n", "AACC\n"]
    both = PeakCollection.from_file("out.intervalcollection", text_file=True)
    assert both.intervals == [peak2, peak1]
    # summit files through a pattern
    write("s_1.fa", ">peak0 " + peak1.to_file_line() + "\nAA\n")
    write("s_2.fa", ">peak0 " + peak2.to_file_line() + "\nCC\n>peak1 " + peak1.to_file_line() + "\nGG\n")
    got = ai.concatenate_sequence_files(types.SimpleNamespace(
        chromosomes="1,2", out_file_name="all.summits.fasta", is_summits="True",
        use_input_file_pattern="s_[chrom].fa"))
    assert [p.sequence for p in got] == ["CC", "AA", "GG"]          # stable among equal scores
    assert len(PeakCollection.from_file("all.summits.intervalcollection").intervals) == 3
    with pytest.raises(FileNotFoundError):
        ai.concatenate_sequence_files(types.SimpleNamespace(
            chromosomes="3", out_file_name="x.fasta", is_summits="True", use_input_file_pattern=None))


def test_sequence_scanner_on_random_json(tmp_path):
    # random graph chunks (junk members, key order, quoted ids, nodes without a sequence):
    # the scanner sees the (id, sequence) pairs the json module sees, or -- when a sequence
    # holds an escape -- refuses the file
    import ctypes
    from hypothesis import HealthCheck, given, settings, strategies as st
    from graph_peak_caller_b200 import _lib
    scalars = st.one_of(st.none(), st.booleans(), st.integers(-10**9, 10**9), st.text(max_size=10))
    junk = st.recursive(scalars, lambda kids: st.one_of(
        st.lists(kids, max_size=3), st.dictionaries(st.text(max_size=6), kids, max_size=3)),
        max_leaves=6)
    ident = st.one_of(st.integers(1, 2**31 - 1), st.integers(1, 2**31 - 1).map(str))
    node = st.fixed_dictionaries({"id": ident}, optional={
        "sequence": st.text("ACGTNacgtn", max_size=12), "name": st.text(max_size=4), "x": junk})
    hard_node = st.fixed_dictionaries({"id": ident, "sequence": st.text("AC\\\"\n", min_size=1, max_size=6)})
    chunk = st.fixed_dictionaries({}, optional={
        "node": st.lists(st.one_of(node, node, node, hard_node), max_size=5),
        "edge": junk, "path": junk, "sequence": st.text(max_size=5)})
    lib = _lib.load()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)

    @settings(derandomize=True, max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(st.lists(chunk, min_size=1, max_size=4), st.booleans())
    def run(chunks, reverse_keys):
        lines = []
        for c in chunks:
            c = dict(c)
            if "node" in c and reverse_keys:
                c["node"] = [dict(reversed(list(n.items()))) for n in c["node"]]
            lines.append(json.dumps(c, separators=(", ", ": ") if reverse_keys else (",", ":")))
        want = [(int(n["id"]), n.get("sequence", "")) for c in chunks for n in c.get("node", [])]
        escaped = any(json.dumps(s) != '"%s"' % s for _, s in want)
        text = np.frombuffer(("\n".join(lines) + "\n").encode("utf-8"), dtype=np.uint8)
        n, m = ctypes.c_uint64(0), ctypes.c_uint64(0)
        status = lib.gpc_parse_vg_sequences_host(vp(text), text.size, None, None, None,
                                                 ctypes.byref(n), ctypes.byref(m))
        if escaped:
            assert status == _lib.GPC_ERR_ARG
            return
        assert status == _lib.GPC_OK
        assert (n.value, m.value) == (len(want), sum(len(s) for _, s in want))
        ids = np.zeros(len(want) + 1, dtype=np.int32)
        ptr = np.zeros(len(want) + 1, dtype=np.int64)
        seq = np.zeros(int(m.value) + 1, dtype=np.uint8)
        assert lib.gpc_parse_vg_sequences_host(vp(text), text.size, vp(ids), vp(ptr), vp(seq),
                                               ctypes.byref(n), ctypes.byref(m)) == _lib.GPC_OK
        assert ids[:len(want)].tolist() == [i for i, _ in want]
        bases = seq[:int(m.value)].tobytes().decode("ascii")
        assert [bases[ptr[k]:ptr[k + 1]] for k in range(len(want))] == [s for _, s in want]
        # too little room is reported, not overrun
        if len(want) > 1:
            n.value, m.value = len(want) - 1, m.value
            assert lib.gpc_parse_vg_sequences_host(vp(text), text.size, vp(ids), vp(ptr), vp(seq),
                                                   ctypes.byref(n), ctypes.byref(m)) == _lib.GPC_ERR_CAPACITY

    run()


def test_interval_sequences_on_random_paths():
    # any path of forward or reverse nodes, any offsets: the slices of the flat array equal
    # the string built node by node
    rng = np.random.default_rng(31)
    g = make_graph(4000, 0.08, seed=9, mnp_frac=0.3)
    sg = SequenceGraph.create_empty_from_ob_graph(g)
    seqs = {}
    for i, n in enumerate(g.node_len):
        seqs[i + g.min_node] = "".join(rng.choice(list("acgt"), size=int(n)))
        if n:
            sg.set_sequence(i + g.min_node, seqs[i + g.min_node].upper())
    for _ in range(300):
        v = int(rng.integers(0, g.n_nodes))
        nodes = [v]
        for _ in range(int(rng.integers(0, 5))):
            nxt = g.succ[g.succ_ptr[v]:g.succ_ptr[v + 1]]
            if nxt.size == 0:
                break
            v = int(rng.choice(nxt))
            nodes.append(v)
        ids = [n + g.min_node for n in nodes]
        if rng.random() < 0.5:                                   # the same walk on the reverse strand
            ids = [-n for n in reversed(ids)]
        strings = [seqs[n] if n > 0 else complement(seqs[-n]) for n in ids]
        if len(strings[0]) == 0 or len(strings[-1]) == 0:
            continue
        start = int(rng.integers(0, len(strings[0])))
        end = int(rng.integers(1, len(strings[-1]) + 1))
        if len(ids) == 1 and end <= start:
            continue
        whole = "".join(strings)
        want = whole[start:len(whole) - (len(strings[-1]) - end)]
        assert sg.get_interval_sequence(Peak(start, end, ids)) == want


def test_foreign_sequence_file_is_a_warning(tmp_path):
    # a .sequences file that is not this package's (e.g. offsetbasedgraph's pickle): no FASTA, no error
    import pickle
    g = reference_graph()
    with open(tmp_path / "1.nobg", "wb") as fh:
        g.to_file(fh)
    with open(tmp_path / "1.nobg.sequences", "wb") as fh:
        pickle.dump({"sequences": ["ACGT"]}, fh)
    m = MultipleGraphsCallpeaks(["1"], [g], None, None, None, None, Reporter(str(tmp_path / "x_")),
                                sequence_retrievers=ci._SequenceGraphFiles([str(tmp_path / "1.nobg.sequences")]))
    m._write_sequences(0, _Caller(PeakCollection([Peak(0, 2, [1, 2], score=3)])))
    assert not (tmp_path / "x_1_sequences.fasta").exists()


def test_create_alignment_fasta(tmp_path):
    # callpeaks_interface.py:94-103: unique alignments (first per start position) as FASTA
    g = reference_graph()
    ci.create_ob_graph(types.SimpleNamespace(
        vg_json_file_name=write(tmp_path / "g.json", VG_TEST_GRAPH),
        out_file_name=str(tmp_path / "1.nobg")))

    def alignment(nodes, offset, reverse=False):
        first = {"node_id": nodes[0], "offset": offset}
        if reverse:
            first["is_reverse"] = True
        mapping = [{"position": first, "edit": [{"from_length": 3, "to_length": 3}]}]
        for n in nodes[1:]:
            pos = {"node_id": n}
            if reverse:
                pos["is_reverse"] = True
            mapping.append({"position": pos, "edit": [{"from_length": 2, "to_length": 2}]})
        return json.dumps({"path": {"mapping": mapping}})

    lines = [alignment([1, 2], 4), alignment([1, 3], 4), alignment([3], 1), alignment([4, 2], 0, True)]
    reads = write(tmp_path / "reads.json", "\n".join(lines) + "\n")
    out = str(tmp_path / "alignments.fasta")
    ci.create_alignment_fasta(types.SimpleNamespace(
        vg_json_file_name=reads, graph=ci.load_graph(str(tmp_path / "1.nobg")),
        data_dir=str(tmp_path), chrom="1", out_file_name=out))
    with open(out) as fh:
        got = fh.read().split("\n")
    # the second alignment starts where the first does (1:4) and is dropped
    assert got == [">1", "ccc" + "tt", ">1", "ccc", ">1", "cagt" + "aa", ""]
