"""CPU: the native vg JSON parsers (gpc_parse_vg_alignments_host,
gpc_parse_vg_graph_host) against the json-module restatement
of pyvg's conversion in oracle/ingest.py, on the reference's own known answers
and fixtures (copied as data under tests/golden/) and on synthetic files."""
import json
import os
import time

import numpy as np
import pytest

from graph_peak_caller_b200 import conversion
from graph_peak_caller_b200.custom_exceptions import IntervalNotInGraphException
from graph_peak_caller_b200.intervals import Intervals, UniqueIntervals
from graph_peak_caller_b200.reads import ReadBatch
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import ingest as oingest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("start_node", "start_off", "end_node", "end_off", "path_ptr", "path")


def same_batch(got, want):
    for f in FIELDS:
        w = want[f] if isinstance(want, dict) else getattr(want, f)
        assert np.array_equal(getattr(got, f), w), f


def write_lines(path, lines, end="\n"):
    with open(path, "w") as fh:
        for line in lines:
            fh.write(line + end)
    return str(path)


def test_known_answer_of_the_reference_examples(tmp_path):
    # reference tests/examples.py:5-31: position {"offset": 10, "node_id": 0}, edits of
    # from_length 4 and 1 -> Interval(10, 15, [0]); position {"offset": 0, "node_id": 1,
    # "is_reverse": true} is a reverse mapping (tests/test_vg.py:27-29)
    edits = [{"to_length": 4, "from_length": 4}, {"to_length": 1, "from_length": 1, "sequence": "N"}]
    lines = [json.dumps({"path": {"name": "path1", "mapping": [
                {"position": {"offset": 10, "node_id": 0}, "edit": edits}]}}),
             json.dumps({"path": {"mapping": [
                 {"position": {"offset": 0, "node_id": 1, "is_reverse": True}, "edit": edits}]}})]
    got = ReadBatch.from_vg_json_file(write_lines(tmp_path / "a.json", lines))
    assert got.start_off.tolist() == [10, 0] and got.end_off.tolist() == [15, 5]
    assert got.path.tolist() == [0, -1] and got.path_ptr.tolist() == [0, 1, 2]
    assert got.start_node.tolist() == [0, -1] and got.end_node.tolist() == [0, -1]
    same_batch(got, oingest.vg_json_lines_to_reads(lines))


def test_reference_fixture_matches_the_restatement():
    name = os.path.join(GOLDEN, "vg_alignments_15.json")   # reference tests/node_range_test_data/
    with open(name) as fh:
        want = oingest.vg_json_lines_to_reads(fh)
    got = ReadBatch.from_vg_json_file(name)
    assert len(got) == 15 and got.n_without_path == 0
    same_batch(got, want)
    # the first alignment by hand: reverse strand, 8 mappings, offset 1, last edit 3 long
    assert got.path[:8].tolist() == [-8235641, -8235640, -8235638, -8235637, -8235635, -8235634,
                                     -8235632, -8235631]
    assert (got.start_off[0], got.end_off[0]) == (1, 3)


def test_split_into_chromosomes_as_the_reference_test(tmp_path):
    # reference tests/test_command_line_interface.py:141-149: three alignments per chromosome
    ranges = {"1": "1:9029588", "2": "9029589:15477266", "3": "15477267:23090691",
              "4": "23090692:29187850", "5": "29187851:37071876"}   # tests/node_range_test_data/
    for chrom, text in ranges.items():
        (tmp_path / ("node_range_%s.txt" % chrom)).write_text(text)
    with open(os.path.join(GOLDEN, "vg_alignments_15.json")) as fh:
        lines = fh.read().splitlines()
    reads = write_lines(tmp_path / "vg_alignments.json", lines)
    unmapped = conversion.split_vg_json_reads_into_chromosomes("1,2,3,4,5", reads,
                                                               str(tmp_path) + "/")
    assert unmapped == 0
    # the reference's own calling form (an args namespace) gives the same files
    import types
    from graph_peak_caller_b200 import callpeaks_interface
    first = {c: (tmp_path / ("vg_alignments_%s.json" % c)).read_bytes() for c in ranges}
    assert callpeaks_interface.split_vg_json_reads_into_chromosomes(types.SimpleNamespace(
        chromosomes="1,2,3,4,5", vg_json_reads_file_name=reads,
        range_files_base_name=str(tmp_path) + "/")) == 0
    assert first == {c: (tmp_path / ("vg_alignments_%s.json" % c)).read_bytes() for c in ranges}
    for chrom, text in ranges.items():
        part = ReadBatch.from_vg_json_file(str(tmp_path / ("vg_alignments_%s.json" % chrom)))
        lo, hi = (int(v) for v in text.split(":"))
        assert len(part) == 3
        assert np.all((np.abs(part.path) >= lo) & (np.abs(part.path) <= hi))


@pytest.mark.parametrize("quoted", [False, True])
def test_synthetic_reads_round_trip(tmp_path, quoted):
    g = make_graph(30000, 0.05, seed=11, mnp_frac=0.1)
    reads = make_reads(g, 4000, 30, 100, seed=12)
    lines = []
    for i in range(len(reads)):
        nodes = reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist()
        lines.append(json.dumps(oingest.read_to_alignment(
            g.node_len, g.min_node, int(reads.start_off[i]), int(reads.end_off[i]), nodes,
            quoted=quoted, name="read %d {[\"" % i)))
        if i % 7 == 0:
            lines.append(json.dumps({"sequence": "ACGT", "name": "unaligned"}))
        if i % 11 == 0:
            lines.append("")
        if i % 13 == 0:      # a path somewhere else than at the top level is not the read's path
            lines.append(json.dumps({"fragment_next": {"path": {"mapping": [
                {"position": {"node_id": 1}}]}}, "path": {"mapping": []}}))
    name = write_lines(tmp_path / "reads.json", lines)
    want = oingest.vg_json_lines_to_reads(lines)
    for threads in (0, 1, 3):
        got = ReadBatch.from_vg_json_file(name, graph=g, n_threads=threads)
        same_batch(got, want)
        same_batch(got, reads)
        assert got.n_without_path == want["n_without_path"] > 0
    # by file name, as the reference's parse_input_file: *.json holds vg alignments
    same_batch(Intervals(name).batch, reads)
    same_batch(conversion.vg_json_file_to_interval_collection(name, g).intervals, reads)
    assert sum(1 for _ in UniqueIntervals(name)) == sum(1 for _ in UniqueIntervals(reads))


def test_large_file_uses_all_chunks(tmp_path):
    g = make_graph(100000, 0.02, seed=13)
    reads = make_reads(g, 20000, 36, 200, seed=14)
    lines = [json.dumps(oingest.read_to_alignment(
        g.node_len, g.min_node, int(reads.start_off[i]), int(reads.end_off[i]),
        reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist())) for i in range(len(reads))]
    name = write_lines(tmp_path / "big.json", lines)
    assert os.path.getsize(name) > (1 << 20)           # above the single-thread cut-off
    t0 = time.perf_counter()
    got = ReadBatch.from_vg_json_file(name, n_threads=4)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = oingest.vg_json_lines_to_reads(lines)
    t_json = time.perf_counter() - t0
    same_batch(got, want)
    same_batch(got, reads)
    assert t_native * 3 < t_json, (t_native, t_json)


def test_node_outside_the_graph(tmp_path):
    g = make_graph(2000, 0.05, seed=15)
    ok = json.dumps(oingest.read_to_alignment(g.node_len, g.min_node, 0, 1, [1]))
    bad = json.dumps({"path": {"mapping": [{"position": {"node_id": g.n_nodes + 5},
                                            "edit": [{"from_length": 3}]}]}})
    name = write_lines(tmp_path / "x.json", [ok, bad])
    assert len(ReadBatch.from_vg_json_file(name)) == 2          # no graph, no check
    with pytest.raises(IntervalNotInGraphException, match="line 2"):
        ReadBatch.from_vg_json_file(name, graph=g)


@pytest.mark.parametrize("line", [
    '{"path": {"mapping": [{"position": {"offset": 3}, "edit": []}]}}',          # no node id
    '{"path": {"mapping": [{"position": {"node_id": 4}, "edit": [{"from_length": 2}]}',
    '{"path": {"mapping": [{"position": {"node_id": 1.5}}]}}',
    '{"path": {"mapping": [{"position": {"node_id": 0, "is_reverse": true}}]}}',
    '["path"]'])
def test_malformed_alignments_are_rejected(tmp_path, line):
    name = write_lines(tmp_path / "bad.json", ['{"path": {"mapping": []}}', line])
    with pytest.raises(Exception, match="malformed line"):
        ReadBatch.from_vg_json_file(name)


def test_empty_and_pathless_files(tmp_path):
    empty = tmp_path / "none.json"
    empty.write_text("")
    got = ReadBatch.from_vg_json_file(str(empty))
    assert len(got) == 0 and got.path_ptr.tolist() == [0] and got.n_without_path == 0
    name = write_lines(tmp_path / "unaligned.json", ['{"sequence": "A"}', '{"path": {}}'], end="\r\n")
    got = ReadBatch.from_vg_json_file(name)
    assert len(got) == 0 and got.n_without_path == 2


# ---- graphs -------------------------------------------------------------------------

def test_reference_graph_fixture(tmp_path):
    # reference tests/vg_test_graph.json and the graph its CLI test expects from it
    # (tests/test_command_line_interface.py:18-27): blocks 7, 4, 7, 4; edges 1->2, 1->3, 2->4, 3->4
    text = ('{"node": [     {"id": 1, "sequence": "TTTCCCC"},     {"id": 2, "sequence": "TTTT"},'
            '     {"id": 3, "sequence": "CCCCTTT"},     {"id": 4, "sequence": "ACTG"} ],'
            '"edge": [     {"from": 1, "to": 2},     {"from": 1, "to": 3},     {"from": 2, "to": 4},'
            '    {"from": 3, "to": 4} ] }')
    name = write_lines(tmp_path / "g.json", [text])
    g = conversion.json_file_to_obg_numpy_graph(name, 0)
    assert g.min_node == 1 and g.node_len.tolist() == [7, 4, 7, 4]
    assert {k: g.adj_list[k] for k in (1, 2, 3, 4)} == {1: [2, 3], 2: [4], 3: [4], 4: []}
    assert g.reverse_adj_list[-4] == [-2, -3]


@pytest.mark.parametrize("chunks,min_node", [(1, 1), (5, 1), (3, 1000)])
def test_synthetic_graph_round_trip(tmp_path, chunks, min_node):
    g0 = make_graph(20000, 0.05, seed=21, mnp_frac=0.1)
    src = np.repeat(np.arange(g0.n_nodes), np.diff(g0.succ_ptr))
    lines = oingest.graph_to_vg_json(g0.node_len, min_node, src, g0.succ, chunks=chunks)
    name = write_lines(tmp_path / "g.json", lines)
    g = conversion.json_file_to_obg_numpy_graph(name)
    assert g.min_node == min_node
    for f in ("node_len", "node_off", "succ_ptr", "succ", "pred_ptr", "pred"):
        assert np.array_equal(getattr(g, f), getattr(g0, f)), f
    ids, lens, a, b = oingest.vg_graph_lines_to_lists(lines)
    assert sorted(zip(a, b)) == sorted(zip((src + min_node).tolist(), (g0.succ + min_node).tolist()))
    assert lens == g0.node_len.tolist()


def test_graph_details(tmp_path):
    # quoted 64-bit ids, keys in another order, an escape inside a skipped string, a hole in the ids
    lines = ['{"path": [{"name": "a \\"b\\" ]}", "mapping": [{"position": {"node_id": "7"}}]}],'
             ' "edge": [{"to": "12", "from": "10", "overlap": 0}],'
             ' "node": [{"sequence": "ACG", "name": "x", "id": "10"}, {"id": "12", "sequence": ""}]}',
             '',
             '{"node": [{"id": 13, "sequence": "ACGTACGT"}], "edge": [{"from": 12, "to": 13}]}']
    g = conversion.json_file_to_obg_numpy_graph(write_lines(tmp_path / "g.json", lines))
    assert g.min_node == 10 and g.node_len.tolist() == [3, 0, 0, 8]
    assert g.adj_list[10] == [12] and g.adj_list[12] == [13]


def test_graph_rejections(tmp_path):
    rev = '{"node": [{"id": 1, "sequence": "A"}, {"id": 2, "sequence": "C"}], ' \
          '"edge": [{"from": 1, "to": 2, "to_end": true}]}'
    with pytest.raises(ValueError, match="reversing edges"):
        conversion.json_file_to_obg_numpy_graph(write_lines(tmp_path / "r.json", [rev]))
    back = '{"node": [{"id": 1, "sequence": "A"}, {"id": 2, "sequence": "C"}], ' \
           '"edge": [{"from": 2, "to": 1}]}'
    with pytest.raises(ValueError, match="ascend"):
        conversion.json_file_to_obg_numpy_graph(write_lines(tmp_path / "b.json", [back]))
    with pytest.raises(Exception, match="malformed"):
        conversion.json_file_to_obg_numpy_graph(write_lines(tmp_path / "m.json", ['{"node": [{"id": 1']))
    with pytest.raises(ValueError, match="no nodes"):
        conversion.json_file_to_obg_numpy_graph(write_lines(tmp_path / "e.json", ['{"edge": []}']))


def test_multigraph_reads_the_references_file_kinds(tmp_path):
    # reference multiplegraphscallpeaks.py:78-104: *.intervalcollection or vg JSON per chromosome
    from graph_peak_caller_b200 import Configuration
    from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
    from graph_peak_caller_b200.peakcollection import IntervalCollection
    from graph_peak_caller_b200.reporter import Reporter
    g0 = make_graph(5000, 0.05, seed=31)
    reads = make_reads(g0, 300, 20, 80, seed=32)
    src = np.repeat(np.arange(g0.n_nodes), np.diff(g0.succ_ptr))
    graph_file = write_lines(tmp_path / "chr1.json",
                             oingest.graph_to_vg_json(g0.node_len, 1, src, g0.succ, chunks=2))
    vg_reads = write_lines(tmp_path / "reads_chr1.json", [
        json.dumps(oingest.read_to_alignment(
            g0.node_len, 1, int(reads.start_off[i]), int(reads.end_off[i]),
            reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist())) for i in range(len(reads))])
    text_reads = str(tmp_path / "reads_chr1.intervalcollection")
    IntervalCollection(list(reads)).to_file(text_reads, text_file=True)
    reads.to_file(str(tmp_path / "reads_chr1.npz"))
    caller = MultipleGraphsCallpeaks(["chr1"], [graph_file], [vg_reads], [text_reads], [None],
                                     Configuration(), Reporter(None))
    g = caller._graph(0)
    assert np.array_equal(g.node_len, g0.node_len) and np.array_equal(g.succ, g0.succ)
    sample, control = caller.get_intervals(vg_reads, text_reads, g)
    assert isinstance(sample, UniqueIntervals) and isinstance(control, UniqueIntervals)
    same_batch(sample.batch, reads)
    same_batch(control.batch, reads)
    same_batch(caller.get_intervals(str(tmp_path / "reads_chr1.npz"), sample, g)[0].batch, reads)


def test_ingest_rate_tool_reports_what_bench_prints():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    import ingest_rate
    out = ingest_rate.measure(n_reads=3000, repeats=1, json_reads=500)
    assert out["reads"] == 3000 and out["reads_per_sec"] > 0 and out["json_module_reads_per_sec"] > 0
    json.dumps(out)
    # the generated lines are what the restatement reads too
    lines = ingest_rate.alignment_lines(200)
    want = oingest.vg_json_lines_to_reads(lines)
    assert len(want["start_node"]) == 200


def test_splitter_equals_the_live_reference(tmp_path):
    # preprocess_interface.split_vg_json_reads_into_chromosomes of the unmodified reference
    # (build container only) writes byte-identical files
    import shutil
    import types
    from oracle import refbridge
    if not refbridge.available():
        pytest.skip("/root/reference not present")
    refbridge._ref()
    from graph_peak_caller import preprocess_interface
    extra = [json.dumps({"path": {"mapping": [{"position": {"node_id": str(k) if k % 2 else k},
                                               "edit": [{"from_length": 1}]}]}})
             for k in (5, 15477266, 15477267)] + ['{"sequence": "ACGT"}']
    # (a node beyond every range makes the reference look up the limits of "unmapped" and
    # fail with KeyError, preprocess_interface.py:102,113-114; here such a line goes to the
    # unmapped file as the reference's log message intends -- covered below, not compared)
    with open(os.path.join(GOLDEN, "vg_alignments_15.json")) as fh:
        lines = fh.read().splitlines() + extra
    ranges = ["1:9029588", "9029589:15477266", "15477267:23090691", "23090692:29187850",
              "29187851:37071876"]
    outs = []
    for side in ("mine", "reference"):
        d = tmp_path / side
        d.mkdir()
        for i, text in enumerate(ranges, 1):
            (d / ("node_range_%d.txt" % i)).write_text(text)
        reads = write_lines(d / "vg_alignments.json", lines)
        if side == "mine":
            unmapped = conversion.split_vg_json_reads_into_chromosomes("1,2,3,4,5", reads,
                                                                       str(d) + "/")
            assert unmapped == 1
        else:
            preprocess_interface.split_vg_json_reads_into_chromosomes(types.SimpleNamespace(
                chromosomes="1,2,3,4,5", vg_json_reads_file_name=reads,
                range_files_base_name=str(d) + "/"))
        outs.append({f: (d / f).read_bytes() for f in sorted(os.listdir(d))})
    assert outs[0] == outs[1] and len(outs[0]) == 12
    beyond = write_lines(tmp_path / "mine" / "far.json", [
        json.dumps({"path": {"mapping": [{"position": {"node_id": 40000000}}]}})])
    assert conversion.split_vg_json_reads_into_chromosomes("1,2", beyond, str(tmp_path / "mine") + "/") == 1
    assert (tmp_path / "mine" / "far_unmapped.json").read_text().count("\n") == 1


def test_scanner_on_random_json(tmp_path):
    # random junk around the path (nested containers, escapes, unicode, numbers in every
    # notation the json module writes), random key order and separators: the structure-aware
    # scanner must see what the json module sees
    from hypothesis import HealthCheck, given, settings, strategies as st
    scalars = st.one_of(st.none(), st.booleans(), st.integers(-10**12, 10**12),
                        st.floats(allow_nan=False, allow_infinity=False), st.text(max_size=12))
    junk = st.recursive(scalars, lambda kids: st.one_of(
        st.lists(kids, max_size=3), st.dictionaries(st.text(max_size=8), kids, max_size=3)),
        max_leaves=8)
    position = st.fixed_dictionaries(
        {"node_id": st.integers(1, 2**31 - 1)},
        optional={"offset": st.integers(0, 5000), "is_reverse": st.booleans(), "name": st.text(max_size=5)})
    edit = st.fixed_dictionaries({}, optional={"from_length": st.integers(0, 300),
                                                "to_length": st.integers(0, 300),
                                                "sequence": st.text("ACGT", max_size=6)})
    mapping = st.fixed_dictionaries({"position": position},
                                    optional={"edit": st.lists(edit, max_size=3),
                                              "rank": st.integers(1, 50), "extra": junk})
    alignment = st.fixed_dictionaries(
        {}, optional={"path": st.fixed_dictionaries({}, optional={
            "mapping": st.lists(mapping, max_size=4), "name": st.text(max_size=6)}),
            "sequence": st.text(max_size=20), "refpos": junk, "fragment_next": junk,
            "annotation": junk, "mapping": junk, "position": junk})

    def dump(obj, rng_bits):
        keys = list(obj)
        if rng_bits & 1:
            keys.reverse()
        sep = ((",", ":"), (", ", ": "), (" ,  ", " :  "))[(rng_bits >> 1) % 3]
        return json.dumps({k: obj[k] for k in keys}, separators=sep,
                          ensure_ascii=bool(rng_bits & 8))

    name = str(tmp_path / "fuzz.json")

    @settings(derandomize=True, max_examples=150, deadline=None,
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(st.lists(st.tuples(alignment, st.integers(0, 15)), min_size=1, max_size=6))
    def run(items):
        lines = [dump(obj, bits) for obj, bits in items]
        with open(name, "w", encoding="utf-8") as fh:
            fh.write("\n".join(lines) + "\n")
        want = oingest.vg_json_lines_to_reads(lines)
        got = ReadBatch.from_vg_json_file(name)
        same_batch(got, want)
        assert got.n_without_path == want["n_without_path"]

    run()


def test_graph_scanner_on_random_json(tmp_path):
    import ctypes
    from hypothesis import HealthCheck, given, settings, strategies as st
    from graph_peak_caller_b200 import _lib
    scalars = st.one_of(st.none(), st.booleans(), st.integers(-10**9, 10**9), st.text(max_size=10))
    junk = st.recursive(scalars, lambda kids: st.one_of(
        st.lists(kids, max_size=3), st.dictionaries(st.text(max_size=6), kids, max_size=3)),
        max_leaves=6)
    ident = st.one_of(st.integers(1, 2**31 - 1), st.integers(1, 2**31 - 1).map(str))
    node = st.fixed_dictionaries({"id": ident}, optional={
        "sequence": st.text("ACGTN\\\"é", max_size=12), "name": st.text(max_size=4)})
    edge = st.fixed_dictionaries({"from": ident, "to": ident}, optional={
        "from_start": st.booleans(), "to_end": st.booleans(), "overlap": st.integers(0, 3)})
    chunk = st.fixed_dictionaries({}, optional={"node": st.lists(node, max_size=5),
                                                "edge": st.lists(edge, max_size=5),
                                                "path": junk, "other": junk})
    name = str(tmp_path / "fuzz_graph.json")
    lib = _lib.load()

    @settings(derandomize=True, max_examples=120, deadline=None,
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(st.lists(chunk, min_size=1, max_size=4), st.booleans())
    def run(chunks, ascii_only):
        lines = [json.dumps(c, ensure_ascii=ascii_only) for c in chunks]
        ids, lens, src, dst = oingest.vg_graph_lines_to_lists(lines)
        text = np.frombuffer(("\n".join(lines) + "\n").encode("utf-8"), dtype=np.uint8)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        n, m = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _lib.check(lib.gpc_parse_vg_graph_host(vp(text), text.size, None, None, ctypes.byref(n),
                                               None, None, ctypes.byref(m)))
        assert (n.value, m.value) == (len(ids), len(src))
        out = [np.zeros(max(k, 1), dtype=np.int32) for k in (len(ids), len(ids), len(src), len(src))]
        _lib.check(lib.gpc_parse_vg_graph_host(vp(text), text.size, vp(out[0]), vp(out[1]),
                                               ctypes.byref(n), vp(out[2]), vp(out[3]),
                                               ctypes.byref(m)))
        assert out[0][:len(ids)].tolist() == ids and out[2][:len(src)].tolist() == src
        assert out[3][:len(src)].tolist() == dst
        if ascii_only:          # escapes count as one character each; raw UTF-8 counts bytes
            assert out[1][:len(ids)].tolist() == lens

    run()


def test_native_splitter_on_random_lines(tmp_path):
    # the native splitter against the reference's regular expression (oracle restatement) on
    # lines built from the fragments that regular expression cares about
    from hypothesis import HealthCheck, given, settings, strategies as st
    pieces = st.sampled_from([b'node_id":', b'node_id": ', b'node_id":  ', b'node_id":"', b'node_id": "',
                              b'"node_id": -', b'node_id', b'":', b'"', b' ', b'\t', b'\r', b'7', b'42',
                              b'150', b'99999999999999999999999', b'{', b'}', b'x', b'\n', b'\n\n',
                              b'{"path": {"mapping": [{"position": {"node_id": "12"}}]}}\n',
                              b'{"path": {"mapping": [{"position": {"offset": 3, "node_id":120}}]}}\n'])
    limits = [(1, 10), (5, 50), (100, 200)]                       # overlapping: the first range wins
    for i, (lo, hi) in enumerate(limits, 1):
        (tmp_path / ("node_range_%d.txt" % i)).write_text("%d:%d" % (lo, hi))

    @settings(derandomize=True, max_examples=200, deadline=None,
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(st.lists(pieces, max_size=40))
    def run(parts):
        data = b"".join(parts)
        with open(tmp_path / "r.json", "wb") as fh:
            fh.write(data)
        want = oingest.split_by_chromosome(data, limits)
        unmapped = conversion.split_vg_json_reads_into_chromosomes("1,2,3", str(tmp_path / "r.json"),
                                                                   str(tmp_path) + "/")
        got = [(tmp_path / ("r_%s.json" % c)).read_bytes() for c in ("1", "2", "3", "unmapped")]
        assert got == want
        assert unmapped == len(oingest._split_newlines(want[-1])) if want[-1] else unmapped == 0

    run()


def test_concurrent_parses_do_not_mix(tmp_path):
    # the library keeps one pending parse between the counting and the filling call; calls
    # from several Python threads (ctypes releases the GIL) must still each get their own file
    import threading
    g = make_graph(20000, 0.05, seed=31, mnp_frac=0.1)
    files, wants = [], []
    for k in range(4):
        reads = make_reads(g, 3000 + 500 * k, 30, 100, seed=40 + k)
        lines = [json.dumps(oingest.read_to_alignment(
            g.node_len, g.min_node, int(reads.start_off[i]), int(reads.end_off[i]),
            reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist())) for i in range(len(reads))]
        files.append(write_lines(tmp_path / ("r%d.json" % k), lines))
        wants.append(reads)
    results = [None] * 16

    def work(j):
        results[j] = ReadBatch.from_vg_json_file(files[j % 4], g)

    threads = [threading.Thread(target=work, args=(j,)) for j in range(16)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for j, got in enumerate(results):
        want = wants[j % 4]
        for f in ("start_node", "start_off", "end_node", "end_off", "path_ptr", "path"):
            assert np.array_equal(getattr(got, f), getattr(want, f)), (j, f)
