"""CPU: the host side of callpeaks_interface.py -- file-name conventions,
configuration, graph / linear-map files, read-length estimation -- i.e.
everything the reference's entry points (callpeaks_interface.py:17-371) do
before the first kernel runs.  The runs themselves are in
tests/test_gpu_zz_interface.py."""
import json
import os
import types

import numpy as np
import pytest

from graph_peak_caller_b200 import callpeaks_interface as ci
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.conversion import load_graph
from graph_peak_caller_b200.custom_exceptions import GraphNotFoundException
from graph_peak_caller_b200.peakcollection import IntervalCollection
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import ingest as oingest


def args_of(**kw):
    return types.SimpleNamespace(**kw)


def write_graph_json(path, g, chunks=2):
    src = np.repeat(np.arange(g.n_nodes), np.diff(g.succ_ptr))
    with open(path, "w") as fh:
        fh.write("\n".join(oingest.graph_to_vg_json(g.node_len, g.min_node, src, g.succ,
                                                    chunks=chunks)) + "\n")
    return str(path)


def write_reads_json(path, g, reads):
    with open(path, "w") as fh:
        for i in range(len(reads)):
            fh.write(json.dumps(oingest.read_to_alignment(
                g.node_len, g.min_node, int(reads.start_off[i]), int(reads.end_off[i]),
                reads.path[reads.path_ptr[i]:reads.path_ptr[i + 1]].tolist())) + "\n")
    return str(path)


def same_graph(a, b):
    assert a.min_node == b.min_node
    for f in ("node_len", "succ_ptr", "succ", "pred_ptr", "pred"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_create_ob_graph_and_linear_map_files(tmp_path):
    g = make_graph(8000, 0.05, seed=41, mnp_frac=0.1)
    name = write_graph_json(tmp_path / "chr7.json", g)
    # default output name: up to the first dot + .nobg (preprocess_interface.py:55-56)
    out = ci.create_ob_graph(args_of(vg_json_file_name=name, out_file_name=None))
    assert out == str(tmp_path / "chr7.nobg") and os.path.isfile(out)
    same_graph(load_graph(out), g)
    other = ci.create_ob_graph(args_of(vg_json_file_name=name,
                                       out_file_name=str(tmp_path / "other.obg")))
    same_graph(load_graph(other), g)
    lm_name = str(tmp_path / "chr7_linear_map.npz")
    ci.find_or_create_linear_map(load_graph(out), lm_name)
    assert LinearMap.from_file(lm_name, g) == LinearMap.from_graph(g)
    stamp = os.path.getmtime(lm_name)
    ci.find_or_create_linear_map(load_graph(out), lm_name)        # found: left alone
    assert os.path.getmtime(lm_name) == stamp


def test_load_graph_errors(tmp_path):
    with pytest.raises(GraphNotFoundException, match="not found"):
        load_graph(str(tmp_path / "missing.nobg"))
    foreign = str(tmp_path / "foreign.nobg")
    with open(foreign, "wb") as fh:
        np.savez(fh, blocks=np.arange(4))
    with pytest.raises(GraphNotFoundException, match="create_ob_graph"):
        load_graph(foreign)


def test_read_lengths_and_estimate(tmp_path):
    g = make_graph(20000, 0.05, seed=42, mnp_frac=0.1)
    reads = make_reads(g, 500, 31, 120, seed=43)
    want = [reads.interval(i, g).length() for i in range(len(reads))]
    assert reads.lengths(g).tolist() == want
    assert int(np.median(want)) == 31 and max(want) == 31      # reads at the graph end are cut
    graph_file = ci.create_ob_graph(args_of(vg_json_file_name=write_graph_json(tmp_path / "g.json", g),
                                            out_file_name=None))
    assert ci.estimate_read_length(write_reads_json(tmp_path / "r.json", g, reads), graph_file) == 31
    text = str(tmp_path / "r.intervalcollection")
    IntervalCollection(list(reads)).to_file(text, text_file=True)
    assert ci.estimate_read_length(text, graph_file) == 31


def test_get_confiugration_and_intervals(tmp_path):
    g = make_graph(3000, 0.05, seed=44)
    reads = make_reads(g, 50, 20, 80, seed=45)
    sample = write_reads_json(tmp_path / "s.json", g, reads)
    a = args_of(control=None, linear_map=None, graph_file_name=str(tmp_path / "g.nobg"),
                fragment_length="80", read_length="20", q_value_threshold="0.01",
                sample=sample, graph=g)
    config = ci.get_confiugration(a)
    assert config.linear_map_name == str(tmp_path / "g_linear_map.npz")
    assert (config.fragment_length, config.read_length, config.has_control) == (80, 20, False)
    assert config.q_values_threshold == 0.01
    a.linear_map, a.control = "given.npz", sample
    config = ci.get_confiugration(a)
    assert config.linear_map_name == "given.npz" and config.has_control
    s, c = ci.get_intervals(a)
    assert type(s).__name__ == type(c).__name__ == "UniqueIntervals"
    assert len(s.batch) == len(c.batch) == len(reads)
    with pytest.raises(SystemExit):
        ci.parse_input_file(str(tmp_path / "nowhere.json"), g)


def test_get_file_names(tmp_path):
    for n in ("b2.nobg", "b1.nobg", "c.txt"):
        (tmp_path / n).write_text("")
    assert ci._get_file_names(None) is None
    assert ci._get_file_names([str(tmp_path / "*.nobg")]) == [str(tmp_path / "b1.nobg"),
                                                              str(tmp_path / "b2.nobg")]
    assert ci._get_file_names(["z", "a"]) == ["a", "z"]
    assert ci._get_file_names(["single"]) == ["single"]


def _two_chromosomes(tmp_path):
    out = {}
    for k, chrom in enumerate(("1", "2")):
        g = make_graph(6000 + 2000 * k, 0.04, seed=50 + k)
        ci.create_ob_graph(args_of(vg_json_file_name=write_graph_json(tmp_path / ("%s.json" % chrom), g),
                                   out_file_name=str(tmp_path / ("%s.nobg" % chrom))))
        s = make_reads(g, 400, 20, 80, seed=60 + k, peak_frac=0.7, peak_spacing=2500)
        c = make_reads(g, 600, 20, 80, seed=70 + k, peak_frac=0.0, dup_frac=0.0)
        write_reads_json(tmp_path / ("sample_%s.json" % chrom), g, s)
        write_reads_json(tmp_path / ("control_%s.json" % chrom), g, c)
        out[chrom] = (g, s, c)
    return out


def test_prepare_callpeaks2(tmp_path):
    data = _two_chromosomes(tmp_path)
    d = str(tmp_path) + "/"
    a = args_of(graph=[d + "*.nobg"], sample=[d + "sample_2.json", d + "sample_1.json"],
                control=[d + "control_*.json"], fragment_length="80", read_length=None,
                out_name=d + "run_", q_threshold="0.01", keep_duplicates="False",
                genome_size="14000", unique_reads="700", stop_after_p_values="True")
    caller = ci.prepare_callpeaks2(a)
    assert caller.graph_file_names == [d + "1.nobg", d + "2.nobg"]
    assert caller.samples == [d + "sample_1.json", d + "sample_2.json"]
    assert caller.controls == [d + "control_1.json", d + "control_2.json"]
    assert caller.names == ["1", "2"]                     # graph file names below data_dir, no extension
    assert caller.linear_maps == [d + "1_linear_map.npz", d + "2_linear_map.npz"]
    for chrom, lm in zip(("1", "2"), caller.linear_maps):
        assert LinearMap.from_file(lm, data[chrom][0]) == LinearMap.from_graph(data[chrom][0])
    config = caller._config
    assert (config.fragment_length, config.read_length) == (80, 20)       # read length estimated
    assert config.q_values_threshold == 0.01 and config.has_control and not config.keep_duplicates
    assert config.global_min == 700 * 80 / 14000
    assert caller.stop_after_p_values and caller._reporter._base_name == d + "run_"
    # one graph: no extra experiment name; no control: the sample doubles as control
    a1 = args_of(graph=[d + "1.nobg"], sample=[d + "sample_1.json"], control=None,
                 fragment_length="80", read_length="20")
    single = ci.prepare_callpeaks2(a1)
    assert single.names == [""] and single.controls == single.samples
    assert not single._config.has_control and single._config.global_min is None
    assert single._reporter._base_name == "" and not single.stop_after_p_values
    # the reference's hard stops
    with pytest.raises(SystemExit):
        ci.prepare_callpeaks2(args_of(graph=[d + "*.nobg"], sample=[d + "sample_1.json"],
                                      control=None, fragment_length="80", read_length="20"))
    with pytest.raises(SystemExit):
        ci.prepare_callpeaks2(args_of(graph=[d + "1.nobg"], sample=[d + "sample_1.json"],
                                      control=None, fragment_length="10", read_length="20"))
    # no fragment length: estimated (tests/test_shiftestimation.py); these few reads do not
    # hold the 100 paired peaks the model asks for (shiftestimation.py:112-116)
    from graph_peak_caller_b200.shiftestimation import NotEnoughPairsException
    with pytest.raises(NotEnoughPairsException):
        ci.prepare_callpeaks2(args_of(graph=[d + "1.nobg"], sample=[d + "sample_1.json"],
                                      control=None, fragment_length=None, read_length="20"))


def test_prepare_callpeaks_whole_genome(tmp_path):
    _two_chromosomes(tmp_path)
    d = str(tmp_path)
    a = args_of(chromosomes="1,2", data_dir=d, sample=d + "/sample.json", control=None,
                fragment_length="80", read_length="20", genome_size="14000", unique_reads="700",
                out_name=None, keep_duplicates="True", stop_after_p_values="False")
    caller = ci.prepare_callpeaks_whole_genome(a)
    assert caller.names == ["1", "2"]
    assert caller.graph_file_names == [d + "/1.nobg", d + "/2.nobg"]
    assert caller.samples == [d + "/sample_1.json", d + "/sample_2.json"]
    assert caller.controls == caller.samples and caller.controls is not caller.samples
    assert caller.linear_maps == [d + "/1_linear_map.npz", d + "/2_linear_map.npz"]
    assert all(os.path.isfile(f) for f in caller.linear_maps)
    assert caller._config.keep_duplicates and not caller._config.has_control
    assert caller._config.global_min == 4.0 and caller._reporter._base_name == ""
    a.sample, a.control = d + "/reads_chrom.intervalcollection", d + "/control.json"
    caller = ci.prepare_callpeaks_whole_genome(a)
    assert caller.samples == [d + "/reads_1.intervalcollection", d + "/reads_2.intervalcollection"]
    assert caller.controls == [d + "/control_1.json", d + "/control_2.json"]
    assert caller._config.has_control


def test_second_half_reads_the_reporters_files(tmp_path):
    # state after p-values as the first half leaves it on disk (reporter.py:21-23,47-49 of the
    # reference), picked up by a caller that was given no reads (multiplegraphscallpeaks.py:147-156)
    from graph_peak_caller_b200 import Configuration
    from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
    from graph_peak_caller_b200.reporter import Reporter
    from graph_peak_caller_b200.sparsediffs import SparseValues
    g = make_graph(4000, 0.05, seed=81)
    base = str(tmp_path / "wg_")
    with open(str(tmp_path / "7.nobg"), "wb") as fh:
        g.to_file(fh)
    p = SparseValues(np.array([0, 10, 25]), np.array([0.0, 3.5, 0.0]))
    p.track_size = g.size_bp
    sub = Reporter(base).get_sub_reporter("7")
    sub.add("pvalues", p)
    sub.add("touched_nodes", {3, 1, 2})
    config = Configuration()
    config.read_length, config.fragment_length = 20, 80
    caller = MultipleGraphsCallpeaks(["7"], [str(tmp_path / "7.nobg")], None, None, None, config,
                                     Reporter(base))
    assert caller.samples is None and caller.my_chromosomes() == [0]
    state = caller._caller_from_files(0)
    assert state.p_values_pileup == p and state.p_values_pileup.track_size == g.size_bp
    assert sorted(state.touched_nodes.tolist()) == [1, 2, 3]
    assert state._reporter._base_name == base + "7_" and state.direct_pileup is None
    same_graph(state.graph, g)


def test_chromosome_costs_from_file_sizes(tmp_path):
    # several ranks: graphs given as files are weighed by file size (no rank parses them all)
    from graph_peak_caller_b200 import Configuration, multigraph
    from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
    from graph_peak_caller_b200.reporter import Reporter
    files, graphs = [], []
    for k, bp in enumerate((3000, 9000, 5000)):
        g = make_graph(bp, 0.05, seed=90 + k)
        files.append(write_graph_json(tmp_path / ("c%d.json" % k), g, chunks=1))
        graphs.append(g)
    config = Configuration()
    config.read_length, config.fragment_length = 20, 80
    by_file = MultipleGraphsCallpeaks(["a", "b", "c"], files, None, None, None, config, Reporter(None))
    costs = by_file._costs()
    assert costs == [os.path.getsize(f) for f in files] and by_file._graphs == {}
    by_graph = MultipleGraphsCallpeaks(["a", "b", "c"], graphs, None, None, None, config,
                                       Reporter(None))
    assert by_graph._costs() == [g.size_bp for g in graphs]
    # both weights pack these three the same way onto two ranks
    assert multigraph.assign_chromosomes(costs, 2) == \
        multigraph.assign_chromosomes(by_graph._costs(), 2) == [[1], [2, 0]]


def test_reporter_file_contract(tmp_path):
    # the file names of the reference's Reporter (reporter.py:17-57)
    from graph_peak_caller_b200.peakcollection import Peak, PeakCollection
    from graph_peak_caller_b200.reporter import Reporter
    from graph_peak_caller_b200.sparsediffs import SparseValues
    g = make_graph(2000, 0.05, seed=77)
    base = str(tmp_path / "run_")
    rep = Reporter(base).get_sub_reporter("chr3")
    assert rep._base_name == base + "chr3_" and Reporter(base).get_sub_reporter("")._base_name == base
    track = SparseValues(np.array([0, 5, 9]), np.array([0.0, 2.0, 0.0]))
    track.track_size = g.size_bp
    for name in ("qvalues", "pvalues", "hole_cleaned", "fragment_pileup", "background_track",
                 "direct_pileup", "thresholded"):
        rep.add(name, track)
    peaks = [Peak.from_flat(1, 3, [1, 2], g, 2.5, (False, False), 5)]
    rep.add("all_max_paths", peaks)
    rep.add("max_paths", peaks)
    rep.add("touched_nodes", {2, 1})
    written = sorted(os.listdir(tmp_path))
    want = ["run_chr3_%s_%s.npy" % (n, k) for n in ("qvalues", "pvalues", "hole_cleaned",
            "fragment_pileup", "background_track", "direct_pileup") for k in ("indexes", "values")]
    want += ["run_chr3_all_max_paths.intervalcollection", "run_chr3_max_paths.intervalcollection",
             "run_chr3_touched_nodes.npy"]
    assert written == sorted(want)                    # "thresholded" is kept but not written
    assert rep.memory["thresholded"] is track
    assert SparseValues.from_sparse_files(base + "chr3_qvalues") == track
    back = PeakCollection.from_file(base + "chr3_max_paths.intervalcollection", graph=g)
    assert back.intervals[0] == peaks[0] and back.intervals[0].score == 2.5
    assert sorted(np.load(base + "chr3_touched_nodes.npy").tolist()) == [1, 2]
    quiet = Reporter(None)
    quiet.add("qvalues", track)
    assert quiet.memory["qvalues"] is track and quiet.get_sub_reporter("x")._base_name is None


def test_helpers_equal_the_live_reference(tmp_path):
    # callpeaks_interface._get_file_names / get_confiugration of the unmodified reference
    from oracle import refbridge
    if not refbridge.available():
        pytest.skip("/root/reference not present")
    refbridge._ref()
    from graph_peak_caller import callpeaks_interface as ref
    for n in ("b2.nobg", "b1.nobg", "a.json"):
        (tmp_path / n).write_text("")
    for pattern in (None, [str(tmp_path / "*.nobg")], ["z", "a", "m"], ["only"],
                    [str(tmp_path / "nothing*")]):
        assert ci._get_file_names(pattern) == ref._get_file_names(pattern)
    for control, linear_map in ((None, None), ("c.json", "lm.npz")):
        a = args_of(control=control, linear_map=linear_map, graph_file_name="dir/chr1.nobg",
                    fragment_length="135", read_length="36", q_value_threshold=None)
        mine, theirs = ci.get_confiugration(a), ref.get_confiugration(a)
        for field in ("has_control", "linear_map_name", "fragment_length", "read_length",
                      "q_values_threshold", "global_min", "keep_duplicates"):
            assert getattr(mine, field) == getattr(theirs, field), field


class _Recorder:
    """Stands in for the reference's MultipleGraphsCallpeaks: keeps what it was given."""
    last = None

    def __init__(self, names, graphs, samples, controls, linear_maps, config, reporter,
                 sequence_retrievers=None, stop_after_p_values=False, **kw):
        type(self).last = dict(names=list(names), graphs=list(graphs), samples=list(samples),
                               controls=list(controls), linear_maps=list(linear_maps),
                               config=dict(config.__dict__), base=reporter._base_name,
                               stop=stop_after_p_values)

    def run(self):
        pass


def _seen(caller):
    return dict(names=caller.names, graphs=caller.graph_file_names, samples=caller.samples,
                controls=caller.controls, linear_maps=caller.linear_maps,
                config=dict(caller._config.__dict__), base=caller._reporter._base_name,
                stop=caller.stop_after_p_values)


def test_run_preparation_equals_the_live_reference(tmp_path, monkeypatch):
    # what run_callpeaks2 / run_callpeaks_whole_genome of the unmodified reference hand to
    # MultipleGraphsCallpeaks, against what the entry points here hand to theirs
    from oracle import refbridge
    if not refbridge.available():
        pytest.skip("/root/reference not present")
    refbridge._ref()
    from graph_peak_caller import callpeaks_interface as ref
    monkeypatch.setattr(ref, "MultipleGraphsCallpeaks", _Recorder)
    _two_chromosomes(tmp_path)
    monkeypatch.chdir(tmp_path)
    for chrom in ("1", "2"):                    # present, so that neither side has to build them
        ci.create_linear_map(load_graph(chrom + ".nobg"), chrom + "_linear_map.npz")
    cases = [
        dict(graph=["*.nobg"], sample=["sample_1.json", "sample_2.json"],
             control=["control_2.json", "control_1.json"], fragment_length="80", read_length="20",
             out_name="x_", q_threshold="0.2", keep_duplicates="True", genome_size="14000",
             unique_reads="700", stop_after_p_values="True", min_fold_enrichment=None,
             max_fold_enrichment=None),
        dict(graph=["1.nobg"], sample=["sample_1.json"], control=None, fragment_length="80",
             read_length="20", out_name=None, q_threshold=None, keep_duplicates="False",
             genome_size=None, unique_reads=None, stop_after_p_values="False",
             min_fold_enrichment=None, max_fold_enrichment=None),
    ]
    for case in cases:
        ref.run_callpeaks2(args_of(**case))
        assert _seen(ci.prepare_callpeaks2(args_of(**case))) == _Recorder.last
    whole = [
        dict(chromosomes="1,2", data_dir=".", sample="sample.json", control="control.json",
             fragment_length="80", read_length="20", genome_size="14000", unique_reads="700",
             out_name="wg_", keep_duplicates="False", stop_after_p_values="True"),
        dict(chromosomes="2", data_dir=".", sample="reads_chrom.intervalcollection", control=None,
             fragment_length="80", read_length="20", genome_size="14000", unique_reads="700",
             out_name=None, keep_duplicates="True", stop_after_p_values="False"),
    ]
    for case in whole:
        ref.run_callpeaks_whole_genome(args_of(**case))
        assert _seen(ci.prepare_callpeaks_whole_genome(args_of(**case))) == _Recorder.last


def test_create_linear_map_interface(tmp_path):
    # preprocess_interface.py:72-80
    g = make_graph(5000, 0.05, seed=3)
    name = str(tmp_path / "g.nobg")
    with open(name, "wb") as fh:
        g.to_file(fh)
    out = ci.create_linear_map_interface(args_of(graph=name, graph_file_name=name,
                                                 out_file_base_name=None))
    assert out == str(tmp_path / "g_linear_map.npz")
    assert LinearMap.from_file(out, g) == LinearMap.from_graph(g)
    out = ci.create_linear_map_interface(args_of(graph=g, graph_file_name=name,
                                                 out_file_base_name=str(tmp_path / "other.npz")))
    assert LinearMap.from_file(str(tmp_path / "other.npz"), g) == LinearMap.from_graph(g)


def test_reporter_writes_the_touched_mask_as_ids(tmp_path):
    # what CallPeaks reports is the device mask's set-like view; the file holds node ids
    import torch
    from graph_peak_caller_b200.reporter import Reporter
    from graph_peak_caller_b200.sample.sparsegraphpileup import TouchedNodes
    touched = TouchedNodes(torch.tensor([1, 0, 0, 1, 1], dtype=torch.uint8), 10)
    Reporter(str(tmp_path / "t_")).add("touched_nodes", touched)
    got = np.load(str(tmp_path / "t_touched_nodes.npy"))
    assert got.tolist() == [10, 13, 14] and got.dtype == np.dtype("int")
    assert set(touched) == {10, 13, 14} and 13 in touched and 11 not in touched
