"""GPU: the reference's args-driven entry points (callpeaks_interface.py:80-371)
from files on disk -- vg graph JSON, vg alignment JSON, linear-map and Reporter
files -- against the same runs made in memory through CallPeaks /
MultipleGraphsCallpeaks (which the other test modules hold against the oracle).
Named to run after the other GPU modules."""
import os

import numpy as np
import pytest

from graph_peak_caller_b200 import CallPeaks, Configuration
from graph_peak_caller_b200 import callpeaks_interface as ci
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.intervals import UniqueIntervals
from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
from graph_peak_caller_b200.peakcollection import PeakCollection
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from test_callpeaks_interface import args_of, write_graph_json, write_reads_json

pytestmark = pytest.mark.gpu
R, F = 20, 80


def shape(peaks):
    return [(p.start_position.offset, p.end_position.offset, list(p.region_paths)) for p in peaks]


def same_peaks(got, want):
    assert shape(got) == shape(want)
    assert np.allclose([p.score for p in got], [p.score for p in want], rtol=1e-9, atol=0)


def chromosomes_on_disk(tmp_path, names=("1", "2")):
    data = {}
    for k, chrom in enumerate(names):
        g = make_graph(15000 + 6000 * k, 0.03, seed=90 + k, mnp_frac=0.1)
        ci.create_ob_graph(args_of(
            vg_json_file_name=write_graph_json(tmp_path / ("%s.json" % chrom), g),
            out_file_name=str(tmp_path / ("%s.nobg" % chrom))))
        s = make_reads(g, 1500 + 400 * k, R, F, seed=100 + k, peak_frac=0.7, peak_spacing=3000)
        c = make_reads(g, 3000, R, F, seed=110 + k, peak_frac=0.0, dup_frac=0.0)
        write_reads_json(tmp_path / ("sample_%s.json" % chrom), g, s)
        write_reads_json(tmp_path / ("control_%s.json" % chrom), g, c)
        data[chrom] = (g, s, c)
    return data


def in_memory(data, global_min=None):
    names = list(data)
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, True
    config.global_min = global_min
    caller = MultipleGraphsCallpeaks(
        names, [data[n][0] for n in names], [data[n][1] for n in names],
        [data[n][2] for n in names], [LinearMap.from_graph(data[n][0]) for n in names],
        config, Reporter(None))
    caller.run()
    return caller.peaks()


def test_run_callpeaks_interface_from_files(tmp_path):
    data = chromosomes_on_disk(tmp_path, names=("1",))
    g, s, c = data["1"]
    d = str(tmp_path) + "/"
    args = args_of(graph=ci.load_graph(d + "1.nobg"), graph_file_name=d + "1.nobg",
                   sample=d + "sample_1.json", control=d + "control_1.json", linear_map=None,
                   fragment_length=str(F), read_length=str(R), out_name=d + "single_",
                   q_value_threshold=None, sequence_graph=None)
    caller = ci.run_callpeaks_interface(args)
    assert os.path.isfile(d + "1_linear_map.npz")                 # created on the way
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, True
    config.linear_map_name = LinearMap.from_graph(g)
    ref = CallPeaks(g, config, Reporter(None))
    ref.run(UniqueIntervals(s), UniqueIntervals(c))
    assert len(ref.max_path_peaks) > 0
    same_peaks(caller.max_path_peaks, ref.max_path_peaks)
    written = PeakCollection.from_file(d + "single_max_paths.intervalcollection")
    assert shape(written.intervals) == shape(ref.max_path_peaks)


def test_run_callpeaks2_two_graphs(tmp_path):
    data = chromosomes_on_disk(tmp_path)
    d = str(tmp_path) + "/"
    genome = sum(g.size_bp for g, _, _ in data.values())
    args = args_of(graph=[d + "*.nobg"], sample=[d + "sample_*.json"],
                   control=[d + "control_*.json"], fragment_length=str(F), read_length=None,
                   out_name=d + "two_", genome_size=str(genome), unique_reads="3000")
    caller = ci.run_callpeaks2(args)
    want = in_memory(data, global_min=3000 * F / genome)
    got = caller.peaks()
    assert sorted(got) == ["1", "2"] and sum(len(v) for v in want.values()) > 0
    for chrom in want:
        same_peaks(got[chrom], want[chrom])
        written = PeakCollection.from_file(d + "two_%s_max_paths.intervalcollection" % chrom)
        assert shape(written.intervals) == shape(want[chrom])
        # create_ob_graph left <graph>.sequences next to every graph: the peaks' sequences
        # are written too (reference multiplegraphscallpeaks.py:157-166); the test graphs
        # hold nothing but "a"
        fasta = PeakCollection.from_fasta_file(d + "two_%s_sequences.fasta" % chrom)
        assert shape(fasta.intervals) == shape(want[chrom])
        assert all(set(p.sequence) == {"a"} and len(p.sequence) == q.length()
                   for p, q in zip(fasta.intervals, want[chrom]))


def test_whole_genome_in_two_halves(tmp_path):
    # first half up to the p-value files, then every chromosome on its own from those
    # files with the q mapping joined over all of them (callpeaks_interface.py:253-371)
    data = chromosomes_on_disk(tmp_path)
    d = str(tmp_path)
    genome = sum(g.size_bp for g, _, _ in data.values())
    first = args_of(chromosomes="1,2", data_dir=d, sample=d + "/sample.json",
                    control=d + "/control.json", fragment_length=str(F), read_length=str(R),
                    genome_size=str(genome), unique_reads="3000", out_name=d + "/wg_",
                    stop_after_p_values="True")
    half = ci.run_callpeaks_whole_genome(first)
    assert not hasattr(half, "_q_value_mapping")
    for chrom in data:
        for f in ("pvalues_indexes.npy", "pvalues_values.npy", "touched_nodes.npy",
                  "direct_pileup_indexes.npy"):
            assert os.path.isfile(d + "/wg_%s_%s" % (chrom, f)), f
    want = in_memory(data, global_min=3000 * F / genome)
    assert sum(len(v) for v in want.values()) > 0
    for chrom in data:
        second = args_of(chromosome=chrom, data_dir=d + "/", out_name=d + "/wg_",
                         fragment_length=str(F), read_length=str(R), q_threshold=None,
                         variant_maps_path=None)
        caller = ci.run_callpeaks_whole_genome_from_p_values(second)
        assert list(caller.peaks()) == [chrom]
        same_peaks(caller.peaks()[chrom], want[chrom])


# ---- summits ---------------------------------------------------------

def test_cut_around_summit_vs_oracle(tmp_path):
    # the peaks and q-values of a run; every cut against the oracle's restatement of
    # PeakCollection.cut_around_summit (pinned to the live reference in tests/test_summits.py)
    from graph_peak_caller_b200.sparsediffs import SparseValues
    from oracle import summits as osummits
    g = make_graph(40000, 0.04, seed=120, mnp_frac=0.1)
    s = make_reads(g, 4000, R, F, seed=121, peak_frac=0.7, peak_spacing=3000)
    c = make_reads(g, 6000, R, F, seed=122, peak_frac=0.0, dup_frac=0.0)
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, True
    config.linear_map_name = LinearMap.from_graph(g)
    caller = CallPeaks(g, config, Reporter(str(tmp_path / "s_")))
    caller.run(UniqueIntervals(s), UniqueIntervals(c))
    peaks = caller.max_path_peaks
    assert len(peaks) > 3
    q = caller.q_values_pileup
    dense = osummits.dense_track(q.indices, q.values, g.size_bp)
    for window, track in ((60, q), (10, SparseValues.from_sparse_files(str(tmp_path / "s_qvalues")))):
        collection = PeakCollection(list(peaks))
        raw = collection.summits(track, window)
        collection.cut_around_summit(track, window)
        assert len(collection.intervals) == len(peaks)
        for peak, row, cut in zip(peaks, raw.tolist(), collection.intervals):
            summit, length, want = osummits.cut_around_summit(
                g.node_off, g.node_len, g.min_node, dense, peak.start_position.offset,
                peak.end_position.offset, list(peak.region_paths), window)
            assert (row[0], row[1]) == (summit, length) and length == peak.length()
            assert (cut.start_position.offset, cut.end_position.offset,
                    list(cut.region_paths)) == want
            assert cut.length() <= 2 * window and cut.score == peak.score


def test_reference_known_answer_for_summits():
    # reference tests/test_command_line_interface.py:121-137
    from graph_peak_caller_b200.graph import Graph
    from graph_peak_caller_b200.peakcollection import Peak
    from graph_peak_caller_b200.sparsediffs import SparseValues
    g = Graph({1: 7, 2: 4, 3: 7, 4: 4}, {1: [2, 3], 2: [4], 3: [4]})
    qvalues = SparseValues(np.array([0]), np.array([3.0]))
    qvalues.track_size = 22
    collection = PeakCollection([Peak(0, 2, [1, 2], graph=g, score=3)])
    collection.cut_around_summit(qvalues, 2)
    cut = collection.intervals[0]
    assert (cut.start_position.offset, cut.end_position.offset, cut.region_paths) == (2, 6, [1])
    assert PeakCollection([]).summits(qvalues).shape == (0, 6)


def test_get_summits_known_answer(tmp_path):
    # reference tests/test_command_line_interface.py:121-138: create_ob_graph, peaks as FASTA,
    # q-value files, get_summits with window 2 -> Peak(2, 6, [1]) reading "tccc"
    from graph_peak_caller_b200 import analysis_interface as ai
    from graph_peak_caller_b200.peakcollection import Peak
    from graph_peak_caller_b200.peakfasta import PeakFasta
    from graph_peak_caller_b200.sequencegraph import SequenceGraph
    from graph_peak_caller_b200.sparsediffs import SparseValues
    from test_sequences import VG_TEST_GRAPH, write
    d = str(tmp_path) + "/"
    ci.create_ob_graph(args_of(vg_json_file_name=write(tmp_path / "vg_test_graph.json", VG_TEST_GRAPH),
                               out_file_name=d + "testgraph.obg"))
    qvalues = SparseValues(np.array([0]), np.array([3.0]))
    qvalues.track_size = 22
    qvalues.to_sparse_files(d + "test_qvalues")
    sequence_graph = SequenceGraph.from_file(d + "testgraph.obg.sequences")
    PeakFasta(sequence_graph).write_max_path_sequences(
        d + "test_max_paths.fasta", PeakCollection([Peak(0, 2, [1, 2], score=3)]))
    ai.get_summits(args_of(graph=d + "testgraph.obg", peaks_fasta_file=d + "test_max_paths.fasta",
                           q_values_base_name=d + "test_qvalues", window_size="2"))
    result = PeakCollection.from_fasta_file(d + "test_max_paths_summits.fasta")
    assert result.intervals[0] == Peak(2, 6, [1])
    assert result.intervals[0].sequence.lower() == "tccc"


def test_count_unique_reads_from_files(tmp_path, capsys):
    # preprocess_interface.py:41-48
    from oracle import pipeline as opipeline
    data = chromosomes_on_disk(tmp_path)
    d = str(tmp_path) + "/"
    got = ci.count_unique_reads(["1", "2"], [d + "1.nobg", d + "2.nobg"],
                                [d + "sample_1.json", d + "sample_2.json"])
    want = sum(len(opipeline.unique_reads(s).start_node) for _, s, _ in data.values())
    assert got == want and capsys.readouterr().out.strip() == str(want)
    assert ci.count_unique_reads_interface(args_of(chromosomes="1,2", graphs_location=d,
                                                   reads_base_name=d + "sample_")) == want
