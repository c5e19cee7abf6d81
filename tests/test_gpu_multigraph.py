"""GPU: several chromosome graphs with one joined q mapping, through
MultipleGraphsCallpeaks, against the oracle's restatement of the reference's
file-based join (sparsepvalues.py:76-93)."""
import numpy as np
import pytest

from graph_peak_caller_b200 import Configuration
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.multiplegraphscallpeaks import MultipleGraphsCallpeaks
from graph_peak_caller_b200.reporter import Reporter
from graph_peak_caller_b200.synthetic import make_graph, make_reads
from oracle import control as ocontrol, pipeline as opipeline, stats as ostats
from parity_util import assert_exact, assert_step_close
from test_gpu_postprocess import assert_same_peaks

pytestmark = pytest.mark.gpu


def test_three_chromosomes_joined_q_values():
    R, F = 15, 70
    names, graphs, samples, controls, lms = [], [], [], [], []
    for k in range(3):
        g = make_graph(12000 + 5000 * k, 0.03, seed=60 + k, mnp_frac=0.1)
        names.append("chr%d" % (k + 1))
        graphs.append(g)
        samples.append(make_reads(g, 1200 + 300 * k, R, F, seed=70 + k, peak_frac=0.7,
                                  peak_spacing=3000))
        controls.append(make_reads(g, 2500, R, F, seed=80 + k, peak_frac=0.0, dup_frac=0.0))
        lms.append(LinearMap.from_graph(g))
    config = Configuration()
    config.read_length, config.fragment_length, config.has_control = R, F, True
    caller = MultipleGraphsCallpeaks(names, graphs, samples, controls, lms, config,
                                     Reporter(None))
    caller.run()
    # oracle: per chromosome to p-values, joined table, per chromosome from p-values
    states = []
    for g, s, c in zip(graphs, samples, controls):
        lm = ocontrol.linear_map_from_graph(g)
        states.append(opipeline.run_to_p_values(g, lm, opipeline.unique_reads(s),
                                                opipeline.unique_reads(c), R, F, True))
    sp, cum = opipeline.joined_p_table([(st["p_values"][0], st["p_values"][1], st["track_size"])
                                        for st in states])
    table = ostats.q_table(sp, cum)
    n_peaks = 0
    for name, g, st in zip(names, graphs, states):
        ref = opipeline.run_from_p_values(g, st, table, R, F)
        mine = caller._callers[names.index(name)]
        q = mine.q_values_pileup
        assert_step_close((q.indices, q.values), ref["q_values"], name + " q", upto=g.size_bp)
        h = mine._reporter.memory["hole_cleaned"]
        assert_exact((h.indices, h.values.astype(bool)), ref["hole_cleaned"], name + " holes")
        got = [(p.start_position.offset, p.end_position.offset, list(p.region_paths),
                (bool(p.info[0]), bool(p.info[1])), float(p.score))
               for p in caller.peaks()[name]]
        assert_same_peaks(got, ref["peaks"])
        n_peaks += len(got)
    assert n_peaks > 0
    # the joined mapping differs from a per-chromosome one (it really is global)
    single = ostats.q_table(*opipeline.p_table_of(*states[0]["p_values"], states[0]["track_size"]))
    assert any(abs(single[k] - table[k]) > 1e-9 for k in single if k in table and k != 0)
