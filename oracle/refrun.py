"""BASELINE INFRASTRUCTURE -- the reference's own CPU callpeaks, timed.

``bench.py --impl reference`` (and the ``cpu_baseline`` leg of the GPU arm) call
this; it never imports the product package.  Two runners over the same flat
inputs (``synth_core.FlatGraph`` / ``FlatReads``, any object with those arrays):

* ``kind == "reference"``: the UNMODIFIED reference (oracle/_ref, installed by
  oracle/build_ref.py, or /root/reference in the build container) on top of
  oracle/refshim -- ``CallPeaks.run_pre_callpeaks / get_p_values /
  get_p_to_q_values_mapping / get_q_values / call_peaks_from_q_values``
  (graph_peak_caller/callpeaks.py:47-123), i.e. ``CallPeaks.run`` split at its own
  method boundaries so that the stages can be timed;
* ``kind == "port"``: the oracle restatement (oracle/pipeline.py) -- fallback when
  the reference cannot be imported.

Preparation (obg graph, obg intervals, linear map file) is ingest / preprocessing
and is done once, outside the timed region; ``run`` is the timed call.
"""
import os
import sys
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for _p in (_ROOT, os.path.join(_HERE, "refshim")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def reference_kind():
    import refcompat
    if not refcompat.available():
        return "port"
    try:
        refcompat.install()
        import graph_peak_caller  # noqa: F401
        return "reference"
    except Exception:          # noqa: BLE001 -- any import problem: use the port
        return "port"


class Prepared:
    """Everything one chromosome's timed call needs, built outside the timed region."""

    def __init__(self, kind, graph, sample, control, read_length, fragment_length,
                 has_control, global_min, name="chr"):
        self.kind, self.name = kind, name
        self.R, self.F = read_length, fragment_length
        self.has_control, self.global_min = has_control, global_min
        self.graph = graph
        self.n_reads_in = len(sample.start_node) + (len(control.start_node) if control is not None else 0)
        self.graph_bp, self.nodes = int(graph.node_off[-1]), int(graph.n_nodes)
        if kind == "reference":
            from oracle import refbridge
            self.tmp = tempfile.mkdtemp(prefix="gpc_ref_")
            self.rg = refbridge.to_ref_graph(graph)
            self.sample = refbridge.to_ref_intervals(sample, self.rg)
            self.control = (refbridge.to_ref_intervals(control, self.rg)
                            if control is not None else self.sample)
            from graph_peak_caller.control.linearmap import LinearMap
            self.lm_file = os.path.join(self.tmp, "lm.npz")
            LinearMap.from_graph(self.rg).to_file(self.lm_file)
        else:
            from oracle import control as ocontrol
            self.sample, self.control = sample, (control if control is not None else sample)
            self.lm = ocontrol.linear_map_from_graph(graph)


def run(prep):
    """One timed callpeaks pass.  Returns (unique reads, seconds, stage seconds, peaks)."""
    if prep.kind == "reference":
        return _run_reference(prep)
    return _run_port(prep)


def _run_reference(prep):
    import graph_peak_caller as gpc
    from graph_peak_caller.intervals import UniqueIntervals
    from graph_peak_caller.reporter import Reporter
    cwd = os.getcwd()
    os.chdir(prep.tmp)                 # the reference reads its direct pileup back from files
    try:
        config = gpc.Configuration()
        config.read_length, config.fragment_length = prep.R, prep.F
        config.linear_map_name = prep.lm_file
        config.has_control = prep.has_control
        config.global_min = prep.global_min
        stages = {}
        t0 = t = time.perf_counter()
        caller = gpc.CallPeaks(prep.rg, config, Reporter(prep.name + "_"))
        s = UniqueIntervals(prep.sample)
        c = UniqueIntervals(prep.control)
        for name, call in (("pileups", lambda: caller.run_pre_callpeaks(s, c)),
                           ("pvalues", caller.get_p_values),
                           ("qvalues", lambda: (caller.get_p_to_q_values_mapping(),
                                                caller.get_q_values())),
                           ("peaks", caller.call_peaks_from_q_values)):
            call()
            now = time.perf_counter()
            stages[name] = now - t
            t = now
        dt = time.perf_counter() - t0
        return s.n_reads + c.n_reads, dt, stages, len(caller.max_path_peaks)
    finally:
        os.chdir(cwd)


def _run_port(prep):
    from oracle import pipeline as opipeline
    timings = {}
    t0 = time.perf_counter()
    su = opipeline.unique_reads(prep.sample)
    cu = opipeline.unique_reads(prep.control)
    out = opipeline.callpeaks(prep.graph, prep.lm, su, cu, prep.R, prep.F, prep.has_control,
                              global_min=prep.global_min, timings=timings)
    dt = time.perf_counter() - t0
    timings["dedup"] = dt - sum(timings.values())
    return len(su.start_node) + len(cu.start_node), dt, timings, len(out["peaks"])
