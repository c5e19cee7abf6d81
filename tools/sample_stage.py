"""Sample-pileup stage alone on the chr22-scale workload, for A/B runs of its kernels.
  python tools/sample_stage.py prep     # generate once, save graph and reads under /tmp
  python tools/sample_stage.py run TAG  # load, time gpc_sample_pileup (environment picks the variant)
"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import _lib, kernels, synthetic
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.graph import Graph
from graph_peak_caller_b200.reads import ReadBatch

G, S = "/tmp/ss_graph.npz", "/tmp/ss_reads.npz"
if sys.argv[1] == "prep":
    cfg = synthetic.CONFIGS["chr22_scale"]
    g, s, c = synthetic.chromosome_inputs(cfg, 0)
    g.to_file(G)
    s.to_file(S)
    print("saved", g, len(s))
else:
    tag = sys.argv[2] if len(sys.argv) > 2 else ""
    g = Graph.from_file(G)
    s = ReadBatch.from_file(S)
    dg, ds = DeviceGraph(g), DeviceReads(s)
    si, ns = kernels.unique_reads(ds)
    for it in range(8):
        _lib.profile_enable(it >= 3)
        out = kernels.sample_pileup(dg, ds, si, ns, 164)
    torch.cuda.synchronize()
    rep = _lib.profile_report()
    print("%-28s dense_runs %.4f ms  walk %.4f ms  runs %d / %d" % (
        tag, rep["dense_runs"][1] / rep["dense_runs"][0], rep["walk_reads"][1] / rep["walk_reads"][0],
        out[0][0].numel(), out[1][0].numel()), flush=True)
