"""Wall time per stage (host + device, synchronised after every stage) over
several passes.  python tools/stage_walltime.py [iters]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from graph_peak_caller_b200 import _lib, kernels, pipeline
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = make_graph(51_000_000, 0.02, seed=20261001)
s = make_reads(g, 5_000_000, 36, 200, seed=1)
c = make_reads(g, 5_000_000, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
lm = LinearMap.from_graph(g)
dg = DeviceGraph(g, lm.as_arrays())
ds, dc = DeviceReads(s), DeviceReads(c)
R, F = 36, 200


def timed(label, acc, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc.setdefault(label, []).append(1e3 * (time.perf_counter() - t))
    return out


acc = {}
for it in range(iters):
    t0 = time.perf_counter()
    si, ns = timed("unique_s", acc, lambda: kernels.unique_reads(ds))
    ci, nc = timed("unique_c", acc, lambda: kernels.unique_reads(dc))
    ev, touched = timed("sample_events", acc, lambda: kernels.sample_events(dg, ds, si, ns, F - R))
    frag, direct = timed("events_to_runs", acc, lambda: kernels.events_to_runs(ev, g.size_bp))
    mv = nc * F / np.float64(dg.linear_length)
    bp, val, _ = timed("control_linear", acc, lambda: kernels.control_linear_track(
        dg, dc, ci, nc, [F, 1000, 10000], F, mv))
    cpos, clam = timed("control_backproject", acc, lambda: kernels.control_backproject(
        dg, touched, bp, val, mv, ns / nc))
    ppos, p = timed("pvalues", acc, lambda: kernels.pvalues(frag[0], frag[1], 1.0, cpos, clam))
    tp, tc = timed("ptable", acc, lambda: kernels.ptable_local(ppos, p, g.size_bp))
    qp, qq = timed("qtable", acc, lambda: kernels.qtable_build(tp, tc, g.size_bp))
    q = timed("qapply", acc, lambda: kernels.qvalues_apply(p, qp, qq))
    thr = timed("threshold", acc, lambda: kernels.threshold(ppos, q, -np.log10(0.05)))
    cl = timed("clean_holes", acc, lambda: kernels.clean_holes(dg, thr[0], thr[1], R, touched))
    raw = timed("max_paths", acc, lambda: kernels.max_paths(dg, cl[0], cl[1], direct[0], direct[1],
                                                             ppos, q, 0))
    ps = timed("peakset_d2h", acc, lambda: pipeline.PeakSet(raw, F))
    acc.setdefault("TOTAL", []).append(1e3 * (time.perf_counter() - t0))
print("%-22s %s" % ("stage", " ".join("%7d" % i for i in range(iters))))
for k, v in acc.items():
    print("%-22s %s" % (k, " ".join("%7.2f" % x for x in v)))
