"""Wall time of whole passes, concurrent control chain on/off, no profiling."""
import sys, time
import torch
sys.path.insert(0, ".")
from graph_peak_caller_b200 import pipeline, kernels
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads
g = make_graph(51_000_000, 0.02, seed=20261001)
s = make_reads(g, 5_000_000, 36, 200, seed=1)
c = make_reads(g, 5_000_000, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
dg = DeviceGraph(g, LinearMap.from_graph(g).as_arrays()); ds, dc = DeviceReads(s), DeviceReads(c)
def one(conc):
    st = pipeline.run_to_p_values(dg, ds, dc, 36, 200, True, None, True)
    tp, tc = pipeline.local_p_table(st.p_values, st.track_size)
    qp, qq = pipeline.q_table(tp, tc, st.track_size)
    _, _, _, raw = pipeline.run_from_p_values(dg, st.p_values, st.touched, st.direct, qp, qq, 36, 200)
    return pipeline.PeakSet(raw, 200)
for conc in (False,):
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t = time.perf_counter(); p = one(conc); torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t))
    print("concurrent=%s: %s  peaks %d" % (conc, " ".join("%.1f" % x for x in ts), len(p)))
