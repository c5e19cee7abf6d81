"""Host-side (Python) time of one device pass: cProfile over pipeline.callpeaks."""
import sys, time, cProfile, pstats
import torch
sys.path.insert(0, ".")
from graph_peak_caller_b200 import pipeline
from graph_peak_caller_b200.control.linearmap import LinearMap
from graph_peak_caller_b200.device import DeviceGraph, DeviceReads
from graph_peak_caller_b200.synthetic import make_graph, make_reads
g = make_graph(51_000_000, 0.02, seed=20261001)
s = make_reads(g, 5_000_000, 36, 200, seed=1)
c = make_reads(g, 5_000_000, 36, 200, seed=2, peak_frac=0.0, dup_frac=0.0)
dg = DeviceGraph(g, LinearMap.from_graph(g).as_arrays())
ds, dc = DeviceReads(s), DeviceReads(c)
for _ in range(4): pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): pipeline.callpeaks(dg, ds, dc, 36, 200, True, None, True)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
